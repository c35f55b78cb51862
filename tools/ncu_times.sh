#!/bin/bash
# per-kernel durations of one warm step (S = 64, rendered frames): tools/ncu_times.sh <tag> [regex]
mkdir -p gpurun_out
tag=$1; regex=${2:-k_}
export LT_BENCH_SYNTH=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:$regex --launch-skip 60 -c 22 --csv --log-file gpurun_out/${tag}_times.csv python tools/morph_bench.py --one > gpurun_out/${tag}_times.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/${tag}_times.csv') if not l.startswith('=='))]
agg={}
for r in rows:
    agg.setdefault((r['ID'],r['Kernel Name'][:44],r['Grid Size']),{})[r['Metric Name']]=r['Metric Value']+' '+r['Metric Unit']
for k,v in agg.items():
    print(k[1],k[2],' | '.join(v.get(m,'') for m in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active')))
PY
