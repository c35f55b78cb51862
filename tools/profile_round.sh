set -x
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
tail -c 600 gpurun_out/r01_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on --launch-skip 75 -c 24 -f -o gpurun_out/r01_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out/
