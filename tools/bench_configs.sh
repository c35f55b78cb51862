#!/bin/bash
# Every bench configuration of BASELINE.json on ONE GPU + the host DMA probe: tools/bench_configs.sh <tag>
mkdir -p gpurun_out
tag=${1:-bc}
python tools/host_dma_probe.py --gpus 1 > gpurun_out/${tag}_host_dma_1.jsonl 2> gpurun_out/${tag}_host_dma_1.err
timeout 600 python bench.py --config single --frames 1000 > gpurun_out/${tag}_single.json 2> gpurun_out/${tag}_single.err
timeout 600 python bench.py --config mixed --steps 5 > gpurun_out/${tag}_mixed.json 2> gpurun_out/${tag}_mixed.err
timeout 600 python bench.py --config 512 --steps 5 > gpurun_out/${tag}_512_n1.json 2> gpurun_out/${tag}_512_n1.err
timeout 600 python bench.py --config 1080p --steps 10 > gpurun_out/${tag}_1080p.json 2> gpurun_out/${tag}_1080p.err
timeout 600 python bench.py --config 2160p --steps 5 > gpurun_out/${tag}_2160p.json 2> gpurun_out/${tag}_2160p.err
for f in single mixed 512_n1 1080p 2160p; do echo "== $f"; tail -c 300 gpurun_out/${tag}_$f.err; head -c 420 gpurun_out/${tag}_$f.json; echo; done
cat gpurun_out/${tag}_host_dma_1.jsonl
