#!/bin/bash
# One multi-GPU gpurun call: host DMA probe, configs[3] (512 streams, strong scaling) and the default line on N GPUs.
#   tools/multi_gpu_session.sh <tag> <N> [what...]     what: probe 512 default
mkdir -p gpurun_out
tag=$1; N=$2; shift; shift
for w in "$@"; do
  case $w in
    probe) timeout 600 python tools/host_dma_probe.py --gpus $N > gpurun_out/${tag}_host_dma_${N}.jsonl 2> gpurun_out/${tag}_host_dma_${N}.err; cat gpurun_out/${tag}_host_dma_${N}.jsonl;;
    512) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 512 --steps 5 > gpurun_out/${tag}_512_n${N}.json 2> gpurun_out/${tag}_512_n${N}.err; tail -c 200 gpurun_out/${tag}_512_n${N}.err; head -c 600 gpurun_out/${tag}_512_n${N}.json; echo;;
    default) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --no-cpu-baseline > gpurun_out/${tag}_default_n${N}.json 2> gpurun_out/${tag}_default_n${N}.err; tail -c 200 gpurun_out/${tag}_default_n${N}.err; head -c 300 gpurun_out/${tag}_default_n${N}.json; echo;;
  esac
done
