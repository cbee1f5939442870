#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/N4_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 200 --warmup 10 > gpurun_out/N4_bench.json 2> gpurun_out/N4_bench.err; echo "bench rc $?" >> gpurun_out/N4_bench.err
tail -3 gpurun_out/N4_bench.err; head -c 400 gpurun_out/N4_bench.json
