#!/bin/bash
# 2 x B200: the two-GPU tests, the exchange step with 3 / 6 lanes, bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/N2_smi.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/N2_multigpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/N2_multigpu_tests.log
for V in "LB200_CULL_LANES=6" "LB200_CULL_LANES=6 LB200_NO_PDL=1" "LB200_CULL_LANES=3 LB200_NO_PDL=1" "LB200_CULL_LANES=8" "LB200_CULL_LANES=6 LB200_EXCHANGE_PIPELINED=1 LB200_NO_PDL=1"; do
  T=$(echo "$V" | tr ' =' '__')
  env $V timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/time_exchange.py > gpurun_out/N2_time_exchange_$T.log 2>&1
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/N2_bench.json 2> gpurun_out/N2_bench.err; echo "bench rc $?" >> gpurun_out/N2_bench.err
tail -3 gpurun_out/N2_multigpu_tests.log; grep 'from C' gpurun_out/N2_time_exchange_*.log; tail -3 gpurun_out/N2_bench.err; head -c 1500 gpurun_out/N2_bench.json
