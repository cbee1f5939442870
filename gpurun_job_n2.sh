#!/bin/bash
# 2 x B200: id-gather push variants, bench.py --gpus 2 with the final defaults
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/N2d_multigpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/N2d_multigpu_tests.log
for G in 2 4 8; do
  LB200_PUSH_GRID=$G timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 profiles/time_gather.py 2>&1 | grep GATHER_TIMES_US >> gpurun_out/N2d_time_gather.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/N2d_bench.json 2> gpurun_out/N2d_bench.err; echo "bench rc $?" >> gpurun_out/N2d_bench.err
tail -3 gpurun_out/N2d_multigpu_tests.log; cat gpurun_out/N2d_time_gather.log; tail -2 gpurun_out/N2d_bench.err; head -c 400 gpurun_out/N2d_bench.json
