#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/N2f_multigpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/N2f_multigpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/N2f_bench.json 2> gpurun_out/N2f_bench.err; echo "bench rc $?" >> gpurun_out/N2f_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/N2f_bench_reference.json 2> gpurun_out/N2f_bench_reference.err; echo "ref rc $?" >> gpurun_out/N2f_bench_reference.err
tail -3 gpurun_out/N2f_multigpu_tests.log; tail -2 gpurun_out/N2f_bench.err; head -c 300 gpurun_out/N2f_bench.json; tail -2 gpurun_out/N2f_bench_reference.err; head -c 200 gpurun_out/N2f_bench_reference.json
