set -x
mkdir -p gpurun_out
python -m pytest tests/test_cull_gpu.py -q -m gpu 2>&1 | tail -5
python bench.py --steps 200 --warmup 10 --only-cull > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 2000 gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json
ncu --set full --clock-control none --import-source on -k regex:cull_pages -s 70 -c 2 -o gpurun_out/cull_v2 python bench.py --steps 20 --warmup 3 --only-cull > gpurun_out/ncu_full.log 2>&1
