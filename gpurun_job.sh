set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --only-cull > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cull_pages -s 70 -c 3 -o gpurun_out/cull_r1 python bench.py --steps 20 --warmup 3 --only-cull > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
