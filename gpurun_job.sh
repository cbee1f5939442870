set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -q -m gpu 2>&1 | tail -8
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; tail -c 1500 gpurun_out/bench_v4.err; python -c "
import json; j=json.load(open('gpurun_out/bench_v4.json'))
print('cull', j['value'], j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'])
for k,v in j.get('paths',{}).items(): print(k, v['value'], v['unit'], v['ms_per_step'], v['roofline']['frac'])
print(j.get('paths_error'))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-600
