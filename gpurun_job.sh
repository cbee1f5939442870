set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -c 1500 gpurun_out/bench_v3.err; python -c "
import json; j=json.load(open('gpurun_out/bench_v3.json'))
print('cull', j['value'], j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'])
for k,v in j.get('paths',{}).items(): print(k, v['value'], v['unit'], v['ms_per_step'], v['roofline']['frac'])
print(j.get('paths_error'), j.get('cpu_baseline'))
"
