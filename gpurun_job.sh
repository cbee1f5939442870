mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --steps 300 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 3 --only-cull > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cull_pages -s 70 -c 2 -o gpurun_out/cull_final python bench.py --steps 20 --warmup 3 --only-cull > gpurun_out/ncu_cull.log 2>&1
python -c "
import json; j=json.load(open('gpurun_out/bench_final.json'))
print('cull', j['value'], j['ms_per_step'], j['roofline'], 'e2e', j['e2e'])
for k,v in j.get('paths',{}).items(): print(k, v['value'], v['unit'], v['ms_per_step'], v['roofline']['frac'])
"
