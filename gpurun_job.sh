mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -x 2>&1 | tail -15
for mode in 0 1; do
LB200_NO_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_n2_$mode.err; python -c "
import json
for l in open('gpurun_out/bench_n2_$mode.json'):
    if l.startswith('{'):
        j=json.loads(l); print('N2 nop2p=$mode', j['value'], j['ms_per_step'], j['gpu_launches'], j['config']['parallelism'])"
done
