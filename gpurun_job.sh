timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -3
for d in 0 7; do
LB200_GATHER_DEBUG=$d timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$d profiles/time_gather.py 2>&1 | grep -E "GATHER_TIMES|illegal" | head -3; echo "debug=$d"
done
