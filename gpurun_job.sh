python -m pytest tests/test_cull_gpu.py -q -m gpu 2>&1 | tail -2
python profiles/time_cull_variants.py 2>&1 | grep CULLVAR | cut -c1-60
