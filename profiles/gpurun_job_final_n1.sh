#!/bin/bash
# round-2 final 1-GPU evidence: tests, smoke, both bench arms, ncu launch list, ncu --set full of the kernels DESIGN.md quotes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/I_smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/I_gputests.log 2>&1; echo "pytest rc $?" >> gpurun_out/I_gputests.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/I_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/I_smoke.log
timeout 900 python bench.py > gpurun_out/I_bench_n1.json 2> gpurun_out/I_bench_n1.err; echo "bench rc $?" >> gpurun_out/I_bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/I_bench_reference.json 2> gpurun_out/I_bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/I_launches.csv python bench.py --only-cull --steps 50 --warmup 3 > gpurun_out/I_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cull_pages_kernel -s 40 -c 2 -o gpurun_out/I_cull_full python bench.py --only-cull --steps 20 --warmup 3 > gpurun_out/I_ncu_cull.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'create_keys_kernel|radix_sort_kernel' -s 56 -c 2 -o gpurun_out/I_sortkeys_full python profiles/time_sortkeys.py > gpurun_out/I_ncu_sortkeys.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:pose_palette_kernel -s 10 -c 1 -o gpurun_out/I_pose_full python profiles/time_anim_variants.py > gpurun_out/I_ncu_pose.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:skin_kernel -s 3 -c 1 -o gpurun_out/I_skin_full python profiles/time_anim_variants.py > gpurun_out/I_ncu_skin.log 2>&1
tail -3 gpurun_out/I_gputests.log; tail -2 gpurun_out/I_smoke.log; tail -2 gpurun_out/I_bench_n1.err; head -c 700 gpurun_out/I_bench_n1.json; ls -la gpurun_out | tail -15
