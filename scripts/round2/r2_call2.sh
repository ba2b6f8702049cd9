#!/bin/bash
# Round-2 GPU call 2: full GPU suite at HEAD, per-layer conv timings (default vs CTA pair), launch list of one step, ncu of the generic conv
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/c2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c2_pytest.txt | cut -c1-300
grep -E "rms-rel|fell|FAILED|passed|failed" gpurun_out/c2_pytest.txt | cut -c1-400 | head -80
timeout 300 python scripts/bench_convs.py --variants default,pair,occ0,persistent --layers l1,l2,l3,l4,l2s,l3s,l4s,l2d 2>&1 | tee gpurun_out/c2_bench_convs.txt
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -1 gpurun_out/c2_bench.json | cut -c1-1800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c2_launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/c2_profile_native.log 2>&1; tail -2 gpurun_out/c2_profile_native.log
python scripts/summarize_launches.py gpurun_out/c2_launches_native.csv "c2 native step (warm)" | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_conv_gemm -c 12 -o gpurun_out/c2_ncu_conv python scripts/bench_convs.py --eager --layers l2,l3 --dirs fwd --variants default,pair > gpurun_out/c2_ncu_conv.log 2>&1; tail -2 gpurun_out/c2_ncu_conv.log
ls -la gpurun_out | tail -12
