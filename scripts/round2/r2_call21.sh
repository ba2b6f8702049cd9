#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c21_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c21_pytest.txt | cut -c1-300
grep -E "FAILED|^E  |Timeout" gpurun_out/c21_pytest.txt | head
python scripts/bench_small_gemm.py 2>&1 | tee gpurun_out/c21_small_gemm.txt
bash scripts/measure_configs.sh ours 1 2>&1 | tee gpurun_out/c21_configs_ours.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/c21_bench_headline.json 2> gpurun_out/c21_bench_headline.err; tail -1 gpurun_out/c21_bench_headline.json | cut -c1-300
