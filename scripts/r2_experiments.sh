#!/bin/bash
# Round-2 opener: one 1-GPU call that validates and measures every opt-in path written blind at the end of round 1.
#   gpurun --timeout 600 -- 'bash scripts/r2_experiments.sh'
# Outputs: gpurun_out/r2_exp_tests.txt, gpurun_out/r2_bench_<flag>.json
mkdir -p gpurun_out
RLR_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_experimental.py -m gpu -q -x > gpurun_out/r2_exp_tests.txt 2>&1
echo "experimental tests exit $?" | tee -a gpurun_out/r2_exp_tests.txt
tail -5 gpurun_out/r2_exp_tests.txt
for flag in NONE RLR_PDL RLR_STRIDED_TMA RLR_IM2COL_STEM RLR_BN_RECOMPUTE RLR_CONV_OCC3; do
    val=1; [ $flag = RLR_CONV_OCC3 ] && val=2
    env $flag=$val timeout 120 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/r2_bench_$flag.json 2> gpurun_out/r2_bench_$flag.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_$flag.json").read().strip().splitlines()[-1])
    print("$flag", d["ms_per_step"], "ms/round", d["value"], "rounds/s")
except Exception as e:
    print("$flag", "FAILED", e)
PY
done
