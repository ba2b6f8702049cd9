#!/bin/bash
# Round-2 GPU call 3: HEAD suite, conv variants per layer (pair BN=256 / deep ring / strided TMA), whole-step A/B of the knobs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/c3_pytest.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c3_pytest.txt | cut -c1-300
timeout 400 python scripts/bench_convs.py --variants default,pair,pairdeep,strided --layers l2,l3,l4,l2s,l3s,l4s,l2d,l3d 2>&1 | tee gpurun_out/c3_bench_convs.txt
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c3_bench_$name.json 2> gpurun_out/c3_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c3_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')} notes={d.get('notes')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c3_bench_{n}.err").read()[-600:])
PY
}
b default NONE=1
b nostem RLR_IM2COL_STEM=0
b pair RLR_CONV_2CTA=1
b pairdeep RLR_CONV_2CTA=2
b strided RLR_STRIDED_TMA=1
b pair_strided RLR_CONV_2CTA=1 RLR_STRIDED_TMA=1
b default2 NONE=1
RLR_STRIDED_TMA=1 timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "fp32_autograd or learns" > gpurun_out/c3_pytest_strided.txt 2>&1; echo "strided net tests rc=$?"; tail -4 gpurun_out/c3_pytest_strided.txt | cut -c1-300
RLR_CONV_2CTA=1 timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "fp32_autograd or learns or conv" > gpurun_out/c3_pytest_pair.txt 2>&1; echo "pair net tests rc=$?"; tail -4 gpurun_out/c3_pytest_pair.txt | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c3_launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/c3_profile_native.log 2>&1; tail -2 gpurun_out/c3_profile_native.log
python scripts/summarize_launches.py gpurun_out/c3_launches_native.csv "c3 native step (warm)" | head -36
