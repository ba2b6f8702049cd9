#!/bin/bash
mkdir -p gpurun_out
b() {   # name, env..., -- bench args
    name=$1; shift
    envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c24_bench_$name.json 2> gpurun_out/c24_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c24_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c24_bench_{n}.err").read()[-800:])
PY
}
b headline_stats_off RLR_EPILOGUE_BN_STATS=0 --
b headline_stats_s1 RLR_EPILOGUE_BN_STATS=1 RLR_EPI_STAT_SLOTS=1 --
b headline_stats_s2 RLR_EPILOGUE_BN_STATS=1 RLR_EPI_STAT_SLOTS=2 --
b headline_stats_s4 RLR_EPILOGUE_BN_STATS=1 RLR_EPI_STAT_SLOTS=4 --
b headline_stats_off2 RLR_EPILOGUE_BN_STATS=0 --
RLR_EPILOGUE_BN_STATS=1 timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "sm100_vs_fp32 or gemm_bf16 or conv2d" > gpurun_out/c24_pytest_stats.txt 2>&1; echo "pytest (stats on) rc=$?"; tail -2 gpurun_out/c24_pytest_stats.txt | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c24_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c24_pytest.txt | cut -c1-300
