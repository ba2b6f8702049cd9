#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "sm100_vs_fp32 or learns or handoff" > gpurun_out/c26_pytest_branch.txt 2>&1; echo "pytest (fwd+bwd branch on) rc=$?"; tail -2 gpurun_out/c26_pytest_branch.txt | cut -c1-200
b() {   # name, env..., -- bench args
    name=$1; shift
    envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c26_bench_$name.json 2> gpurun_out/c26_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c26_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c26_bench_{n}.err").read()[-800:])
PY
}
b headline_branch1 RLR_BWD_BRANCH=1 --
b headline_branch0 RLR_BWD_BRANCH=0 --
b headline_branch1b RLR_BWD_BRANCH=1 --
b headline_branch0b RLR_BWD_BRANCH=0 --
b resnet34_branch1 RLR_BWD_BRANCH=1 -- --model resnet34
b resnet34_branch0 RLR_BWD_BRANCH=0 -- --model resnet34
