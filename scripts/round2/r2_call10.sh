#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c10_pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c10_pytest.txt | cut -c1-300
grep -E "FAILED|Error|vs-" gpurun_out/c10_pytest.txt | head -20
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c10_bench_$name.json 2> gpurun_out/c10_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c10_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round handoff={d['config'].get('fused_handoff')} fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c10_bench_{n}.err").read()[-800:])
PY
}
b default NONE=1
b bn_occ RLR_BN_OCC=1
b default2 NONE=1
b bn_occ2 RLR_BN_OCC=1
