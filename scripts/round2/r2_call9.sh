#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c9_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c9_pytest.txt | cut -c1-300
grep -E "FAILED|Error" gpurun_out/c9_pytest.txt | head
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c9_bench_$name.json 2> gpurun_out/c9_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c9_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round handoff={d['config'].get('fused_handoff')} fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c9_bench_{n}.err").read()[-800:])
PY
}
b default NONE=1
b bn_unroll1 RLR_BN_UNROLL=1
b bn_slots1 RLR_BN_SLOTS=1
b bn_old RLR_BN_UNROLL=1 RLR_BN_SLOTS=1
b nosplit RLR_SPLIT_PRODUCER=0
b default2 NONE=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c9_launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/c9_profile_native.log 2>&1; tail -1 gpurun_out/c9_profile_native.log
