#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c8_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c8_pytest.txt | cut -c1-300
grep -E "FAILED|Error" gpurun_out/c8_pytest.txt | head
timeout 300 python scripts/bench_convs.py --variants default,notmastore,split --layers l2,l3,l4,l2s,l3d --dirs fwd,dgrad 2>&1 | tee gpurun_out/c8_bench_convs.txt
timeout 200 python scripts/trace_conv.py --layers l3,l4 --dirs fwd 2>&1 | tee gpurun_out/c8_trace.txt
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c8_bench_$name.json 2> gpurun_out/c8_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c8_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round handoff={d['config'].get('fused_handoff')} fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c8_bench_{n}.err").read()[-800:])
PY
}
b default NONE=1
b notmastore RLR_TMA_STORE=0
b split RLR_SPLIT_PRODUCER=1
b default2 NONE=1
