#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/trace_conv.py --layers l2,l3,l4 --dirs fwd 2>&1 | tee gpurun_out/c7_trace.txt
timeout 200 python scripts/trace_conv.py --layers l3,l4 --dirs fwd --split 1 2>&1 | tee -a gpurun_out/c7_trace.txt
timeout 200 python scripts/trace_conv.py --layers l3 --dirs fwd --occ3 0 --split 1 2>&1 | tee -a gpurun_out/c7_trace.txt
timeout 300 python scripts/bench_convs.py --variants default,split,persistent,persistent+split --layers l2,l3,l4 --dirs fwd,dgrad 2>&1 | tee gpurun_out/c7_bench_convs.txt
timeout 600 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c7_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c7_pytest.txt | cut -c1-300
grep -E "FAILED|unfused-vs-unfused" gpurun_out/c7_pytest.txt | head
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c7_bench_$name.json 2> gpurun_out/c7_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c7_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round handoff={d['config'].get('fused_handoff')} fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c7_bench_{n}.err").read()[-800:])
PY
}
b default NONE=1
b split RLR_SPLIT_PRODUCER=1
b nooverlap RLR_WGRAD_OVERLAP=0
b default2 NONE=1
