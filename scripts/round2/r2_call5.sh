#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/trace_conv.py --layers l3,l4 --dirs fwd --pair 1 2>&1 | tee gpurun_out/c5_trace.txt
timeout 200 python scripts/trace_conv.py --layers l3 --dirs fwd --pair 2 2>&1 | tee -a gpurun_out/c5_trace.txt
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c5_pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c5_pytest.txt | cut -c1-300
grep -E "FAILED|Error" gpurun_out/c5_pytest.txt | head -20
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c5_bench_$name.json 2> gpurun_out/c5_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c5_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round handoff={d['config'].get('fused_handoff')} fallbacks={d.get('library_fallbacks')} phases={d.get('phase_ms_per_round_rank0')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c5_bench_{n}.err").read()[-800:])
PY
}
b default NONE=1
b strided RLR_STRIDED_TMA=1
b nohandoff_env NONE=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c5_launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/c5_profile_native.log 2>&1; tail -1 gpurun_out/c5_profile_native.log
grep -E "gather_im2col" gpurun_out/c5_launches_native.csv | awk -F'","' '{print $NF}' | head -3
