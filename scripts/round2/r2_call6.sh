#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c6_pytest.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c6_pytest.txt | cut -c1-300
grep -E "FAILED|rms-rel difference" gpurun_out/c6_pytest.txt | head -20
RLR_PDL=1 timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -k "fp32_autograd or learns or handoff or bn_kernels or conv" > gpurun_out/c6_pytest_pdl.txt 2>&1; echo "PDL net tests rc=$?"; tail -4 gpurun_out/c6_pytest_pdl.txt | cut -c1-300
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c6_bench_$name.json 2> gpurun_out/c6_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c6_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round handoff={d['config'].get('fused_handoff')} fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c6_bench_{n}.err").read()[-800:])
PY
}
b default NONE=1
b pdl RLR_PDL=1
b pdl_pair RLR_PDL=1 RLR_CONV_2CTA=2
b pair2 RLR_CONV_2CTA=2
b default2 NONE=1
timeout 300 python scripts/graph_overhead.py 2>&1 | tail -5
timeout 200 python scripts/trace_conv.py --layers l2,l3 --dirs fwd 2>&1 | grep "==" | tee gpurun_out/c6_trace_gap.txt
timeout 200 python scripts/trace_conv.py --layers l3 --dirs fwd --pair 2 2>&1 | grep "==" | tee -a gpurun_out/c6_trace_gap.txt
RLR_WGRAD_OVERLAP=1 timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -k "fp32_autograd or learns" > gpurun_out/c6_pytest_overlap.txt 2>&1; echo "overlap net tests rc=$?"; tail -3 gpurun_out/c6_pytest_overlap.txt | cut -c1-300
b overlap RLR_WGRAD_OVERLAP=1
b overlap_pdl RLR_WGRAD_OVERLAP=1 RLR_PDL=1
b default3 NONE=1
