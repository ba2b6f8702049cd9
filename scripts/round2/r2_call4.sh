#!/bin/bash
# Round-2 GPU call 4: per-CTA timeline of the generic conv kernel; bisect the streaming-mode regression; gather_im2col timing
mkdir -p gpurun_out
timeout 300 python scripts/trace_conv.py --layers l2,l3,l4 --dirs fwd 2>&1 | tee gpurun_out/c4_trace.txt
timeout 120 python scripts/trace_conv.py --layers l3 --dirs fwd --occ3 0 2>&1 | tee -a gpurun_out/c4_trace.txt
T=tests/test_gpu_native.py::test_streaming_mode_after_resident_rounds_uses_valid_indices
for v in "NONE=1" "RLR_SPLITK=0" "RLR_HEAD_V2=0" "RLR_IM2COL_STEM=0" "RLR_SPLITK=0 RLR_HEAD_V2=0 RLR_IM2COL_STEM=0 RLR_BN_RECOMPUTE=0"; do
  env $v timeout 200 python -m pytest $T -m gpu -q -x > gpurun_out/c4_stream_$(echo $v | tr ' =' '__').txt 2>&1; echo "stream test [$v] rc=$?  $(grep -E 'assert \(' gpurun_out/c4_stream_$(echo $v | tr ' =' '__').txt | head -1 | cut -c1-120)"
done
timeout 600 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c4_pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c4_pytest.txt | cut -c1-300
timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; tail -1 gpurun_out/c4_bench.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c4_launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/c4_profile_native.log 2>&1; tail -1 gpurun_out/c4_profile_native.log
grep -E "gather_im2col|stem|linear_small|pad_rows|unpad" gpurun_out/c4_launches_native.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -12
