#!/bin/bash
# One short GPU call: numerics of the opt-in persistent conv scheduler, then the headline bench with it on and off.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_native.py -m gpu -x -q -k "persistent_conv" > gpurun_out/persist_test.log 2>&1
echo "test exit $?" >> gpurun_out/persist_test.log
tail -3 gpurun_out/persist_test.log
RLR_PERSISTENT_CONV=1 timeout 90 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/persist_bench_on.json 2> gpurun_out/persist_bench_on.err
tail -1 gpurun_out/persist_bench_on.json
timeout 90 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/persist_bench_off.json 2> gpurun_out/persist_bench_off.err
tail -1 gpurun_out/persist_bench_off.json
