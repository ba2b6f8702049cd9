#!/bin/bash
# One short GPU call: numerics of the opt-in generic-conv schedulers, then the headline bench with one of them on and off.
# usage: persist_check.sh [ENV_FLAG]   (default RLR_PERSISTENT_CONV; e.g. RLR_CONV_OCC3)
FLAG=${1:-RLR_PERSISTENT_CONV}
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_native.py -m gpu -x -q -k "persistent_conv" > gpurun_out/persist_test.log 2>&1
echo "test exit $?" >> gpurun_out/persist_test.log
tail -3 gpurun_out/persist_test.log
env $FLAG=1 timeout 45 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/${FLAG}_on.json 2> gpurun_out/${FLAG}_on.err
tail -1 gpurun_out/${FLAG}_on.json | cut -c1-260
timeout 45 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/${FLAG}_off.json 2> gpurun_out/${FLAG}_off.err
tail -1 gpurun_out/${FLAG}_off.json | cut -c1-260
