#!/bin/bash
# Round-2 multi-GPU check of whatever flags survived scripts/r2_experiments.sh:
#   gpurun --gpus 2 --timeout 900 -- 'FLAGS="RLR_PDL=1 RLR_STRIDED_TMA=1" bash scripts/r2_multi.sh'
# fused-aggregation tests (default build), then the headline bench at N ranks without and with the flags.
N=${N:-2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multi_tests.txt 2>&1
echo "multi tests exit $? ($(tail -1 gpurun_out/r2_multi_tests.txt))"
run() {   # name, env...
    name=$1; shift
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_multi_$name.json 2> gpurun_out/r2_multi_$name.err
    tail -1 gpurun_out/r2_multi_$name.json | cut -c1-200
}
run default NONE=1
[ -n "$FLAGS" ] && run flags $FLAGS
