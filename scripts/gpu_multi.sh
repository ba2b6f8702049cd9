#!/bin/bash
# multi-GPU check (run with `gpurun --gpus N -- 'N=<N> bash scripts/gpu_multi.sh'`):
#   1. fused P2P aggregation / hand-off / 1000-epoch flag-reuse stress tests (tests/test_gpu_multi.py)
#   2. compute-sanitizer memcheck + racecheck over the 2-rank aggregation worker (all child processes)
#   3. headline bench at N ranks (prints agg_check_max_abs_err / all_ranks_equal from the post-run oracle comparison)
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/pytest_multi.txt 2>&1; echo "multi tests rc=$?"; tail -6 gpurun_out/pytest_multi.txt | cut -c1-300
if [ "${SANITIZE:-1}" = "1" ]; then
  for tool in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 \
        python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "fused_p2p_aggregate_matches_oracle and auto" > gpurun_out/sanitize_multi_$tool.txt 2>&1
    echo "compute-sanitizer $tool over the 2-rank fused aggregation: rc=$? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_multi_$tool.txt | sort | uniq -c | tr '\n' ';')"
  done
fi
for extra in "" "--no_fused_handoff_cli"; do :; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_ours_n$N.json 2> gpurun_out/bench_ours_n$N.err
tail -1 gpurun_out/bench_ours_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench N=%d: %.1f ms/round, %.3f rounds/s, e2e %s, agg_check %s, handoff %s, phases %s' % (d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d.get('agg_check'), d['config'].get('fused_handoff'), d.get('phase_ms_per_round_rank0')))"
