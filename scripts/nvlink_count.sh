#!/bin/bash
# NVLink byte accounting of the fused aggregation + broadcast (2 GPUs): NVML link counters of GPU 0 before / after a run of R federated
# rounds of the headline config, against the traffic model of DESIGN.md section 2.2.
#   gpurun --gpus 2 -- 'bash scripts/nvlink_count.sh'
mkdir -p gpurun_out
snap() { nvidia-smi nvlink -gt d -i 0 | python -c "
import re, sys
tx = rx = 0
for l in sys.stdin:
    m = re.search(r'Data (Tx|Rx):\s*(\d+)\s*KiB', l)
    if m:
        if m.group(1) == 'Tx': tx += int(m.group(2))
        else: rx += int(m.group(2))
print(tx, rx)"; }
nvidia-smi nvlink -gt d -i 0 | head -6 > gpurun_out/nvlink_sample.txt
for cfg in "20 fused" "4 fused" "20 nccl"; do set -- $cfg
  b0=$(snap)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps $1 --warmup 3 --no_e2e --backend $2 > gpurun_out/nvlink_bench_$1_$2.json 2> gpurun_out/nvlink_bench_$1_$2.err
  b1=$(snap)
  python - "$1" "$2" $b0 $b1 <<'PY' | tee -a gpurun_out/nvlink_count.txt
import sys
steps, backend, tx0, rx0, tx1, rx1 = sys.argv[1], sys.argv[2], *map(int, sys.argv[3:7])
rounds = int(steps) + 3 + 1                      # warm-up + timed + the aggregation-check round
print(f"backend={backend} rounds={rounds}: GPU0 NVLink data tx {(tx1 - tx0) / 1024:.1f} MiB, rx {(rx1 - rx0) / 1024:.1f} MiB "
      f"-> per round tx {(tx1 - tx0) / 1024 / rounds:.1f} MiB, rx {(rx1 - rx0) / 1024 / rounds:.1f} MiB")
PY
done
cat gpurun_out/nvlink_count.txt
