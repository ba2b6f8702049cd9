#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python scripts/graph_overhead.py 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --trainer native --no_e2e 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
RLR_SMALL_BN64=1 timeout 600 python bench.py --steps 3 --warmup 3 --trainer native --no_e2e 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
