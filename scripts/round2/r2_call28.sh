#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c28_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c28_pytest.txt | cut -c1-300
bash scripts/measure_configs.sh ours 1 2>&1 | tee gpurun_out/c28_configs_ours.txt
