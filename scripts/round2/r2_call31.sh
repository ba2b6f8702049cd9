#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x > gpurun_out/c31_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c31_pytest.txt | cut -c1-300
