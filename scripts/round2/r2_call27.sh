#!/bin/bash
# final verification at HEAD (1 GPU): full GPU tests, smoke, default bench (both arms), agents-in-flight A/B on the large model
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c27_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c27_pytest.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/c27_bench_default.json 2> gpurun_out/c27_bench_default.err; tail -1 gpurun_out/c27_bench_default.json | cut -c1-700
for nf in 1 2; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e --agents 8 --agents_in_flight $nf > gpurun_out/c27_bench_k8_flight$nf.json 2> gpurun_out/c27_bench_k8_flight$nf.err
  tail -1 gpurun_out/c27_bench_k8_flight$nf.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('resnet18 8 agents 1 GPU, agents_in_flight=$nf: %.1f ms/round' % d['ms_per_step'])" || tail -3 gpurun_out/c27_bench_k8_flight$nf.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/c27_bench_reference.json 2> gpurun_out/c27_bench_reference.err; tail -1 gpurun_out/c27_bench_reference.json | cut -c1-300
