#!/bin/bash
# Round-2 opener: ONE 1-GPU call that validates and measures every opt-in path written blind at the end of round 1.
#   gpurun --timeout 1500 -- 'bash scripts/r2_experiments.sh'
# Each experiment: its numerics test (own process, own timeout: a deadlocked kernel only loses that experiment), then -- only if
# the test passed -- the headline bench with the flag on.  The default build is benched first and last (box drift); finally all
# passing flags together.
# Outputs: gpurun_out/r2_exp_<name>.txt, gpurun_out/r2_bench_<name>.json, summary in gpurun_out/r2_summary.txt
mkdir -p gpurun_out
: > gpurun_out/r2_summary.txt
bench() {   # name, env assignments (one word each)
    name=$1; shift
    env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
    python - "$name" <<'PY' | tee -a gpurun_out/r2_summary.txt
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{name}.json").read().strip().splitlines()[-1])
    print(f"bench {name}: {d['ms_per_step']:.1f} ms/round  {d['value']:.4f} rounds/s")
except Exception as e:
    print(f"bench {name}: FAILED ({e})")
PY
}
bench default_first NONE=1
good=""
#   name          pytest -k expression          flag
while read -r name expr flag; do
    [ -z "$name" ] && continue
    timeout 150 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -k "$expr" > gpurun_out/r2_exp_$name.txt 2>&1
    rc=$?
    echo "test $name: exit $rc ($(tail -1 gpurun_out/r2_exp_$name.txt))" | tee -a gpurun_out/r2_summary.txt
    if [ $rc -eq 0 ]; then
        bench $name $flag
        good="$good $flag"
    fi
done <<'LIST'
strided      strided_tma                  RLR_STRIDED_TMA=1
stem         im2col_stem                  RLR_IM2COL_STEM=1
bnmask       recomputed_relu_mask         RLR_BN_RECOMPUTE=1
occ3x        occ3_level2                  RLR_CONV_OCC3=2
head         head_kernels_v2              RLR_HEAD_V2=1
splitk       splitk_gemm                  RLR_SPLITK=1
halo3        halo3_kernel                 RLR_HALO3=1
pdl          programmatic_dependent       RLR_PDL=1
pair         cta_pair                     RLR_CONV_2CTA=1
LIST
[ -n "$good" ] && bench combined $good      # every experiment whose test passed, together
# concurrency of several agents per GPU (the reference's README workload: FMNIST CNN, 10 agents on one GPU)
timeout 200 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -k agents_in_flight > gpurun_out/r2_exp_inflight.txt 2>&1
echo "test inflight: exit $? ($(tail -1 gpurun_out/r2_exp_inflight.txt))" | tee -a gpurun_out/r2_summary.txt
timeout 200 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -k deeper_family > gpurun_out/r2_exp_family.txt 2>&1
echo "test resnet34/vgg16 on native kernels: exit $? ($(grep "logit rel" gpurun_out/r2_exp_family.txt | tr '\n' ';') $(tail -1 gpurun_out/r2_exp_family.txt))" | tee -a gpurun_out/r2_summary.txt
for v in 0 1; do
    RLR_GEMM_SMALL_BN64=$v timeout 120 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -k small_batch_gemm 2>&1 \
        | grep "gemm \|passed\|failed" | sed "s/^/RLR_GEMM_SMALL_BN64=$v  /" | tee -a gpurun_out/r2_summary.txt
done
readme="--model cnn_mnist --data fmnist --train_size 60000 --agents 10 --steps 3 --warmup 3 --no_e2e"
timeout 600 python bench.py --impl reference $readme > gpurun_out/r2_readme_reference.json 2> gpurun_out/r2_readme_reference.err
echo "README workload (FMNIST CNN, 10 agents, 1 GPU) reference: $(tail -1 gpurun_out/r2_readme_reference.json | cut -c1-160)" | tee -a gpurun_out/r2_summary.txt
for nf in 1 2 4; do
    timeout 200 python bench.py $readme --agents_in_flight $nf > gpurun_out/r2_readme_ours_$nf.json 2> gpurun_out/r2_readme_ours_$nf.err
    echo "README workload ours, agents_in_flight=$nf: $(tail -1 gpurun_out/r2_readme_ours_$nf.json | cut -c1-160)" | tee -a gpurun_out/r2_summary.txt
done
bench default_last NONE=1
cat gpurun_out/r2_summary.txt
