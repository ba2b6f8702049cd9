// Split-K variant of the plain tcgen05 GEMM for small-M / deep-K problems (default for K >= 1024 with at most 16 output
// tiles, e.g. fc1 of the reference CNNs; verified on B200, RLR_SPLITK=0 turns it off).
//
// The first dense layer of the reference's FMNIST CNN is out[256][128] = x[256][9216] W[128][9216]^T: two 128 x 128 output tiles
// that each walk 144 k-blocks in sequence (~40 us of pure pipeline latency on two of 148 SMs); at the runner's batch size 64 it is a
// single tile.  Here grid.z CTAs share a tile's k range, each accumulates its slice in TMEM and adds it into an fp32 workspace
// [M][N] with red.global.add.v4.f32 (zeroed by the caller); `splitk_finish_kernel` then applies bias / ReLU and packs bf16.
// Same roles and operand layouts as gemm.cu (warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue; K-major SW128 operands).
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"

namespace rlr {

using namespace umma;

cudaError_t make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, const uint32_t* elem_strides = nullptr);  // gemm.cu

namespace {

constexpr int SBM = 128, SBN = 128, SBK = 64, SThreads = 192, SStages = 4;
constexpr int SStageBytes = (SBM + SBN) * SBK * 2;          // 32 KB
constexpr int SSmem = SStages * SStageBytes + 1024 + 1024;

struct SplitKParams {
    int M, N, num_kb, kb_per_split;
    float* ws;                 // [M][N] fp32, zero on entry
};

struct __align__(8) SKShared {
    uint64_t full[SStages];
    uint64_t empty[SStages];
    uint64_t tmem_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(SThreads, 1)
umma_gemm_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SplitKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    SKShared* sh = reinterpret_cast<SKShared*>(smem + SStages * SStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int kb0 = blockIdx.z * p.kb_per_split;
    const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
    const int nkb = kb1 - kb0;                               // >= 1 by construction of the grid

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SStages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        mbar_init(&sh->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, SBN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = sh->tmem_base;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&sh->empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * SStageBytes;
                mbar_expect_tx(&sh->full[stage], SStageBytes);
                tma_load_2d(&tmA, &sh->full[stage], sa, (kb0 + i) * SBK, tile_m * SBM);
                tma_load_2d(&tmB, &sh->full[stage], sa + SBM * SBK * 2, (kb0 + i) * SBK, tile_n * SBN);
                if (++stage == SStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(SBM, SBN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&sh->full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * SStageBytes);
                const uint32_t sb = sa + SBM * SBK * 2;
#pragma unroll
                for (int k = 0; k < SBK / 16; ++k)
                    umma_bf16(tmem_acc, smem_desc_sw128(sa + k * 32, 16, 1024), smem_desc_sw128(sb + k * 32, 16, 1024), idesc,
                              (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(&sh->empty[stage]);
                if (++stage == SStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&sh->tmem_full);
        }
    } else {
        const int lane_base = (warp & 3) * 32;
        const int row = tile_m * SBM + lane_base + lane;
        mbar_wait(&sh->tmem_full, 0);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < SBN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)lane_base << 16) + c0, v);
            const int col = tile_n * SBN + c0;
            if (row < p.M) {
                float* dst = p.ws + (size_t)row * p.N + col;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (col + j < p.N)          // N % 8 == 0: whole 4-groups are valid or not
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                                     "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_acc, SBN);
}

// out = bf16(act(ws + bias)); also re-zeroes the workspace so the next call needs no memset
// drop.thr != 0: dropout fused into this finishing pass (keep-mask of output element 4 q + i, common.cuh: dropout_keep8)
__global__ void __launch_bounds__(256) splitk_finish_kernel(float* __restrict__ ws, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                              long long n4, int N, int relu, DropSpec drop) {
    pdl_wait();
    pdl_trigger();
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        float4 v = *reinterpret_cast<const float4*>(ws + 4 * q);
        *reinterpret_cast<float4*>(ws + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) {
            const int c = (int)((4 * q) % N);
            v.x += bias[c]; v.y += bias[c + 1]; v.z += bias[c + 2]; v.w += bias[c + 3];
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (drop.thr) {
            const uint32_t keep = dropout_keep8(drop, q >> 1) >> ((q & 1) * 4);       // elements 4 q .. 4 q + 3 of their group of eight
            v.x = (keep & 1) ? v.x * drop.scale : 0.f; v.y = (keep & 2) ? v.y * drop.scale : 0.f;
            v.z = (keep & 4) ? v.z * drop.scale : 0.f; v.w = (keep & 8) ? v.w * drop.scale : 0.f;
        }
        *reinterpret_cast<uint2*>(out + 4 * q) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

}  // namespace

// out[M][N] (bf16) = act(A[M][K] B[N][K]^T + bias) through `ws` ([M][N] fp32, ZERO on entry; left zero on exit).  K % 64 == 0, N % 8 == 0.
cudaError_t launch_gemm_splitk_bf16(const void* A, const void* B, void* out, float* ws, int M, int N, int K, const float* bias, int relu,
                                    int num_sms, cudaStream_t st, const DropSpec* drop) {
    if (K % SBK || N % 8 || M <= 0) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        RLR_CUDA_CHECK(cudaFuncSetAttribute(umma_gemm_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SSmem));
        configured = true;
    }
    CUtensorMap tmA, tmB;
    {
        const uint64_t d[2] = {(uint64_t)K, (uint64_t)M}, s[1] = {(uint64_t)K * 2};
        const uint32_t b[2] = {SBK, SBM};
        RLR_CUDA_CHECK(make_tmap_bf16(&tmA, A, 2, d, s, b));
    }
    {
        const uint64_t d[2] = {(uint64_t)K, (uint64_t)N}, s[1] = {(uint64_t)K * 2};
        const uint32_t b[2] = {SBK, SBN};
        RLR_CUDA_CHECK(make_tmap_bf16(&tmB, B, 2, d, s, b));
    }
    SplitKParams p{};
    p.M = M; p.N = N; p.num_kb = K / SBK; p.ws = ws;
    const int m_tiles = (M + SBM - 1) / SBM, n_tiles = (N + SBN - 1) / SBN;
    int splits = num_sms / (m_tiles * n_tiles);                 // one wave
    if (splits > p.num_kb / 2) splits = p.num_kb / 2;            // at least two k-blocks per CTA
    if (splits < 1) splits = 1;
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;   // no empty CTA
    RLR_CUDA_CHECK(launch_kernel(umma_gemm_splitk_kernel, dim3(m_tiles, n_tiles, splits), dim3(SThreads), (size_t)SSmem, st, tmA, tmB, p));
    const long long n4 = (long long)M * N / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > num_sms * 4) blocks = num_sms * 4;
    return launch_kernel(splitk_finish_kernel, dim3((int)blocks), dim3(256), (size_t)0, st, ws, bias, reinterpret_cast<__nv_bfloat16*>(out), n4, N,
                         relu, drop ? *drop : DropSpec{});
}

}  // namespace rlr
