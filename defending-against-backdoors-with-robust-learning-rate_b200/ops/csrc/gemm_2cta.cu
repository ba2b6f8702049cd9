// CTA-pair variant of the implicit-GEMM conv / GEMM kernel of gemm.cu (opt-in: RLR_CONV_2CTA=1).
// Verified on B200 (tests/test_gpu_variants.py); its main loop is faster (layer 3: 11.1 vs 13.6 us) but the 128 x 256 epilogue and the
// single TMA stream per SM make the round 1-3 % slower than the default (profiles/r2_step_ab.md, profiles/r2_ncu_generic_conv.md).
//
// A thread-block cluster of two CTAs (adjacent M tiles, same N tile) runs ONE tcgen05.mma.cta_group::2 stream: the instruction
// has M = 256 (128 rows from each CTA's shared memory, 128 accumulator lanes in each CTA's TMEM) and N = BN, and each CTA stages
// only HALF of the B tile (BN/2 filter rows) -- the tensor cores of the two SMs read the two halves from both shared memories.
// Per SM and MMA that is 4 KB of A + BN/2 x 32 B of B instead of 4 KB + BN x 32 B (the N <= 128 MMAs of gemm.cu are bounded by
// shared-memory operand bandwidth, docs/NOTES_ROUND1.md), and the filter tile is fetched from L2 once per pair.  BN = 256 becomes
// possible (256 TMEM columns, two pairs per SM pair).
//
// Roles per CTA as in gemm.cu (warp 0 TMA producer, warp 1 MMA, warps 2-5 epilogue) with the pair protocol of the CUTLASS sm100
// collectives:  * both producers issue cp.async.bulk.tensor.cta_group::2 loads into their OWN shared memory that complete on the
// LEADER's full barrier (barrier address with the peer bit cleared); only the leader arms it (expect_tx of both CTAs' bytes)
//               * only the leader's MMA thread issues MMAs; its tcgen05.commit.cta_group::2 multicasts the "stage free" /
// "accumulator complete" arrivals to the barriers of both CTAs
//               * TMEM is allocated / freed with the cta_group::2 forms by the same warp of both CTAs, with cluster barriers after
// the mbarrier initialisation and before the deallocation.
// B operands: K-major (forward convolutions, GEMM) or MN-major boxes of the un-transposed forward filter (data gradients, kBMN:
// each CTA loads the 64-wide N groups of ITS half of the tile); no statistics epilogue.
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "umma.cuh"

namespace rlr {

using namespace umma;

namespace {

constexpr int CBM = 128, CBK = 64, CThreads = 192, CMaxStages = 8;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the even (leader) CTA of the pair

// kDeep = 0: ~96 KB of operand ring per CTA, two CTAs (of two different pairs) per SM so one's epilogue overlaps the other's main loop;
// kDeep = 1: the whole shared memory as ONE deep ring (192 KB in flight per SM from a single CTA) -- for grids of at most one CTA per SM
template <int BN, int kDeep>
struct CCfg {
    static constexpr int kABytes = CBM * CBK * 2;               // 16 KB: this CTA's 128 rows
    static constexpr int kBBytes = (BN / 2) * CBK * 2;          // this CTA's half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (kDeep ? 192 : 96) * 1024 / kStageBytes;      // 4 / 3 (BN = 128 / 256), deep: 8 / 6
    static constexpr int kRingBytes = kStages * kStageBytes;
    static constexpr int kPitch = BN * 2 + 16;
    static constexpr int kStagingBytes = CBM * kPitch;          // aliases the ring once the accumulator is complete
    static_assert(kStagingBytes <= kRingBytes, "staging aliases the operand ring");
    static constexpr int kSmemBytes = kRingBytes + 1024 + 1024;
};

struct __align__(8) CShared {
    uint64_t full[CMaxStages];  // used in the leader CTA only (armed by the leader, completed by both CTAs' TMA loads)
    uint64_t empty[CMaxStages]; // in both CTAs: one multicast arrival per consumed stage
    uint64_t tmem_full;         // in both CTAs: accumulator complete
    uint32_t tmem_base;
    uint32_t pad;
    int row_index[CBM];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// loads land in the executing CTA's shared memory; the transaction bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2),
                   "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// one arrival on the barrier at this shared-memory offset in EVERY CTA of `cta_mask` once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

template <int BN, bool kBMN, int kDeep>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CThreads, kDeep ? 1 : 2)
umma_conv_gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGemmParams p) {
    using Cfg = CCfg<BN, kDeep>;
    constexpr int CStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    CShared* sh = reinterpret_cast<CShared*>(smem + Cfg::kRingBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();        // 0 = leader (even M tile), 1 = peer
    const bool leader = cta_rank == 0;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y; // cluster = M tiles (2i, 2i+1); an odd tail tile is fully masked
    long long* dbg = p.dbg ? p.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;   // timeline, see ConvGemmParams::dbg
    if (dbg && threadIdx.x == 0) { dbg[0] = (long long)gtimer(); uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[7] = sm; }

    int n0 = 0, h0 = 0, w0 = 0;
    if (p.mode == 1) {
        const int tw_i = tile_m % p.tiles_w, th_i = (tile_m / p.tiles_w) % p.tiles_h, tn_i = tile_m / (p.tiles_w * p.tiles_h);
        w0 = tw_i * p.TW; h0 = th_i * p.TH; n0 = tn_i * p.TN;
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < CStages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        mbar_init(&sh->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(&sh->tmem_base, BN);
    if (warp >= 2) {
        const int r = threadIdx.x - 64;
        int gi;
        if (p.mode == 1) {
            const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
            const int w = w0 + tw, h = h0 + th, n = n0 + tn;
            gi = (w < p.Wo && h < p.Ho && n < p.NB) ? ((n * p.OutH + h * p.out_stride + p.out_ph) * p.OutW + w * p.out_stride + p.out_pw) : -1;
        } else {
            gi = tile_m * CBM + r;
            if (gi >= p.M) gi = -1;
        }
        sh->row_index[r] = gi;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers are initialised and both TMEM allocations are done before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_acc = sh->tmem_base;
    pdl_wait();
    pdl_trigger();
    if (dbg && threadIdx.x == 0) dbg[1] = (long long)gtimer();

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) =============================================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&sh->empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kABytes;
                if (leader) mbar_expect_tx(&sh->full[stage], 2 * Cfg::kStageBytes);     // bytes of BOTH CTAs land on this barrier
                const int tap = p.mode == 1 ? kb / p.cblocks : 0, cb = p.mode == 1 ? kb - tap * p.cblocks : 0;
                if (p.mode == 1) {
                    tma_load_4d_pair(&tmA, &sh->full[stage], sa, cb * CBK, w0 * p.in_stride + p.dw[tap], h0 * p.in_stride + p.dh[tap],
                                     n0 + p.dn[tap]);
                } else {
                    tma_load_2d_pair(&tmA, &sh->full[stage], sa, kb * CBK, tile_m * CBM);
                }
                const int n_half = tile_n * BN + (int)cta_rank * (BN / 2);          // first B row / column of MY half of the tile
                if (kBMN) {
                    for (int g = 0; g < BN / 128; ++g)
                        tma_load_2d_pair(&tmB, &sh->full[stage], sb + g * 8192, p.wtap[tap] * p.wcols + n_half + g * 64, cb * CBK);
                } else {
                    tma_load_2d_pair(&tmB, &sh->full[stage], sb, kb * CBK, n_half);
                }
                if (dbg && kb == 0) dbg[2] = (long long)gtimer();
                if (++stage == CStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer: one thread of the LEADER CTA =============================================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(2 * CBM, BN, 0, kBMN ? 1 : 0);   // M = 256 across the pair
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&sh->full[stage], phase);
                tc_fence_after();
                if (dbg && kb == 0) dbg[3] = (long long)gtimer();
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
                for (int k = 0; k < CBK / 16; ++k) {
                    // descriptors are offsets in the issuing CTA's shared memory; the peer's operands sit at the same offsets
                    const uint64_t da = smem_desc_sw128(sa + k * 32, 16, 1024);
                    const uint64_t db = kBMN ? smem_desc_sw128(sb + k * 2048, 8192, 1024) : smem_desc_sw128(sb + k * 32, 16, 1024);
                    umma_bf16_pair(tmem_acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit_pair(&sh->empty[stage], 0x3);      // stage free in both CTAs
                if (++stage == CStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(&sh->tmem_full, 0x3);             // accumulators complete in both CTAs
            if (dbg) dbg[4] = (long long)gtimer();
        }
    } else {
        // ================================ epilogue (both CTAs; identical to gemm.cu without statistics) ========================
        const int et = threadIdx.x - 64;
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        mbar_wait(&sh->tmem_full, 0);
        tc_fence_after();
        if (dbg && et == 0) dbg[5] = (long long)gtimer();
        uint8_t* staging = smem;                         // every MMA of the pair has completed: the operand ring is idle
        const int col0 = tile_n * BN;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)lane_base << 16) + c0, v);
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
                if (p.bias && col0 + c0 + j < p.N) { a += p.bias[col0 + c0 + j]; b += p.bias[col0 + c0 + j + 1]; }
                if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                packed[j >> 1] = pack_bf16x2(a, b);
            }
            uint4* dst = reinterpret_cast<uint4*>(staging + row * Cfg::kPitch + c0 * 2);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        constexpr int kChunks = BN * 2 / 16;
        constexpr int kIters = CBM * kChunks / 128;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
        for (int it0 = 0; it0 < kIters; it0 += 8) {
            uint4 oldv[8];
            if (p.accumulate) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int idx = et + (it0 + j) * 128, r = idx / kChunks, ch = idx - r * kChunks;
                    const int gi = sh->row_index[r];
                    oldv[j] = make_uint4(0, 0, 0, 0);
                    if (gi >= 0 && col0 + ch * 8 < p.N) oldv[j] = *reinterpret_cast<const uint4*>(out + (size_t)gi * p.ldc + col0 + ch * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int idx = et + (it0 + j) * 128, r = idx / kChunks, ch = idx - r * kChunks;
                const int gi = sh->row_index[r];
                if (gi < 0 || col0 + ch * 8 >= p.N) continue;
                uint4 val = *reinterpret_cast<const uint4*>(staging + r * Cfg::kPitch + ch * 16);
                if (p.accumulate) {
                    const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&oldv[j]);
                    __nv_bfloat162* v2 = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 a = __bfloat1622float2(v2[q]), b = __bfloat1622float2(o2[q]);
                        v2[q] = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
                    }
                }
                *reinterpret_cast<uint4*>(out + (size_t)gi * p.ldc + col0 + ch * 8) = val;
            }
        }
        if (dbg && et == 0) dbg[6] = (long long)gtimer();      // end of this thread's share of the epilogue stores
    }
    // ---- teardown: neither CTA may free TMEM or exit while its partner can still touch it --------------------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2cta(tmem_acc, BN);
}

}  // namespace

// tmB must have been encoded with a {64, BN/2} box (K-major B) or the {64, 64} box of the forward filter view (b_mn).  Grid: M tiles rounded up to whole pairs x N tiles.
template <int BN, int kDeep>
static cudaError_t launch_2cta_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p_in, int m_tiles, cudaStream_t st) {
    using Cfg = CCfg<BN, kDeep>;
    ConvGemmParams p = p_in;
    p.dbg = conv_trace_buf();
    static bool configured = false;
    if (!configured) {
        RLR_CUDA_CHECK(cudaFuncSetAttribute(umma_conv_gemm_2cta_kernel<BN, false, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        RLR_CUDA_CHECK(cudaFuncSetAttribute(umma_conv_gemm_2cta_kernel<BN, true, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        configured = true;
    }
    const dim3 grid((m_tiles + 1) / 2 * 2, (p.N + BN - 1) / BN);
    if (p.b_mn) return launch_kernel(umma_conv_gemm_2cta_kernel<BN, true, kDeep>, grid, dim3(CThreads), (size_t)Cfg::kSmemBytes, st, tmA, tmB, p);
    return launch_kernel(umma_conv_gemm_2cta_kernel<BN, false, kDeep>, grid, dim3(CThreads), (size_t)Cfg::kSmemBytes, st, tmA, tmB, p);
}

// tmB must have been encoded with a {64, BN/2} box (K-major B) or the {64, 64} box of the forward filter view (b_mn).  Grid: M tiles rounded
// up to whole pairs x N tiles.  deep = 1: one CTA per SM with the whole shared memory as operand ring (grids that do not exceed the SM count).
template <int BN>
cudaError_t launch_2cta_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, int m_tiles, cudaStream_t st, int deep) {
    return deep ? launch_2cta_cfg<BN, 1>(tmA, tmB, p, m_tiles, st) : launch_2cta_cfg<BN, 0>(tmA, tmB, p, m_tiles, st);
}
template cudaError_t launch_2cta_bn<128>(const CUtensorMap&, const CUtensorMap&, const ConvGemmParams&, int, cudaStream_t, int);
template cudaError_t launch_2cta_bn<256>(const CUtensorMap&, const CUtensorMap&, const ConvGemmParams&, int, cudaStream_t, int);

}  // namespace rlr
