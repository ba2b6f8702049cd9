// Persistent variant of the implicit-GEMM conv / GEMM kernel of gemm.cu (opt-in: RLR_PERSISTENT_CONV=1).
//
// Same operands, tap tables and epilogue semantics as umma_conv_gemm_kernel, but ONE CTA per SM loops over output tiles
// (tile = blockIdx.x, += gridDim.x; consecutive tiles share the filter n-tile), with a deeper operand ring (7 x 24 KB for
// BN = 64, 5 x 32 KB for BN = 128), a dedicated staging tile and DOUBLE-BUFFERED TMEM accumulators (2 x BN columns): the
// producer and the MMA thread run straight through tile boundaries while the four epilogue warps drain tile i during the
// main loop of tile i+1 -- the structure validated in conv_halo.cu.  No statistics path (BatchNorm statistics are taken by
// the streaming pass, see models/native.py).
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"
#include "stem_gather.cuh"
#include "umma.cuh"

namespace rlr {

using namespace umma;

constexpr int PBM = 128, PBK = 64, PThreads = 192;

// kOcc = 1: one CTA per SM with a deep ring (7 x 24 KB / 5 x 32 KB).  kOcc = 2: two CTAs per SM with a 3-stage ring -- for the stem GEMM,
// whose one-k-block tiles are EPILOGUE bound: two CTAs give the SM two epilogue warp-groups (and two TMA streams) to drain them.
template <int BN, int kOcc = 1>
struct PCfg {
    static constexpr int kStages = kOcc == 2 ? 3 : (BN == 64 ? 7 : 5);
    static constexpr int kABytes = PBM * PBK * 2;
    static constexpr int kBBytes = BN * PBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kPitch = BN * 2 + 16;
    static constexpr int kStagingBytes = PBM * kPitch;
    static constexpr int kRingBytes = kStages * kStageBytes;
    static constexpr int kTailBytes = BN == 128 ? 2048 : 1024;    // PShared (704 B) + the n-tile's bias values (BN floats)
    static constexpr int kSmemBytes = kRingBytes + kStagingBytes + 1024 + kTailBytes;
};

struct __align__(8) PShared {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    int row_index[PBM];
};
static_assert(sizeof(PShared) <= 704, "the bias slice starts 704 bytes into the tail");

template <int BN, bool kBMN, int kOcc = 1>
__global__ void __launch_bounds__(PThreads, kOcc)
umma_conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGemmParams p,
                                 const int m_tiles, const int total_tiles) {
    using Cfg = PCfg<BN, kOcc>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + Cfg::kRingBytes;
    PShared* sh = reinterpret_cast<PShared*>(staging + Cfg::kStagingBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&sh->acc_full[a], 1); mbar_init(&sh->acc_empty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&sh->tmem_base, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = sh->tmem_base;

    // tile -> (tile_m, tile_n) and image-space origin
    auto origin = [&](int tile, int& tile_m, int& tile_n, int& n0, int& h0, int& w0) {
        tile_m = tile % m_tiles; tile_n = tile / m_tiles;
        n0 = h0 = w0 = 0;
        if (p.mode == 1) {
            const int tw_i = tile_m % p.tiles_w, th_i = (tile_m / p.tiles_w) % p.tiles_h, tn_i = tile_m / (p.tiles_w * p.tiles_h);
            w0 = tw_i * p.TW; h0 = th_i * p.TH; n0 = tn_i * p.TN;
        }
    };

    if (warp == 0) {
        // stem GEMM (b_src: one k-block, one n-tile): the B tile is built ONCE per CTA by this warp in the B slot of stage 0 (no stage ever
        // loads a B tile in this mode, so the slot is never overwritten) and every tile's MMAs read it there; per tile only A is fetched
        if (p.b_src) stem_gather_b(p, smem + Cfg::kABytes, BN, 0, lane);
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int tile_m, tile_n, n0, h0, w0;
                origin(tile, tile_m, tile_n, n0, h0, w0);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&sh->empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    uint8_t* sb = sa + Cfg::kABytes;
                    mbar_expect_tx(&sh->full[stage], p.b_src ? Cfg::kABytes : Cfg::kStageBytes);
                    const int tap = p.mode == 1 ? kb / p.cblocks : 0, cb = p.mode == 1 ? kb - tap * p.cblocks : 0;
                    if (p.mode == 1) tma_load_4d(&tmA, &sh->full[stage], sa, cb * PBK, w0 * p.in_stride + p.dw[tap], h0 * p.in_stride + p.dh[tap],
                                                n0 + p.dn[tap]);
                    else tma_load_2d(&tmA, &sh->full[stage], sa, kb * PBK, tile_m * PBM);
                    if (p.b_src) {
                        // resident B tile (above)
                    } else if (kBMN) {
                        for (int g = 0; g < BN / 64; ++g)
                            tma_load_2d(&tmB, &sh->full[stage], sb + g * 8192, p.wtap[tap] * p.wcols + tile_n * BN + g * 64, cb * PBK);
                    } else {
                        tma_load_2d(&tmB, &sh->full[stage], sb, kb * PBK, tile_n * BN);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(PBM, BN, 0, kBMN ? 1 : 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&sh->acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_acc + acc * BN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&sh->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sb = p.b_src ? smem_u32(smem + Cfg::kABytes) : sa + Cfg::kABytes;
#pragma unroll
                    for (int k = 0; k < PBK / 16; ++k) {
                        const uint64_t da = smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = kBMN ? smem_desc_sw128(sb + k * 2048, 8192, 1024) : smem_desc_sw128(sb + k * 32, 16, 1024);
                        umma_bf16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&sh->empty[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&sh->acc_full[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        const int et = threadIdx.x - 64;
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        constexpr int kChunks = BN * 2 / 16;
        constexpr int kIters = PBM * kChunks / 128;
        int acc = 0;
        uint32_t acc_phase = 0;
        float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + 704);
        int bias_tile_n = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int tile_m, tile_n, n0, h0, w0;
            origin(tile, tile_m, tile_n, n0, h0, w0);
            const int col0 = tile_n * BN;
            if (p.bias && tile_n != bias_tile_n) {   // (re)load this n-tile's bias values; consecutive tiles share the n-tile
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = et; i < BN; i += 128) bias_s[i] = (col0 + i < p.N) ? p.bias[col0 + i] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_tile_n = tile_n;
            }
            {   // global output row of my tile row (-1 = masked); staging / row_index of the previous tile are free (end-of-tile barrier)
                int gi;
                if (p.mode == 1) {
                    const int tw = row % p.TW, th = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
                    const int w = w0 + tw, h = h0 + th, n = n0 + tn;
                    gi = (w < p.Wo && h < p.Ho && n < p.NB)
                             ? ((n * p.OutH + h * p.out_stride + p.out_ph) * p.OutW + w * p.out_stride + p.out_pw) : -1;
                } else {
                    gi = tile_m * PBM + row;
                    if (gi >= p.M) gi = -1;
                }
                sh->row_index[row] = gi;
            }
            mbar_wait(&sh->acc_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + ((uint32_t)lane_base << 16) + acc * BN + c0, v);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
                    if (p.bias) { a += bias_s[c0 + j]; b += bias_s[c0 + j + 1]; }
                    if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                    packed[j >> 1] = pack_bf16x2(a, b);
                }
                uint4* dst = reinterpret_cast<uint4*>(staging + row * Cfg::kPitch + c0 * 2);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            tc_fence_before();
            mbar_arrive(&sh->acc_empty[acc]);                  // accumulator drained: the MMA thread may start tile i+2 in it
            asm volatile("bar.sync 1, 128;" ::: "memory");    // staging tile + row_index complete
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
            asm volatile("bar.sync 1, 128;" ::: "memory");    // staging / row_index may be overwritten by the next tile
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_acc, 2 * BN);
}

template <int BN>
cudaError_t launch_persistent_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, int m_tiles, int num_sms,
                                 cudaStream_t st, int occ) {
    const int n_tiles = (p.N + BN - 1) / BN;
    const int total = m_tiles * n_tiles;
    if (occ == 2 && !p.b_mn) {
        using Cfg = PCfg<BN, 2>;
        static bool configured2 = false;
        if (!configured2) {
            RLR_CUDA_CHECK(cudaFuncSetAttribute(umma_conv_gemm_persistent_kernel<BN, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
            configured2 = true;
        }
        const int grid = total < 2 * num_sms ? total : 2 * num_sms;
        umma_conv_gemm_persistent_kernel<BN, false, 2><<<grid, PThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p, m_tiles, total);
        return cudaGetLastError();
    }
    using Cfg = PCfg<BN>;
    static bool configured = false;
    if (!configured) {
        RLR_CUDA_CHECK(cudaFuncSetAttribute(umma_conv_gemm_persistent_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        RLR_CUDA_CHECK(cudaFuncSetAttribute(umma_conv_gemm_persistent_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        configured = true;
    }
    const int grid = total < num_sms ? total : num_sms;
    if (p.b_mn) umma_conv_gemm_persistent_kernel<BN, true><<<grid, PThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p, m_tiles, total);
    else umma_conv_gemm_persistent_kernel<BN, false><<<grid, PThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p, m_tiles, total);
    return cudaGetLastError();
}
template cudaError_t launch_persistent_bn<64>(const CUtensorMap&, const CUtensorMap&, const ConvGemmParams&, int, int, cudaStream_t, int);
template cudaError_t launch_persistent_bn<128>(const CUtensorMap&, const CUtensorMap&, const ConvGemmParams&, int, int, cudaStream_t, int);

}  // namespace rlr
