// Fused federated server step for sm_100a: one pass over the flat parameter vector that
//   * reads every participating agent's local parameters w_k (local HBM, or a PEER GPU's HBM through
//     NVLink/NVSwitch-mapped pointers) and forms the update d_k = w_k - w_g in registers (never materialised),
//   * takes the per-coordinate sign vote  s = |sum_k sign(d_k)|  ->  lr = +server_lr if s >= theta else -server_lr
//     (Robust Learning Rate; reference src/aggregation.py:48-54),
//   * aggregates: data-size-weighted mean (:57-64) | lower coordinate median (:66-69) | sign majority (:71-75),
//   * adds optional Gaussian noise (in-kernel Philox; :34-35) BEFORE the lr multiply, like the reference,
//   * applies the server step  w_g' = w_g + lr * agg  (:38-40) in fp64 then rounds to fp32,
//   * and writes w_g' (+ its bf16 GEMM-operand shadow) to every GPU: one NVLS `multimem.st` per 16 bytes when a
//     multicast mapping exists, else one P2P store per peer.  This store IS the next round's broadcast
//     (reference src/federated.py:72).
// Across GPUs rank r owns coordinates [begin,end); the kernel is reduce-scatter ∘ compute ∘ all-gather in one
// launch, bracketed by release/acquire flag barriers on peer-mapped signal words (no NCCL, no host sync).
// Coordinates >= n_vote (BatchNorm running statistics) get a plain weighted mean, no vote.
#include "common.cuh"
#include "kernels.h"

namespace rlr {

constexpr int kAggThreads = 256;
constexpr int kMaxAgents = 128;

template <int K>
__device__ __forceinline__ float lower_median_fixed(float (&v)[K]) {
    // odd-even transposition network, fully unrolled -> registers only
#pragma unroll
    for (int pass = 0; pass < K; ++pass) {
#pragma unroll
        for (int j = pass & 1; j + 1 < K; j += 2) {
            const float lo = fminf(v[j], v[j + 1]), hi = fmaxf(v[j], v[j + 1]);
            v[j] = lo; v[j + 1] = hi;
        }
    }
    return v[(K - 1) / 2];  // torch.median returns the LOWER median for even K
}

__device__ __forceinline__ float lower_median_dyn(float* v, int K) {
    for (int i = 1; i < K; ++i) {  // insertion sort in local memory (large participant counts only)
        const float x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
    return v[(K - 1) / 2];
}

__device__ __forceinline__ int sgn(float d) { return (d > 0.f) - (d < 0.f); }

// Cross-GPU barrier executed by the first `world` threads of ONE block.
__device__ __forceinline__ void xgpu_barrier(uint32_t* const* flag_ptrs, int slot_base, int rank, int world, uint32_t epoch) {
    const int t = threadIdx.x;
    if (t < world) {
        st_release_sys(flag_ptrs[t] + slot_base + rank, epoch);           // tell peer t "rank is here"
        const uint32_t* mine = flag_ptrs[rank] + slot_base + t;           // wait for peer t
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) { __nanosleep(64); }
    }
}

// MODE: 0 avg, 1 comed, 2 sign.  KT: compile-time K for the median network (0 = runtime K).
template <int MODE, int KT>
__global__ void __launch_bounds__(kAggThreads) fused_aggregate_kernel(AggParams p) {
    __shared__ const float* s_w[kMaxAgents];
    __shared__ double s_wt[kMaxAgents];
    __shared__ float s_sc[kMaxAgents];
    __shared__ unsigned long long s_scratch[32];
    const int K = KT > 0 ? KT : p.K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        s_w[k] = p.w_agents[k];
        s_wt[k] = p.weights[k];
        s_sc[k] = p.scales ? p.scales[k] : 1.0f;
    }
    // ---- barrier-in: every peer's local training has finished and its w_k is globally visible -------------
    if (p.world > 1) {
        if (blockIdx.x == 0) {
            xgpu_barrier(p.flag_ptrs, 0, p.rank, p.world, p.epoch);
            __syncthreads();
            if (threadIdx.x == 0) st_release_gpu(p.local_sync, p.epoch);
        } else if (threadIdx.x == 0) {
            while ((int32_t)(ld_acquire_gpu(p.local_sync) - p.epoch) < 0) { __nanosleep(32); }
        }
    }
    __syncthreads();

    const Philox ph(p.seed);
    const double inv_total = 1.0 / p.total_weight;
    unsigned long long flipped = 0;
    const long long n4 = (p.end - p.begin) >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const long long i = p.begin + (q << 2);
        const bool tail = i >= p.n_vote;
        const float4 g4 = ld_f4(p.w_global + i);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        int s[4] = {0, 0, 0, 0};
        double acc[4] = {0., 0., 0., 0.};
        float agg[4];
        if (MODE == 1 && !tail) {
            if constexpr (KT > 0) {
                float v0[KT], v1[KT], v2[KT], v3[KT];
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    const float4 w = ld_f4(s_w[k] + i);
                    const float sc = s_sc[k];
                    v0[k] = (w.x - g[0]) * sc; v1[k] = (w.y - g[1]) * sc; v2[k] = (w.z - g[2]) * sc; v3[k] = (w.w - g[3]) * sc;
                    s[0] += sgn(v0[k]); s[1] += sgn(v1[k]); s[2] += sgn(v2[k]); s[3] += sgn(v3[k]);
                }
                agg[0] = lower_median_fixed<KT>(v0); agg[1] = lower_median_fixed<KT>(v1);
                agg[2] = lower_median_fixed<KT>(v2); agg[3] = lower_median_fixed<KT>(v3);
            } else {
                float v0[kMaxAgents], v1[kMaxAgents], v2[kMaxAgents], v3[kMaxAgents];
                for (int k = 0; k < K; ++k) {
                    const float4 w = ld_f4(s_w[k] + i);
                    const float sc = s_sc[k];
                    v0[k] = (w.x - g[0]) * sc; v1[k] = (w.y - g[1]) * sc; v2[k] = (w.z - g[2]) * sc; v3[k] = (w.w - g[3]) * sc;
                    s[0] += sgn(v0[k]); s[1] += sgn(v1[k]); s[2] += sgn(v2[k]); s[3] += sgn(v3[k]);
                }
                agg[0] = lower_median_dyn(v0, K); agg[1] = lower_median_dyn(v1, K);
                agg[2] = lower_median_dyn(v2, K); agg[3] = lower_median_dyn(v3, K);
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float4 w = ld_f4(s_w[k] + i);
                const float sc = s_sc[k];
                const float d[4] = {(w.x - g[0]) * sc, (w.y - g[1]) * sc, (w.z - g[2]) * sc, (w.w - g[3]) * sc};
                const double wt = s_wt[k];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    s[c] += sgn(d[c]);
                    if (MODE == 0 || tail) acc[c] += wt * (double)d[c];
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (MODE == 0 || tail) agg[c] = (float)(acc[c] * inv_total);
                else agg[c] = (float)((s[c] > 0) - (s[c] < 0));  // sign majority
            }
        }
        float out[4];
        if (tail) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out[c] = (float)((double)g[c] + acc[c] * inv_total);
        } else {
            float nz[4] = {0.f, 0.f, 0.f, 0.f};
            if (p.noise_std > 0.f) {
                const float4 z = philox_normal4(ph, (uint64_t)(i >> 2), p.noise_stream);
                nz[0] = z.x * p.noise_std; nz[1] = z.y * p.noise_std; nz[2] = z.z * p.noise_std; nz[3] = z.w * p.noise_std;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                // avg keeps its fp64 mean; comed/sign values are exact in fp32
                const double a = ((MODE == 0) ? acc[c] * inv_total : (double)agg[c]) + (double)nz[c];
                const bool keep = (p.theta <= 0) || (abs(s[c]) >= p.theta);
                flipped += keep ? 0 : 1;
                const double lr = keep ? (double)p.server_lr : -(double)p.server_lr;
                out[c] = (float)((double)g[c] + lr * a);
            }
        }
        const float4 o4 = make_float4(out[0], out[1], out[2], out[3]);
        const uint2 b4 = make_uint2(pack_bf16x2(out[0], out[1]), pack_bf16x2(out[2], out[3]));
        if (p.use_multimem) {
            multimem_st_f4(p.out_ptrs[0] + i, o4);
            if (p.out_bf16_ptrs) multimem_st_b2(reinterpret_cast<uint2*>(p.out_bf16_ptrs[0] + i), b4);
        } else {
            for (int d = 0; d < p.n_out; ++d) {
                st_f4(p.out_ptrs[d] + i, o4);
                if (p.out_bf16_ptrs) *reinterpret_cast<uint2*>(p.out_bf16_ptrs[d] + i) = b4;
            }
        }
    }
    // ---- statistics: number of coordinates whose learning rate was flipped ---------------------------------
    if (p.flipped) {
        const unsigned long long tot = block_sum<unsigned long long>(flipped, s_scratch);
        if (threadIdx.x == 0 && tot) atomicAdd(p.flipped, tot);
    }
    // ---- barrier-out: my slice has landed everywhere; wait until every peer's slice has landed here ---------
    if (p.world > 1) {
        __threadfence_system();
        __syncthreads();
        __shared__ int s_last;
        if (threadIdx.x == 0) {
            const unsigned prev = atomicAdd(p.local_sync + 1, 1u);
            s_last = (prev == gridDim.x - 1);
            if (s_last) { p.local_sync[1] = 0; }
            __threadfence();
        }
        __syncthreads();
        if (s_last) xgpu_barrier(p.flag_ptrs, p.world, p.rank, p.world, p.epoch);
    }
}

template <int MODE, int KT>
static cudaError_t launch_one(const AggParams& p, int grid, cudaStream_t st) {
    fused_aggregate_kernel<MODE, KT><<<grid, kAggThreads, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fused_aggregate(const AggParams& p, int num_sms, cudaStream_t st) {
    if (p.K < 1 || p.K > kMaxAgents) return cudaErrorInvalidValue;
    if (((p.end - p.begin) & 3) || (p.begin & 3) || (p.n_vote & 3)) return cudaErrorInvalidValue;
    const long long n4 = (p.end - p.begin) >> 2;
    long long want = (n4 + kAggThreads - 1) / kAggThreads;
    // all CTAs co-resident (<= 8 per SM) so the intra-kernel flag barriers can never starve
    int grid = (int)(want < 1 ? 1 : (want > (long long)num_sms * 8 ? (long long)num_sms * 8 : want));
    switch (p.mode) {
        case 0: return launch_one<0, 0>(p, grid, st);
        case 2: return launch_one<2, 0>(p, grid, st);
        case 1:
            switch (p.K) {
                case 1: return launch_one<1, 1>(p, grid, st);
                case 2: return launch_one<1, 2>(p, grid, st);
                case 3: return launch_one<1, 3>(p, grid, st);
                case 4: return launch_one<1, 4>(p, grid, st);
                case 5: return launch_one<1, 5>(p, grid, st);
                case 6: return launch_one<1, 6>(p, grid, st);
                case 7: return launch_one<1, 7>(p, grid, st);
                case 8: return launch_one<1, 8>(p, grid, st);
                default: return launch_one<1, 0>(p, grid, st);
            }
        default: return cudaErrorInvalidValue;
    }
}

// ---- per-agent update L2 norms  ||w_k - w_g||_2^2  (server clipping + the Norms/* diagnostics) -------------
__global__ void __launch_bounds__(256) update_sqnorm_kernel(const float* const* w_agents, const float* w_global,
                                                              long long n, double* out /*[K]*/) {
    __shared__ double scratch[32];
    const float* w = w_agents[blockIdx.y];
    double acc = 0.0;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 a = ld_f4(w + i), g = ld_f4(w_global + i);
        const float d0 = a.x - g.x, d1 = a.y - g.y, d2 = a.z - g.z, d3 = a.w - g.w;
        acc += (double)(d0 * d0 + d1 * d1) + (double)(d2 * d2 + d3 * d3);
    }
    const double tot = block_sum<double>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, tot);
}

cudaError_t launch_update_sqnorm(const float* const* w_agents, const float* w_global, long long n, int K, double* out,
                                 int num_sms, cudaStream_t st) {
    if (n & 3) return cudaErrorInvalidValue;
    RLR_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * K, st));
    long long want = (n / 4 + 255) / 256;
    int gx = (int)(want > num_sms * 4 ? num_sms * 4 : (want < 1 ? 1 : want));
    update_sqnorm_kernel<<<dim3(gx, K), 256, 0, st>>>(w_agents, w_global, n, out);
    return cudaGetLastError();
}

}  // namespace rlr
