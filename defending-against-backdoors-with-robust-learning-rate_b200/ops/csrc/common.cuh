// Shared device helpers for the sm_100a kernels of b200-robust-fl.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "dropspec.h"

#define RLR_CUDA_CHECK(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) return _e;                                                        \
    } while (0)

namespace rlr {

constexpr int kWarp = 32;

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------------
// Kernels launched through launch_kernel() with PDL enabled (RLR_PDL=1 / set_pdl) may start while their predecessor in the stream
// is still draining: they run their prologue (barrier init, TMEM allocation, tensor-map prefetch, index arithmetic) and then block
// in pdl_wait() until the predecessor grid has completed and its memory is visible.  Rules every such kernel follows: NO global
// memory access before pdl_wait(); pdl_wait() is executed by every thread; pdl_trigger() (lets the successor start launching)
// comes after it.  Without the launch attribute both instructions are no-ops, so the same kernels serve plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_pdl;             // -1: read RLR_PDL on first use (default off); defined in elementwise.cu
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    if (!pdl_enabled()) {
        kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ unsigned long long gtimer() {     // nanosecond timer common to all SMs
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Column sums of a 32(lanes) x 32(values) tile held one row per lane: after the call lane l returns sum_over_lanes v[l].
// Butterfly that halves the live values per stage: 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#define RLR_TSTAGE(O, N)                                                         \
    {                                                                            \
        const bool up = (lane & (O)) != 0;                                       \
        _Pragma("unroll") for (int i = 0; i < (N); ++i) {                        \
            const float send = up ? v[i] : v[i + (N)];                           \
            const float keep = up ? v[i + (N)] : v[i];                           \
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, (O));               \
        }                                                                        \
    }
    RLR_TSTAGE(16, 16) RLR_TSTAGE(8, 8) RLR_TSTAGE(4, 4) RLR_TSTAGE(2, 2) RLR_TSTAGE(1, 1)
#undef RLR_TSTAGE
    return v[0];
}

// Block-wide sum; result valid in thread 0. `scratch` must hold >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
    if (wid == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

// ---- streaming 128-bit global accesses (read-once / write-once data) --------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// Peer (NVLink) / freshly written data: plain relaxed load, never the non-coherent path.
__device__ __forceinline__ float4 ld_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_f1(const float* p) {
    float r;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_f4(float* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// NVLS multicast store: one store lands in every GPU bound to the multicast object.
__device__ __forceinline__ void multimem_st_f4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_b2(uint2* mc, uint2 v) {  // 4 packed bf16
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)) : "memory");
}

// ---- cross-GPU flags (system scope) ------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- Philox4x32-10 counter RNG (Salmon et al.), used for dropout masks and server noise ---------------
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint64_t stream) const {
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
__device__ __forceinline__ float u32_to_unit(uint32_t x) {  // (0,1]
    return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float4 philox_normal4(const Philox& ph, uint64_t ctr, uint64_t stream) {
    const uint4 u = ph(ctr, stream);
    const float r0 = sqrtf(-2.0f * logf(u32_to_unit(u.x))), r1 = sqrtf(-2.0f * logf(u32_to_unit(u.z)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u32_to_unit(u.y), &s0, &c0);
    sincospif(2.0f * u32_to_unit(u.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// ---- fused dropout --------------------------------------------------------------------------------------------------------------
// Keep-mask of a tensor viewed as a flat array of elements: the 8 elements [8 q, 8 q + 8) share ONE Philox4x32-10 call keyed by
// (seed; counter q, stream = (step << 20) ^ node), 16 random bits per element, keep iff bits >= p * 65536.  Every kernel that produces
// or back-propagates through a dropped tensor (maxpool / GEMM epilogues, their backward kernels, the stand-alone dropout kernels)
// evaluates THIS function, so no mask tensor is ever written.  Reference: nn.Dropout2d(p=.5) on 2-D activations (src/models.py:17-19,
// 40-44) = element-wise dropout; bit-parity with torch's Philox stream is not a goal, mask statistics are tested.
__device__ __forceinline__ uint32_t dropout_keep8(const DropSpec& d, long long q) {     // bit i = element 8 q + i is kept
    const Philox ph(d.seed);
    const uint4 u = ph((uint64_t)q, ((uint64_t)(*d.step) << 20) ^ d.stream);
    const uint32_t r[4] = {u.x, u.y, u.z, u.w};
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) keep |= (uint32_t)(((r[i >> 1] >> (16 * (i & 1))) & 0xffffu) >= d.thr) << i;
    return keep;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace rlr
