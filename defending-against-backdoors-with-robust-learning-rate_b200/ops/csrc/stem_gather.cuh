// Stem GEMM (tiny-K first layer): the B operand tile is built by ONE warp from the un-padded bf16 filter b_src[N][b_ld] (b_kvalid <= 64 valid
// columns) straight into the 128-byte-swizzled K-major shared-memory layout the tensor core reads -- no padded filter copy in global memory.
// With wait_flags the warp first acquires the broadcast-ready words of the slices that hold the filter (round hand-off fused with the first
// GEMM: the filter may still be in flight from the other GPUs' aggregation kernels).  Used by gemm.cu and gemm_persistent.cu.
#pragma once
#include "common.cuh"
#include "gemm.h"

namespace rlr {

__device__ __forceinline__ void stem_gather_b(const ConvGemmParams& p, uint8_t* sb0, int BN, int tile_n, int lane) {
    if (p.wait_flags) {
        const uint32_t epoch = *p.wait_epoch;
        for (int r = p.wait_lo + lane; r <= p.wait_hi; r += 32)
            while ((int32_t)(ld_acquire_sys(p.wait_flags + r) - epoch) < 0) { __nanosleep(32); }
        __syncwarp();
    }
    // zero the tile (K padding and rows beyond N), then scatter the valid elements: consecutive lanes read consecutive filter elements
    // (coalesced, 8 independent loads in flight per lane) and store them at their swizzled position: 16-byte chunk c of row r lives at
    // chunk (c ^ (r & 7)) of that row
    for (int i = lane; i < BN * 8; i += 32) *reinterpret_cast<uint4*>(sb0 + i * 16) = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const int rows = min(BN, p.N - tile_n * BN), total = rows * p.b_kvalid;
    const __nv_bfloat16* src = p.b_src + (size_t)tile_n * BN * p.b_ld;
    for (int e0 = 0; e0 < total; e0 += 32 * 8) {
        unsigned short v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + j * 32 + lane;
            const int r = e / p.b_kvalid, col = e - r * p.b_kvalid;
            v[j] = e < total ? __bfloat16_as_ushort(src[(size_t)r * p.b_ld + col]) : (unsigned short)0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + j * 32 + lane;
            if (e < total) {
                const int r = e / p.b_kvalid, col = e - r * p.b_kvalid;
                *reinterpret_cast<unsigned short*>(sb0 + r * 128 + (((col >> 3) ^ (r & 7)) << 4) + (col & 7) * 2) = v[j];
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes above -> tensor-core (async proxy) reads
    __syncwarp();
}

}  // namespace rlr
