// Memory-bound kernels of the local-training loop: batch gather+normalise from the device-resident dataset,
// trojan stamping, fused clip-grad-norm + momentum-SGD (+ PGD projection) over the FLAT parameter buffers, and the
// loss / evaluation reductions.  Reference call sites: src/agent.py:41-60 (step), src/utils.py:52-54 (per-sample
// host transforms), :160-178 (poisoning), :128-157 (evaluation).
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace rlr {

// ---- programmatic dependent launch switch (common.cuh) -------------------------------------------------------------------------
int g_pdl = -1;
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
bool pdl_enabled() {
    if (g_pdl < 0) { const char* e = getenv("RLR_PDL"); g_pdl = (e && atoi(e) > 0) ? 1 : 0; }
    return g_pdl > 0;
}

// ------------------------------------------------------------------------------------------------------------
// batch = normalize(dataset[perm[cursor : cursor+B]])   (uint8/float NHWC  ->  fp32/bf16, NCHW or padded NHWC)
// ------------------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void gather_normalize_kernel(const TIn* __restrict__ data, const int64_t* __restrict__ idx,
                                        const int* __restrict__ cursor, const int64_t* __restrict__ targets,
                                        TOut* __restrict__ out, int64_t* __restrict__ out_labels, int B, int HW, int C,
                                        int c_pad, int nchw, float4 mean, float4 inv_std, float in_scale) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * HW) return;
    const int b = t / HW, px = t - b * HW;
    const int64_t src = idx[(cursor ? *cursor : 0) + b];
    if (px == 0 && out_labels) out_labels[b] = targets[src];
    const TIn* in = data + (src * HW + px) * C;
    const float mu[4] = {mean.x, mean.y, mean.z, mean.w};
    const float is[4] = {inv_std.x, inv_std.y, inv_std.z, inv_std.w};
    if (nchw) {
        for (int c = 0; c < C; ++c)
            out[((int64_t)b * C + c) * HW + px] = (TOut)(((float)in[c] * in_scale - mu[c]) * is[c]);
    } else {
        TOut* o = out + ((int64_t)b * HW + px) * c_pad;
        for (int c = 0; c < C; ++c) o[c] = (TOut)(((float)in[c] * in_scale - mu[c]) * is[c]);
        for (int c = C; c < c_pad; ++c) o[c] = (TOut)0.f;
    }
}

cudaError_t launch_gather_normalize(const void* data, int in_is_float, const int64_t* idx, const int* cursor,
                                    const int64_t* targets, void* out, int out_kind, int64_t* out_labels, int B, int H,
                                    int W, int C, int c_pad, int nchw, const float* mean, const float* stdv,
                                    cudaStream_t st) {
    if (C > 4 || B <= 0) return cudaErrorInvalidValue;
    float mu[4] = {0, 0, 0, 0}, is[4] = {1, 1, 1, 1};
    for (int c = 0; c < C; ++c) { mu[c] = mean[c]; is[c] = 1.0f / stdv[c]; }
    const float4 m4 = make_float4(mu[0], mu[1], mu[2], mu[3]), s4 = make_float4(is[0], is[1], is[2], is[3]);
    const int HW = H * W, total = B * HW, threads = 256, blocks = (total + threads - 1) / threads;
    const float sc = in_is_float ? 1.0f : (1.0f / 255.0f);
#define RLR_GN(TI, TO)                                                                                          \
    gather_normalize_kernel<TI, TO><<<blocks, threads, 0, st>>>((const TI*)data, idx, cursor, targets, (TO*)out, \
                                                                 out_labels, B, HW, C, c_pad, nchw, m4, s4, sc)
    if (in_is_float) { if (out_kind == 0) RLR_GN(float, float); else RLR_GN(float, __nv_bfloat16); }
    else             { if (out_kind == 0) RLR_GN(uint8_t, float); else RLR_GN(uint8_t, __nv_bfloat16); }
#undef RLR_GN
    return cudaGetLastError();
}

// end of a local step: the batch cursor moves on and (optionally) the Philox step counter of the dropout masks is bumped --
// one single-thread node instead of a torch `+= 1` inside the captured step
__global__ void advance_cursor_kernel(int* cursor, int delta, long long* step) {
    *cursor += delta;
    if (step) *step += 1;
}
cudaError_t launch_advance_cursor(int* cursor, int delta, long long* step, cudaStream_t st) {
    advance_cursor_kernel<<<1, 1, 0, st>>>(cursor, delta, step);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// row padding for tiny-K operands: dst[r][0..Kp) = src[r][0..K) followed by zeros (bf16).  Used for the stem conv's channel-
// padded input / filter and the im2col filter matrix ([Cout][k*k*Cin] -> [Cout][64]); the inverse adds the valid columns of an
// fp32 [R][Kp] gradient into the [R][K] slice of the flat gradient.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         long long R, int K, int Kp) {
    const long long n = R * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / Kp;
        const int c = (int)(i - r * Kp);
        dst[i] = c < K ? src[r * K + c] : __float2bfloat16(0.f);
    }
}
cudaError_t launch_pad_rows(const __nv_bfloat16* src, __nv_bfloat16* dst, long long R, int K, int Kp, int num_sms, cudaStream_t st) {
    if (K > Kp || R <= 0) return cudaErrorInvalidValue;
    const long long want = (R * Kp + 255) / 256;
    const int grid = (int)(want > (long long)num_sms * 8 ? (long long)num_sms * 8 : want);
    pad_rows_kernel<<<grid, 256, 0, st>>>(src, dst, R, K, Kp);
    return cudaGetLastError();
}
__global__ void __launch_bounds__(256) unpad_add_kernel(const float* __restrict__ src, float* __restrict__ dst, long long R, int K, int Kp) {
    const long long n = R * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K;
        dst[i] += src[r * Kp + (i - r * K)];
    }
}
cudaError_t launch_unpad_add(const float* src, float* dst, long long R, int K, int Kp, int num_sms, cudaStream_t st) {
    if (K > Kp || R <= 0) return cudaErrorInvalidValue;
    const long long want = (R * K + 255) / 256;
    const int grid = (int)(want > (long long)num_sms * 8 ? (long long)num_sms * 8 : want);
    unpad_add_kernel<<<grid, 256, 0, st>>>(src, dst, R, K, Kp);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// trojan stamping: apply a compiled pixel program to selected images in place
// mode 0: set (all channels) | 1: uint8 wrap-around add | 2: float subtract
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void stamp_pixels_kernel(T* __restrict__ data, const int64_t* __restrict__ sel, int S,
                                    const int* __restrict__ rows, const int* __restrict__ cols,
                                    const float* __restrict__ vals, int P, int H, int W, int C, int mode) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)S * P * C) return;
    const int c = (int)(t % C);
    const int pi = (int)((t / C) % P);
    const int64_t s = t / ((int64_t)C * P);
    T* px = data + ((sel[s] * H + rows[pi]) * W + cols[pi]) * C + c;
    const float v = vals[pi];
    if (mode == 0) *px = (T)v;
    else if (mode == 1) *px = (T)(uint8_t)((unsigned)(*px) + (unsigned)v);  // wraps mod 256 like numpy uint8
    else *px = (T)((float)(*px) - v);
}

cudaError_t launch_stamp_pixels(void* data, int is_float, const int64_t* sel, int S, const int* rows, const int* cols,
                                const float* vals, int P, int H, int W, int C, int mode, cudaStream_t st) {
    if (S <= 0 || P <= 0) return cudaSuccess;
    const int64_t total = (int64_t)S * P * C;
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads);
    if (is_float) stamp_pixels_kernel<float><<<blocks, threads, 0, st>>>((float*)data, sel, S, rows, cols, vals, P, H, W, C, mode);
    else stamp_pixels_kernel<uint8_t><<<blocks, threads, 0, st>>>((uint8_t*)data, sel, S, rows, cols, vals, P, H, W, C, mode);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// round start: w_local <- w_global (the "broadcast" consumer), bf16 operand shadow, momentum <- 0
// (reference: deepcopy + vector_to_parameters src/federated.py:72; fresh optimizer src/agent.py:37-38)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) round_init_kernel(const float* __restrict__ wg, float* __restrict__ wl,
                                                           __nv_bfloat16* __restrict__ wb, float* __restrict__ mom,
                                                           long long n4) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 v = ld_f4(wg + 4 * q);
        if (wl) st_f4(wl + 4 * q, v);
        if (wb) *reinterpret_cast<uint2*>(wb + 4 * q) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        if (mom) st_f4(mom + 4 * q, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}
static inline int grid_for(long long n4, int threads, int num_sms, int per_sm) {
    long long want = (n4 + threads - 1) / threads;
    long long cap = (long long)num_sms * per_sm;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}
cudaError_t launch_round_init(const float* w_global, float* w_local, __nv_bfloat16* w_bf16, float* mom, long long n,
                              cudaStream_t st) {
    if (n & 3) return cudaErrorInvalidValue;
    round_init_kernel<<<grid_for(n / 4, 256, 148, 8), 256, 0, st>>>(w_global, w_local, w_bf16, mom, n / 4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// fused clip_grad_norm_(.,max) + SGD(momentum) [+ ||w-w0||^2 for PGD] over flat buffers   (src/agent.py:50-60)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ x, long long n4, double* out) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 v = ld_f4(x + 4 * q);
        acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    const double tot = block_sum<double>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(out, tot);
}
cudaError_t launch_sqnorm(const float* x, long long n, double* out, int num_sms, cudaStream_t st) {
    if (n & 3) return cudaErrorInvalidValue;
    sqnorm_kernel<<<grid_for(n / 4, 256, num_sms, 4), 256, 0, st>>>(x, n / 4, out);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                         float* __restrict__ m, const float* __restrict__ w0,
                                                         __nv_bfloat16* __restrict__ wb, long long n4, float lr,
                                                         float momentum, float max_grad_norm,
                                                         const double* __restrict__ g_sqnorm, double* d_sqnorm, long long n4_pgd) {
    __shared__ double scratch[32];
    float coef = 1.0f;
    if (max_grad_norm > 0.f && g_sqnorm) {
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
        coef = fminf(1.0f, max_grad_norm / ((float)sqrt(*g_sqnorm) + 1e-6f));
    }
    double dacc = 0.0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const float4 gv = ld_f4(g + 4 * q), mv = ld_f4(m + 4 * q), wv = ld_f4(w + 4 * q);
        float4 mn, wn;
        mn.x = momentum * mv.x + coef * gv.x; mn.y = momentum * mv.y + coef * gv.y;
        mn.z = momentum * mv.z + coef * gv.z; mn.w = momentum * mv.w + coef * gv.w;
        wn.x = wv.x - lr * mn.x; wn.y = wv.y - lr * mn.y; wn.z = wv.z - lr * mn.z; wn.w = wv.w - lr * mn.w;
        st_f4(m + 4 * q, mn);
        st_f4(w + 4 * q, wn);
        if (d_sqnorm) {
            // PGD radius is measured over the model parameters only ([0, n_pgd): the reference projects parameters_to_vector(),
            // src/agent.py:54-60); BatchNorm running statistics stored behind them never count and are never rescaled
            if (q < n4_pgd) {
                const float4 o = ld_f4(w0 + 4 * q);
                const float d0 = wn.x - o.x, d1 = wn.y - o.y, d2 = wn.z - o.z, d3 = wn.w - o.w;
                dacc += (double)(d0 * d0 + d1 * d1) + (double)(d2 * d2 + d3 * d3);
            }
        } else if (wb) {
            *reinterpret_cast<uint2*>(wb + 4 * q) = make_uint2(pack_bf16x2(wn.x, wn.y), pack_bf16x2(wn.z, wn.w));
        }
    }
    if (d_sqnorm) {
        const double tot = block_sum<double>(dacc, scratch);
        if (threadIdx.x == 0) atomicAdd(d_sqnorm, tot);
    }
}
cudaError_t launch_sgd_step(float* w, const float* g, float* m, const float* w0, __nv_bfloat16* w_bf16, long long n,
                            float lr, float momentum, float max_grad_norm, const double* g_sqnorm, double* d_sqnorm,
                            int num_sms, cudaStream_t st, long long n_pgd) {
    if ((n & 3) || (n_pgd & 3)) return cudaErrorInvalidValue;
    if (n_pgd <= 0 || n_pgd > n) n_pgd = n;
    sgd_step_kernel<<<grid_for(n / 4, 256, num_sms, 4), 256, 0, st>>>(w, g, m, w0, w_bf16, n / 4, lr, momentum,
                                                                       max_grad_norm, g_sqnorm, d_sqnorm, n_pgd / 4);
    return cudaGetLastError();
}

// PGD: w <- w0 + (w - w0) / max(1, ||w - w0|| / clip)   (src/agent.py:54-60), no host sync for the norm
__global__ void __launch_bounds__(256) pgd_project_kernel(float* __restrict__ w, const float* __restrict__ w0,
                                                            __nv_bfloat16* __restrict__ wb, long long n4, float clip,
                                                            const double* __restrict__ d_sqnorm, long long n4_pgd) {
    const float denom = fmaxf(1.0f, (float)sqrt(*d_sqnorm) / clip);
    const float inv = 1.0f / denom;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        float4 wv = ld_f4(w + 4 * q);
        if (denom > 1.0f && q < n4_pgd) {
            const float4 o = ld_f4(w0 + 4 * q);
            wv.x = o.x + (wv.x - o.x) * inv; wv.y = o.y + (wv.y - o.y) * inv;
            wv.z = o.z + (wv.z - o.z) * inv; wv.w = o.w + (wv.w - o.w) * inv;
            st_f4(w + 4 * q, wv);
        }
        if (wb) *reinterpret_cast<uint2*>(wb + 4 * q) = make_uint2(pack_bf16x2(wv.x, wv.y), pack_bf16x2(wv.z, wv.w));
    }
}
cudaError_t launch_pgd_project(float* w, const float* w0, __nv_bfloat16* w_bf16, long long n, float clip,
                               const double* d_sqnorm, int num_sms, cudaStream_t st, long long n_pgd) {
    if ((n & 3) || (n_pgd & 3)) return cudaErrorInvalidValue;
    if (n_pgd <= 0 || n_pgd > n) n_pgd = n;
    pgd_project_kernel<<<grid_for(n / 4, 256, num_sms, 4), 256, 0, st>>>(w, w0, w_bf16, n / 4, clip, d_sqnorm, n_pgd / 4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// softmax cross-entropy forward+backward (mean reduction): one thread per row, C <= 32
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void softmax_xent_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, T* __restrict__ dlogits,
                                    float* loss_sum, int* correct, int B, int C, float grad_scale) {
    __shared__ float scratch[32];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f;
    int ok = 0;
    if (b < B) {
        float v[32];
        float mx = -INFINITY;
        int arg = 0;
        for (int c = 0; c < C; ++c) {
            v[c] = (float)logits[(int64_t)b * C + c];
            if (v[c] > mx) { mx = v[c]; arg = c; }
        }
        float se = 0.f;
        for (int c = 0; c < C; ++c) { v[c] = __expf(v[c] - mx); se += v[c]; }
        const int y = (int)labels[b];
        const float inv = 1.0f / se;
        loss = -(__logf(v[y] * inv + 1e-30f));
        ok = (arg == y);
        if (dlogits)
            for (int c = 0; c < C; ++c) dlogits[(int64_t)b * C + c] = (T)((v[c] * inv - (c == y ? 1.f : 0.f)) * grad_scale);
    }
    const float tot = block_sum<float>(loss, scratch);
    if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, tot);
    if (correct) {
        const int nok = __syncthreads_count(ok);
        if (threadIdx.x == 0 && nok) atomicAdd(correct, nok);
    }
}
cudaError_t launch_softmax_xent(const void* logits, int kind, const int64_t* labels, void* dlogits, float* loss_sum,
                                int* correct, int B, int C, float grad_scale, cudaStream_t st) {
    if (C > 32 || B <= 0) return cudaErrorInvalidValue;
    const int threads = 128, blocks = (B + threads - 1) / threads;
    if (kind == 0) softmax_xent_kernel<float><<<blocks, threads, 0, st>>>((const float*)logits, labels, (float*)dlogits, loss_sum, correct, B, C, grad_scale);
    else softmax_xent_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>((const __nv_bfloat16*)logits, labels, (__nv_bfloat16*)dlogits, loss_sum, correct, B, C, grad_scale);
    return cudaGetLastError();
}

// evaluation: sum of per-sample losses + confusion matrix, all on device (reference loops over samples on the
// host with .item() syncs, src/utils.py:144-152)
template <typename T>
__global__ void eval_metrics_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, int B, int C,
                                    double* loss_sum, long long* confusion) {
    __shared__ double scratch[32];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double loss = 0.0;
    if (b < B) {
        float mx = -INFINITY;
        int arg = 0;
        for (int c = 0; c < C; ++c) {
            const float v = (float)logits[(int64_t)b * C + c];
            if (v > mx) { mx = v; arg = c; }
        }
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf((float)logits[(int64_t)b * C + c] - mx);
        const int y = (int)labels[b];
        loss = (double)(logf(se) - ((float)logits[(int64_t)b * C + y] - mx));
        atomicAdd((unsigned long long*)(confusion + (int64_t)y * C + arg), 1ull);
    }
    const double tot = block_sum<double>(loss, scratch);
    if (threadIdx.x == 0) atomicAdd(loss_sum, tot);
}
cudaError_t launch_eval_metrics(const void* logits, int kind, const int64_t* labels, int B, int C, double* loss_sum,
                                long long* confusion, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    const int threads = 128, blocks = (B + threads - 1) / threads;
    if (kind == 0) eval_metrics_kernel<float><<<blocks, threads, 0, st>>>((const float*)logits, labels, B, C, loss_sum, confusion);
    else eval_metrics_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>((const __nv_bfloat16*)logits, labels, B, C, loss_sum, confusion);
    return cudaGetLastError();
}

}  // namespace rlr
