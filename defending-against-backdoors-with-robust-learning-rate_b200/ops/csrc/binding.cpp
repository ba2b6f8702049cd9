// Python bindings (pybind11 through torch/extension.h) for the sm_100a kernels and the small native runtime
// pieces (CUDA-IPC symmetric buffers).  All launches go to the CURRENT torch CUDA stream so they compose with
// stream capture (CUDA graphs) and with torch ops on the same stream.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kernels.h"
#include "gemm.h"

namespace {

inline cudaStream_t cur_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }
inline int num_sms() { return at::cuda::getCurrentDeviceProperties()->multiProcessorCount; }
inline void check(cudaError_t e, const char* what) {
    TORCH_CHECK(e == cudaSuccess, what, ": ", cudaGetErrorString(e));
}
template <typename T>
inline T* ptr_or_null(const c10::optional<at::Tensor>& t) {
    return t.has_value() && t->defined() ? reinterpret_cast<T*>(t->data_ptr()) : nullptr;
}
#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda() && (x).is_contiguous(), #x " must be a contiguous CUDA tensor")

// ---------------------------------------------------------------------------------------------------------------
void fused_aggregate(at::Tensor w_agent_ptrs, at::Tensor weights, c10::optional<at::Tensor> scales, double total_weight,
                     int64_t w_global_ptr, at::Tensor out_ptrs, c10::optional<at::Tensor> out_bf16_ptrs,
                     bool use_multimem, int64_t begin, int64_t end, int64_t n_vote, int64_t mode, int64_t theta,
                     double server_lr, double noise_std, int64_t seed, int64_t noise_stream,
                     c10::optional<at::Tensor> flipped, c10::optional<at::Tensor> flag_ptrs,
                     c10::optional<at::Tensor> local_sync, int64_t rank, int64_t world, int64_t epoch, bool handoff) {
    CHECK_CUDA(w_agent_ptrs); CHECK_CUDA(weights); CHECK_CUDA(out_ptrs);
    TORCH_CHECK(w_agent_ptrs.scalar_type() == at::kLong && out_ptrs.scalar_type() == at::kLong, "pointer tables must be int64");
    TORCH_CHECK(weights.scalar_type() == at::kDouble, "weights must be float64");
    c10::cuda::CUDAGuard guard(w_agent_ptrs.device());
    rlr::AggParams p{};
    p.w_agents = reinterpret_cast<const float* const*>(w_agent_ptrs.data_ptr());
    p.weights = weights.data_ptr<double>();
    p.scales = ptr_or_null<const float>(scales);
    p.total_weight = total_weight;
    p.w_global = reinterpret_cast<const float*>(w_global_ptr);
    p.out_ptrs = reinterpret_cast<float* const*>(out_ptrs.data_ptr());
    p.out_bf16_ptrs = ptr_or_null<__nv_bfloat16* const>(out_bf16_ptrs);
    p.n_out = (int)out_ptrs.numel();
    p.use_multimem = use_multimem ? 1 : 0;
    p.K = (int)w_agent_ptrs.numel();
    p.begin = begin; p.end = end; p.n_vote = n_vote;
    p.mode = (int)mode; p.theta = (int)theta;
    p.server_lr = (float)server_lr; p.noise_std = (float)noise_std;
    p.seed = (uint64_t)seed; p.noise_stream = (uint64_t)noise_stream;
    p.flipped = ptr_or_null<unsigned long long>(flipped);
    p.flag_ptrs = ptr_or_null<uint32_t* const>(flag_ptrs);
    p.local_sync = ptr_or_null<uint32_t>(local_sync);
    p.rank = (int)rank; p.world = (int)world; p.epoch = (uint32_t)epoch;
    p.handoff = handoff ? 1 : 0;
    TORCH_CHECK(world <= 1 || (p.flag_ptrs && p.local_sync), "multi-GPU aggregation needs flag_ptrs and local_sync");
    check(rlr::launch_fused_aggregate(p, num_sms(), cur_stream()), "fused_aggregate");
}

// wait for broadcast slices [first, last] of round `epoch` (ready_ptr = address of this rank's ready words, 0 = nothing to wait for) and
// copy the BatchNorm-statistics tail of the broadcast buffer into the trainer's parameters
void acquire_slices(int64_t ready_ptr, int64_t first, int64_t last, c10::optional<at::Tensor> epoch, c10::optional<at::Tensor> tail_src,
                    c10::optional<at::Tensor> tail_dst) {
    TORCH_CHECK(ready_ptr == 0 || (epoch.has_value() && epoch->defined() && epoch->scalar_type() == at::kInt), "epoch must be an int32 device tensor");
    const float* src = ptr_or_null<const float>(tail_src);
    float* dst = ptr_or_null<float>(tail_dst);
    const long long n = (src && dst) ? (long long)tail_dst->numel() : 0;
    TORCH_CHECK(!(src && dst) || tail_src->numel() == tail_dst->numel(), "acquire_slices: tail size mismatch");
    check(rlr::launch_acquire_slices(reinterpret_cast<const uint32_t*>(ready_ptr), (int)first, (int)last,
                                     ready_ptr ? reinterpret_cast<const uint32_t*>(epoch->data_ptr()) : nullptr, src, dst, n, cur_stream()),
          "acquire_slices");
}

void update_sqnorm(at::Tensor w_agent_ptrs, int64_t w_global_ptr, int64_t n, at::Tensor out) {
    CHECK_CUDA(w_agent_ptrs); CHECK_CUDA(out);
    TORCH_CHECK(out.scalar_type() == at::kDouble && out.numel() >= w_agent_ptrs.numel());
    c10::cuda::CUDAGuard guard(out.device());
    check(rlr::launch_update_sqnorm(reinterpret_cast<const float* const*>(w_agent_ptrs.data_ptr()),
                                    reinterpret_cast<const float*>(w_global_ptr), n, (int)w_agent_ptrs.numel(),
                                    out.data_ptr<double>(), num_sms(), cur_stream()), "update_sqnorm");
}

// ---------------------------------------------------------------------------------------------------------------
void gather_normalize(at::Tensor data, at::Tensor idx, c10::optional<at::Tensor> cursor, c10::optional<at::Tensor> targets,
                      at::Tensor out, c10::optional<at::Tensor> out_labels, int64_t B, int64_t c_pad, bool nchw,
                      std::vector<double> mean, std::vector<double> stdv) {
    CHECK_CUDA(data); CHECK_CUDA(idx); CHECK_CUDA(out);
    TORCH_CHECK(data.dim() == 4 && idx.scalar_type() == at::kLong);
    const int H = data.size(1), W = data.size(2), C = data.size(3);
    TORCH_CHECK((int)mean.size() == C && (int)stdv.size() == C);
    const int in_is_float = data.scalar_type() == at::kFloat;
    TORCH_CHECK(in_is_float || data.scalar_type() == at::kByte, "dataset must be uint8 or float32");
    const int out_kind = out.scalar_type() == at::kFloat ? 0 : 1;
    TORCH_CHECK(out_kind == 0 || out.scalar_type() == at::kBFloat16, "out must be fp32 or bf16");
    TORCH_CHECK(out.numel() >= B * H * W * (nchw ? C : c_pad), "out too small");
    float mu[4], sd[4];
    for (int c = 0; c < C; ++c) { mu[c] = (float)mean[c]; sd[c] = (float)stdv[c]; }
    c10::cuda::CUDAGuard guard(data.device());
    check(rlr::launch_gather_normalize(data.data_ptr(), in_is_float, idx.data_ptr<int64_t>(), ptr_or_null<const int>(cursor),
                                       ptr_or_null<const int64_t>(targets), out.data_ptr(), out_kind,
                                       ptr_or_null<int64_t>(out_labels), (int)B, H, W, C, (int)c_pad, nchw ? 1 : 0, mu, sd,
                                       cur_stream()), "gather_normalize");
}

// gather + normalise + im2col: A[B*Ho*Wo][64] (bf16) for a k x k / pad stem convolution over data[N,H,W,C]  (C*k*k <= 64)
void gather_im2col(at::Tensor data, at::Tensor idx, c10::optional<at::Tensor> cursor, c10::optional<at::Tensor> targets, at::Tensor A,
                   c10::optional<at::Tensor> out_labels, int64_t B, int64_t k, int64_t pad, std::vector<double> mean, std::vector<double> stdv) {
    CHECK_CUDA(data); CHECK_CUDA(idx); CHECK_CUDA(A);
    TORCH_CHECK(data.dim() == 4 && idx.scalar_type() == at::kLong && A.scalar_type() == at::kBFloat16);
    const int H = data.size(1), W = data.size(2), C = data.size(3);
    TORCH_CHECK((int)mean.size() == C && (int)stdv.size() == C && C * k * k <= 64);
    const int in_is_float = data.scalar_type() == at::kFloat;
    TORCH_CHECK(in_is_float || data.scalar_type() == at::kByte, "dataset must be uint8 or float32");
    const int64_t Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
    TORCH_CHECK(A.numel() >= B * Ho * Wo * 64, "A too small");
    float mu[4], sd[4];
    for (int c = 0; c < C; ++c) { mu[c] = (float)mean[c]; sd[c] = (float)stdv[c]; }
    c10::cuda::CUDAGuard guard(data.device());
    check(rlr::launch_gather_im2col(data.data_ptr(), in_is_float, idx.data_ptr<int64_t>(), ptr_or_null<const int>(cursor),
                                    ptr_or_null<const int64_t>(targets), reinterpret_cast<__nv_bfloat16*>(A.data_ptr()),
                                    ptr_or_null<int64_t>(out_labels), (int)B, H, W, C, (int)k, (int)pad, mu, sd, cur_stream()), "gather_im2col");
}

void stamp_pixels(at::Tensor data, at::Tensor sel, at::Tensor rows, at::Tensor cols, at::Tensor vals, int64_t mode) {
    CHECK_CUDA(data); CHECK_CUDA(sel); CHECK_CUDA(rows); CHECK_CUDA(cols); CHECK_CUDA(vals);
    TORCH_CHECK(data.dim() == 4 && sel.scalar_type() == at::kLong && rows.scalar_type() == at::kInt &&
                cols.scalar_type() == at::kInt && vals.scalar_type() == at::kFloat);
    const int is_float = data.scalar_type() == at::kFloat;
    TORCH_CHECK(is_float || data.scalar_type() == at::kByte);
    c10::cuda::CUDAGuard guard(data.device());
    check(rlr::launch_stamp_pixels(data.data_ptr(), is_float, sel.data_ptr<int64_t>(), (int)sel.numel(), rows.data_ptr<int>(),
                                   cols.data_ptr<int>(), vals.data_ptr<float>(), (int)rows.numel(), data.size(1), data.size(2),
                                   data.size(3), (int)mode, cur_stream()), "stamp_pixels");
}

void advance_cursor(at::Tensor cursor, int64_t delta, c10::optional<at::Tensor> step) {
    CHECK_CUDA(cursor);
    TORCH_CHECK(cursor.scalar_type() == at::kInt);
    TORCH_CHECK(!(step.has_value() && step->defined()) || step->scalar_type() == at::kLong, "step counter must be int64");
    c10::cuda::CUDAGuard guard(cursor.device());
    check(rlr::launch_advance_cursor(cursor.data_ptr<int>(), (int)delta, reinterpret_cast<long long*>(ptr_or_null<int64_t>(step)),
                                     cur_stream()), "advance_cursor");
}

// zero a contiguous tensor with a memset node on the current stream (no fill kernel inside captured steps)
void memset_zero(at::Tensor t) {
    CHECK_CUDA(t);
    c10::cuda::CUDAGuard guard(t.device());
    check(cudaMemsetAsync(t.data_ptr(), 0, (size_t)t.numel() * t.element_size(), cur_stream()), "memset_zero");
}

void pad_rows(at::Tensor src, at::Tensor dst) {
    CHECK_CUDA(src); CHECK_CUDA(dst);
    TORCH_CHECK(src.scalar_type() == at::kBFloat16 && dst.scalar_type() == at::kBFloat16 && src.dim() == 2 && dst.dim() == 2 &&
                src.size(0) == dst.size(0), "pad_rows: bf16 [R,K] -> [R,Kp]");
    c10::cuda::CUDAGuard guard(src.device());
    check(rlr::launch_pad_rows(reinterpret_cast<const __nv_bfloat16*>(src.data_ptr()), reinterpret_cast<__nv_bfloat16*>(dst.data_ptr()),
                               src.size(0), (int)src.size(1), (int)dst.size(1), num_sms(), cur_stream()), "pad_rows");
}

void unpad_add(at::Tensor src, at::Tensor dst) {
    CHECK_CUDA(src); CHECK_CUDA(dst);
    TORCH_CHECK(src.scalar_type() == at::kFloat && dst.scalar_type() == at::kFloat && src.dim() == 2 && dst.dim() == 2 &&
                src.size(0) == dst.size(0), "unpad_add: fp32 [R,Kp] -> [R,K]");
    c10::cuda::CUDAGuard guard(src.device());
    check(rlr::launch_unpad_add(src.data_ptr<float>(), dst.data_ptr<float>(), src.size(0), (int)dst.size(1), (int)src.size(1), num_sms(),
                                cur_stream()), "unpad_add");
}

// ---------------------------------------------------------------------------------------------------------------
void round_init(at::Tensor w_global, c10::optional<at::Tensor> w_local, c10::optional<at::Tensor> w_bf16,
                c10::optional<at::Tensor> mom) {
    CHECK_CUDA(w_global);
    c10::cuda::CUDAGuard guard(w_global.device());
    check(rlr::launch_round_init(w_global.data_ptr<float>(), ptr_or_null<float>(w_local), ptr_or_null<__nv_bfloat16>(w_bf16),
                                 ptr_or_null<float>(mom), w_global.numel(), cur_stream()), "round_init");
}

void sqnorm(at::Tensor x, at::Tensor out) {
    CHECK_CUDA(x); CHECK_CUDA(out);
    TORCH_CHECK(x.scalar_type() == at::kFloat && out.scalar_type() == at::kDouble);
    c10::cuda::CUDAGuard guard(x.device());
    check(rlr::launch_sqnorm(x.data_ptr<float>(), x.numel(), out.data_ptr<double>(), num_sms(), cur_stream()), "sqnorm");
}

void sgd_step(at::Tensor w, at::Tensor g, at::Tensor m, c10::optional<at::Tensor> w0, c10::optional<at::Tensor> w_bf16,
              double lr, double momentum, double max_grad_norm, c10::optional<at::Tensor> g_sqnorm,
              c10::optional<at::Tensor> d_sqnorm, int64_t n_pgd, c10::optional<at::Tensor> w_in) {
    CHECK_CUDA(w); CHECK_CUDA(g); CHECK_CUDA(m);
    TORCH_CHECK(w.numel() == g.numel() && w.numel() == m.numel());
    TORCH_CHECK(!(d_sqnorm.has_value() && d_sqnorm->defined()) || (w0.has_value() && w0->defined()), "PGD needs w0");
    c10::cuda::CUDAGuard guard(w.device());
    check(rlr::launch_sgd_step(w.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), ptr_or_null<const float>(w0),
                               ptr_or_null<__nv_bfloat16>(w_bf16), w.numel(), (float)lr, (float)momentum,
                               (float)max_grad_norm, ptr_or_null<const double>(g_sqnorm), ptr_or_null<double>(d_sqnorm),
                               num_sms(), cur_stream(), n_pgd, ptr_or_null<const float>(w_in)), "sgd_step");
}

void pgd_project(at::Tensor w, at::Tensor w0, c10::optional<at::Tensor> w_bf16, double clip, at::Tensor d_sqnorm, int64_t n_pgd) {
    CHECK_CUDA(w); CHECK_CUDA(w0); CHECK_CUDA(d_sqnorm);
    c10::cuda::CUDAGuard guard(w.device());
    check(rlr::launch_pgd_project(w.data_ptr<float>(), w0.data_ptr<float>(), ptr_or_null<__nv_bfloat16>(w_bf16), w.numel(),
                                  (float)clip, d_sqnorm.data_ptr<double>(), num_sms(), cur_stream(), n_pgd), "pgd_project");
}

// ---------------------------------------------------------------------------------------------------------------
void softmax_xent(at::Tensor logits, at::Tensor labels, c10::optional<at::Tensor> dlogits, c10::optional<at::Tensor> loss_sum,
                  c10::optional<at::Tensor> correct, double grad_scale) {
    CHECK_CUDA(logits); CHECK_CUDA(labels);
    TORCH_CHECK(logits.dim() == 2 && labels.scalar_type() == at::kLong);
    const int kind = logits.scalar_type() == at::kFloat ? 0 : 1;
    TORCH_CHECK(kind == 0 || logits.scalar_type() == at::kBFloat16);
    c10::cuda::CUDAGuard guard(logits.device());
    check(rlr::launch_softmax_xent(logits.data_ptr(), kind, labels.data_ptr<int64_t>(),
                                   dlogits.has_value() && dlogits->defined() ? dlogits->data_ptr() : nullptr,
                                   ptr_or_null<float>(loss_sum), ptr_or_null<int>(correct), (int)logits.size(0),
                                   (int)logits.size(1), (float)grad_scale, cur_stream()), "softmax_xent");
}

void eval_metrics(at::Tensor logits, at::Tensor labels, at::Tensor loss_sum, at::Tensor confusion) {
    CHECK_CUDA(logits); CHECK_CUDA(labels); CHECK_CUDA(loss_sum); CHECK_CUDA(confusion);
    TORCH_CHECK(loss_sum.scalar_type() == at::kDouble && confusion.scalar_type() == at::kLong);
    const int kind = logits.scalar_type() == at::kFloat ? 0 : 1;
    c10::cuda::CUDAGuard guard(logits.device());
    check(rlr::launch_eval_metrics(logits.data_ptr(), kind, labels.data_ptr<int64_t>(), (int)logits.size(0), (int)logits.size(1),
                                   loss_sum.data_ptr<double>(), reinterpret_cast<long long*>(confusion.data_ptr<int64_t>()),
                                   cur_stream()), "eval_metrics");
}

// ---------------------------------------------------------------------------------------------------------------
// CUDA-IPC symmetric buffers: cudaMalloc'ed slabs whose handles are exchanged through torch.distributed and opened
// by every peer (fallback when torch's symmetric memory / multicast is unavailable).
// ---------------------------------------------------------------------------------------------------------------
std::pair<int64_t, py::bytes> ipc_alloc(int64_t nbytes, int64_t device) {
    c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
    void* p = nullptr;
    check(cudaMalloc(&p, (size_t)nbytes), "ipc_alloc/cudaMalloc");
    check(cudaMemset(p, 0, (size_t)nbytes), "ipc_alloc/cudaMemset");
    cudaIpcMemHandle_t h;
    check(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    return {reinterpret_cast<int64_t>(p), py::bytes(reinterpret_cast<const char*>(&h), sizeof(h))};
}
int64_t ipc_open(const std::string& handle, int64_t device) {
    TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
    c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle.data(), sizeof(h));
    void* p = nullptr;
    check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return reinterpret_cast<int64_t>(p);
}
void ipc_close(int64_t ptr) { check(cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)), "cudaIpcCloseMemHandle"); }
void ipc_free(int64_t ptr) { check(cudaFree(reinterpret_cast<void*>(ptr)), "cudaFree"); }

at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, at::ScalarType dtype, int64_t device) {
    auto opts = at::TensorOptions().dtype(dtype).device(at::kCUDA, (c10::DeviceIndex)device);
    return at::from_blob(reinterpret_cast<void*>(ptr), sizes, [](void*) {}, opts);
}

}  // namespace

void register_gemm_bindings(py::module_& m);  // gemm_binding.cpp

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "b200-robust-fl sm_100a kernels";
    m.def("fused_aggregate", &fused_aggregate, py::arg("w_agent_ptrs"), py::arg("weights"), py::arg("scales"), py::arg("total_weight"),
          py::arg("w_global_ptr"), py::arg("out_ptrs"), py::arg("out_bf16_ptrs"), py::arg("use_multimem"), py::arg("begin"), py::arg("end"),
          py::arg("n_vote"), py::arg("mode"), py::arg("theta"), py::arg("server_lr"), py::arg("noise_std"), py::arg("seed"),
          py::arg("noise_stream"), py::arg("flipped"), py::arg("flag_ptrs"), py::arg("local_sync"), py::arg("rank"), py::arg("world"),
          py::arg("epoch"), py::arg("handoff") = false);
    m.def("aggregate_max_agents", &rlr::aggregate_max_agents);
    m.def("acquire_slices", &acquire_slices);
    m.def("update_sqnorm", &update_sqnorm);
    m.def("gather_normalize", &gather_normalize);
    m.def("gather_im2col", &gather_im2col);
    m.def("stamp_pixels", &stamp_pixels);
    m.def("advance_cursor", &advance_cursor, py::arg("cursor"), py::arg("delta"), py::arg("step") = py::none());
    m.def("memset_zero", &memset_zero);
    m.def("pad_rows", &pad_rows);
    m.def("unpad_add", &unpad_add);
    m.def("round_init", &round_init);
    m.def("sqnorm", &sqnorm);
    m.def("sgd_step", &sgd_step, py::arg("w"), py::arg("g"), py::arg("m"), py::arg("w0"), py::arg("w_bf16"), py::arg("lr"),
          py::arg("momentum"), py::arg("max_grad_norm"), py::arg("g_sqnorm"), py::arg("d_sqnorm"), py::arg("n_pgd") = 0, py::arg("w_in") = py::none());
    m.def("pgd_project", &pgd_project, py::arg("w"), py::arg("w0"), py::arg("w_bf16"), py::arg("clip"), py::arg("d_sqnorm"),
          py::arg("n_pgd") = 0);
    m.def("softmax_xent", &softmax_xent);
    m.def("eval_metrics", &eval_metrics);
    m.def("ipc_alloc", &ipc_alloc);
    m.def("ipc_open", &ipc_open);
    m.def("ipc_close", &ipc_close);
    m.def("ipc_free", &ipc_free);
    m.def("tensor_from_ptr", &tensor_from_ptr);
    register_gemm_bindings(m);
}
