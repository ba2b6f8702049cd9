"""Operator layer.

Every op has two implementations behind one Python signature:

* **native** -- a hand-written sm_100a kernel from ``ops/_C.so`` (sources in ``ops/csrc``), used whenever the tensors
  live on a CUDA device.  If the extension is missing on a GPU box the op raises -- there is no silent eager fallback.
* **oracle** -- a plain fp32/fp64 PyTorch re-statement of the same semantics, used on CPU (the plumbing config and the
  CPU test-suite) and as the numerical reference the GPU tests compare the kernels against.

The reference has no operator layer at all: each row of that table maps to one function here.
"""
from __future__ import annotations

import importlib.util
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_C.so")
_ext = None
_ext_err = None
_lock = threading.Lock()

MODE_IDS = {"avg": 0, "comed": 1, "sign": 2}
MAX_FUSED_AGENTS = 1024   # kMaxAgents of ops/csrc/aggregate.cu (participant tables of the fused kernel)


class _Counter:
    """Counts calls into the native extension (each call launches >= 1 of our kernels); used for ``gpu_launches``."""
    calls = 0


class _CountingExt:
    def __init__(self, mod):
        self._mod = mod
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            real = getattr(self._mod, name)
            if not callable(real):
                return real

            def fn(*a, _real=real, **k):
                _Counter.calls += 1
                return _real(*a, **k)
            self._cache[name] = fn
        return fn


def launch_calls() -> int:
    return _Counter.calls


# ---- library fall-throughs ---------------------------------------------------------------------------------------------
# The sm100 back-end of a layer primitive may meet a shape no kernel of ours covers and fall through to a PyTorch library call
# (cuBLAS / cuDNN / ATen).  Every such fall-through is recorded here by call site; ``RLR_STRICT=1`` turns it into an error.  The
# bench prints the counters (``library_fallbacks``) and tests/test_gpu_native.py asserts they stay empty for every zoo model.
_fallbacks: dict = {}


def note_fallback(site: str, detail: str = ""):
    _fallbacks[site] = _fallbacks.get(site, 0) + 1
    if os.environ.get("RLR_STRICT", "0") == "1":
        raise RuntimeError(f"RLR_STRICT=1: sm100 back-end fell through to a library call at {site} {detail}")


def fallback_calls() -> dict:
    """{call site: count} of library fall-throughs of the sm100 back-end since the last reset (empty = none)."""
    return dict(_fallbacks)


def reset_fallbacks():
    _fallbacks.clear()


def zero_(t):
    """Zero a tensor: a memset node on CUDA (no ATen fill kernel inside captured steps), ``zero_()`` on CPU."""
    if t.is_cuda and t.is_contiguous():
        ext().memset_zero(t)
    else:
        t.zero_()
    return t


def native_available() -> bool:
    """True if the compiled extension can be imported (it may still be unusable without a GPU)."""
    try:
        ext()
        return True
    except Exception:  # noqa: BLE001
        return False


def ext():
    """The compiled extension module; raises with build instructions if it is missing."""
    global _ext, _ext_err
    if _ext is not None:
        return _ext
    with _lock:
        if _ext is not None:
            return _ext
        if not os.path.exists(_SO):
            raise RuntimeError(f"native extension {_SO} not built; run `python -m rlr_b200.ops.build` "
                               "(or __graft_entry__.build())")
        spec = importlib.util.spec_from_file_location("rlr_b200.ops._C", _SO)
        mod = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception as e:  # noqa: BLE001
            _ext_err = e
            raise
        _ext = _CountingExt(mod)
    return _ext


def _i32(x, device):
    return torch.as_tensor(x, dtype=torch.int32, device=device)


# =====================================================================================================================
# data path
# =====================================================================================================================
def gather_normalize(data, idxs, mean, std, dtype=torch.float32, nhwc=False, c_pad=None, out=None,
                     cursor=None, targets=None, out_labels=None, batch=None):
    """``normalize(data[idxs])``: raw NHWC uint8/float pixels -> fp32/bf16 batch.

    ``data`` [N,H,W,C]; ``idxs`` int64 sample indices (with ``cursor``: a device int32 offset into ``idxs`` so a CUDA
    graph can replay the launch over successive batches; ``batch`` = batch size then).  Output NCHW, or NHWC padded to
    ``c_pad`` channels when ``nhwc``.  Same arithmetic as ``ToTensor`` + ``Normalize`` (src/utils.py:101,112-115).
    """
    N, H, W, C = data.shape
    B = int(batch if batch is not None else idxs.shape[0])
    c_pad = int(c_pad or C)
    if out is None:
        shape = (B, H, W, c_pad) if nhwc else (B, C, H, W)
        out = torch.empty(shape, dtype=dtype, device=data.device)
    if data.is_cuda:
        ext().gather_normalize(data, idxs, cursor, targets, out, out_labels, B, c_pad, not nhwc,
                               [float(m) for m in mean], [float(s) for s in std])
        return out
    off = int(cursor.item()) if cursor is not None else 0
    sel = idxs[off:off + B]
    x = data[sel].to(torch.float32)
    if data.dtype == torch.uint8:
        x = x / 255.0
    x = (x - torch.tensor(mean, dtype=torch.float32)) / torch.tensor(std, dtype=torch.float32)
    if nhwc:
        if c_pad > C:
            x = torch.nn.functional.pad(x, (0, c_pad - C))
        out.copy_(x.to(out.dtype))
    else:
        out.copy_(x.permute(0, 3, 1, 2).to(out.dtype))
    if out_labels is not None and targets is not None:
        out_labels[:B] = targets[sel]
    return out


def gather_im2col(data, idxs, mean, std, k, pad, out, cursor=None, targets=None, out_labels=None, batch=None):
    """Batch assembly fused with the first layer's im2col (SURVEY.md K1+K2, tiny-K stems: C*k*k <= 64): row (b, ho, wo) of ``out``
    [B*Ho*Wo, 64] (bf16) is the k x k x C patch of the normalised image around that output pixel in (tap, channel) order, zero
    padded -- the A operand of the stem convolution as a single-k-block tcgen05 GEMM.  Same cursor / label contract as
    ``gather_normalize``."""
    N, H, W, C = data.shape
    B = int(batch if batch is not None else idxs.shape[0])
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    if data.is_cuda:
        ext().gather_im2col(data, idxs, cursor, targets, out, out_labels, B, int(k), int(pad), [float(m) for m in mean], [float(s) for s in std])
        return out
    off = int(cursor.item()) if cursor is not None else 0
    sel = idxs[off:off + B]
    x = data[sel].to(torch.float32)
    if data.dtype == torch.uint8:
        x = x / 255.0
    x = (x - torch.tensor(mean, dtype=torch.float32)) / torch.tensor(std, dtype=torch.float32)          # [B,H,W,C]
    xp = torch.nn.functional.pad(x, (0, 0, pad, pad, pad, pad))
    cols = [xp[:, dy:dy + Ho, dx:dx + Wo, :] for dy in range(k) for dx in range(k)]                       # (tap, channel) order
    A = torch.cat(cols, dim=-1).reshape(B * Ho * Wo, k * k * C)
    out[:B * Ho * Wo].zero_()
    out[:B * Ho * Wo, :k * k * C] = A.to(out.dtype)
    if out_labels is not None and targets is not None:
        out_labels[:B] = targets[sel]
    return out


def stamp_pixels(data, sel, rows, cols, vals, mode):
    """Apply a compiled trojan pixel program to images ``sel`` of ``data`` [N,H,W,C] in place."""
    if len(rows) == 0 or sel.numel() == 0:
        return data
    if data.is_cuda:
        dev = data.device
        ext().stamp_pixels(data, sel.to(dev, torch.int64).contiguous(), _i32(rows, dev), _i32(cols, dev),
                           torch.as_tensor(vals, dtype=torch.float32, device=dev), int(mode))
        return data
    r = torch.as_tensor(rows, dtype=torch.int64)
    c = torch.as_tensor(cols, dtype=torch.int64)
    v = torch.as_tensor(vals, dtype=torch.float32)
    s = sel.to(torch.int64)[:, None]
    cur = data[s, r[None, :], c[None, :], :]                      # [S,P,C]
    if mode == 0:
        new = v[None, :, None].expand_as(cur).to(data.dtype)
    elif mode == 1:  # uint8 wrap-around add
        new = ((cur.to(torch.int64) + v[None, :, None].to(torch.int64)) % 256).to(data.dtype)
    else:
        new = (cur.to(torch.float32) - v[None, :, None]).to(data.dtype)
    data[s, r[None, :], c[None, :], :] = new
    return data


# =====================================================================================================================
# server step
# =====================================================================================================================
def aggregate_oracle(w_global, w_agents, weights, mode="avg", theta=0, server_lr=1.0, noise=None, n_vote=None,
                     scales=None):
    """fp64 PyTorch statement of the server step (reference src/aggregation.py:19-75).

    ``w_agents``: list of local parameter vectors; updates are ``w_k - w_global``.  ``noise``: optional pre-sampled
    noise vector (added to the aggregate BEFORE the lr multiply).  Coordinates ``>= n_vote`` get a plain weighted mean.
    Returns ``(new_global_fp32, n_flipped)``.
    """
    g = w_global.double()
    n = g.numel()
    n_vote = n if n_vote is None else int(n_vote)
    ups = [(w.double() - g) for w in w_agents]
    wt = torch.as_tensor(weights, dtype=torch.float64, device=g.device)
    mean_raw = sum(w_ * u for w_, u in zip(wt, ups)) / wt.sum()     # tail coordinates (BatchNorm statistics) are never clip-scaled
    if scales is not None:
        ups = [u * float(s) for u, s in zip(ups, scales)]
    mean = sum(w_ * u for w_, u in zip(wt, ups)) / wt.sum()
    signs = sum(torch.sign(u) for u in ups)
    if mode == "avg":
        agg = mean.clone()
    elif mode == "comed":
        agg = torch.median(torch.stack(ups, dim=1), dim=1).values
    elif mode == "sign":
        agg = torch.sign(signs)
    else:
        raise ValueError(mode)
    if noise is not None:
        agg = agg + noise.double()
    lr = torch.full_like(g, float(server_lr))
    flipped = 0
    if theta > 0:
        neg = signs.abs() < theta
        neg[n_vote:] = False
        lr[neg] = -float(server_lr)
        flipped = int(neg.sum())
    new = g + lr * agg
    if n_vote < n:
        new[n_vote:] = g[n_vote:] + mean_raw[n_vote:]
    return new.float(), flipped


def aggregate_partials(w_global, w_local_agents, local_weights, n_vote=None, scales=None):
    """This rank's share of the server step for the additive aggregators: ``(vote, wsum)`` with
    ``vote = sum_k sign(w_k - w_g)`` (float32: small integers) and ``wsum = sum_k n_k (w_k - w_g)`` (float64) over the LOCAL
    participants.  Summed over ranks (all_reduce) they are exactly the ``signs`` and ``mean * sum(n)`` of ``aggregate_oracle``."""
    g = w_global.double()
    vote = torch.zeros_like(w_global, dtype=torch.float32)
    wsum = torch.zeros_like(g)
    nv = g.numel() if n_vote is None else int(n_vote)
    for i, (w, nk) in enumerate(zip(w_local_agents, local_weights)):
        u = w.double() - g
        if scales is not None:
            u = u.clone()
            u[:nv] *= float(scales[i])           # server clipping scales the voted coordinates only
        vote += torch.sign(u).float()
        wsum += float(nk) * u
    return vote, wsum


def aggregate_from_partials(w_global, vote, wsum, total_weight, mode="avg", theta=0, server_lr=1.0, noise=None, n_vote=None):
    """Finish the server step from globally reduced partials (same formulas and order as ``aggregate_oracle``; avg / sign only)."""
    if mode not in ("avg", "sign"):
        raise ValueError(f"aggregate_from_partials: mode {mode!r} is not additive (coordinate median needs every update)")
    g = w_global.double()
    n = g.numel()
    n_vote = n if n_vote is None else int(n_vote)
    mean = wsum / float(total_weight)
    signs = vote.double()
    agg = mean.clone() if mode == "avg" else torch.sign(signs)
    if noise is not None:
        agg = agg + noise.double()
    lr = torch.full_like(g, float(server_lr))
    flipped = 0
    if theta > 0:
        neg = signs.abs() < theta
        neg[n_vote:] = False
        lr[neg] = -float(server_lr)
        flipped = int(neg.sum())
    new = g + lr * agg
    if n_vote < n:
        new[n_vote:] = g[n_vote:] + mean[n_vote:]
    return new.float(), flipped


class PtrTable:
    """Device int64 table of raw pointers (kept with the tensors it points into, so they stay alive)."""

    def __init__(self, ptrs, device, keep=()):
        self.tensor = torch.tensor([int(p) for p in ptrs], dtype=torch.int64, device=device)
        self.keep = tuple(keep)


def fused_aggregate(w_global, w_agents, weights, mode="avg", theta=0, server_lr=1.0, noise_std=0.0, seed=0,
                    noise_stream=0, n_vote=None, scales=None, out=None, out_bf16=None, flipped=None):
    """Single-process fused server step: ``out <- w_global + lr ⊙ agg({w_k - w_global})`` in one kernel.

    On CUDA this launches ``fused_aggregate_kernel`` (ops/csrc/aggregate.cu); on CPU it runs the fp64 oracle (with
    torch-sampled noise).  ``out`` may alias ``w_global``.  Returns the tensor written.  The multi-GPU variant (peer
    pointers, multicast stores, in-kernel barriers) is driven by ``parallel.fused_agg.FusedAggregator``.
    """
    n = w_global.numel()
    n_vote = n if n_vote is None else int(n_vote)
    out = w_global if out is None else out
    if not w_global.is_cuda:
        noise = None
        if noise_std > 0:
            gen = torch.Generator().manual_seed(int(seed) * 1000003 + int(noise_stream))
            noise = torch.randn(n, generator=gen, dtype=torch.float64) * noise_std
            noise[n_vote:] = 0
        new, nflip = aggregate_oracle(w_global, w_agents, weights, mode, theta, server_lr, noise, n_vote, scales)
        out.copy_(new)
        if out_bf16 is not None:
            out_bf16.copy_(new.to(torch.bfloat16))
        if flipped is not None:
            flipped += nflip
        return out
    dev = w_global.device
    if len(w_agents) > MAX_FUSED_AGENTS:
        # more participants than the kernel's pointer / weight tables hold (1024): exact torch evaluation on the device (fp64, same
        # semantics) -- recorded as a library fall-through
        note_fallback("fused_aggregate", f"K={len(w_agents)} > {MAX_FUSED_AGENTS}")
        noise = None
        if noise_std > 0:
            gen = torch.Generator(device=dev).manual_seed(int(seed) * 1000003 + int(noise_stream))
            noise = torch.randn(n, generator=gen, dtype=torch.float64, device=dev) * noise_std
            noise[n_vote:] = 0
        new, nflip = aggregate_oracle(w_global, w_agents, weights, mode, theta, server_lr, noise, n_vote, scales)
        out.copy_(new)
        if out_bf16 is not None:
            out_bf16.copy_(new.to(torch.bfloat16))
        if flipped is not None:
            flipped += nflip
        return out
    assert n % 4 == 0 and n_vote % 4 == 0, "flat buffers are padded to multiples of 4"
    for w in w_agents:
        assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.numel() == n
    agents = PtrTable([w.data_ptr() for w in w_agents], dev, w_agents)
    outs = PtrTable([out.data_ptr()], dev)
    outs_b = PtrTable([out_bf16.data_ptr()], dev) if out_bf16 is not None else None
    wt = torch.as_tensor(weights, dtype=torch.float64).to(dev)
    sc = torch.as_tensor(scales, dtype=torch.float32).to(dev) if scales is not None else None
    ext().fused_aggregate(agents.tensor, wt, sc, float(sum(float(x) for x in weights)), w_global.data_ptr(), outs.tensor,
                          outs_b.tensor if outs_b else None, False, 0, n, n_vote, MODE_IDS[mode], int(theta),
                          float(server_lr), float(noise_std), int(seed), int(noise_stream), flipped, None, None, 0, 1, 0)
    return out


def update_norms(w_global, w_agents, n=None):
    """L2 norms of the agents' updates ``||w_k - w_global||`` (server clipping, src/aggregation.py:77-81, and the
    Norms/* diagnostic, :83-100) -> float64 tensor [K].  ``n``: only the first ``n`` coordinates count (the model parameters:
    BatchNorm running statistics stored behind ``n_vote`` are not part of the reference's parameter vector)."""
    n = w_global.numel() if n is None else int(n)
    if not w_global.is_cuda:
        return torch.stack([(w[:n].double() - w_global[:n].double()).norm() for w in w_agents])
    dev = w_global.device
    tab = PtrTable([w.data_ptr() for w in w_agents], dev, w_agents)
    out = torch.zeros(len(w_agents), dtype=torch.float64, device=dev)
    ext().update_sqnorm(tab.tensor, w_global.data_ptr(), n, out)
    return out.sqrt()


# =====================================================================================================================
# optimiser over flat buffers
# =====================================================================================================================
def round_init(w_global, w_local=None, w_bf16=None, mom=None):
    """Start of an agent's round: ``w_local <- w_global``, refresh the bf16 operand shadow, zero the momentum."""
    if w_global.is_cuda:
        ext().round_init(w_global, w_local, w_bf16, mom)
        return
    if w_local is not None:
        w_local.copy_(w_global)
    if w_bf16 is not None:
        w_bf16.copy_(w_global.to(torch.bfloat16))
    if mom is not None:
        mom.zero_()


class FlatSGD:
    """Fused ``clip_grad_norm_(., max_grad_norm)`` + momentum SGD + optional PGD projection over flat buffers.

    Reference: src/agent.py:37-38 (SGD, fresh momentum each round), :50 (clip 10), :54-60 (PGD onto the L2 ball of
    radius ``clip`` around the round's global params).  All norms stay on the device (the reference syncs to the host
    for ``max(1, norm/clip)``); the whole step is 2 kernels (+2 with PGD) regardless of the number of tensors.
    """

    def __init__(self, n, device, lr, momentum, max_grad_norm=10.0, pgd_clip=0.0, n_pgd=None):
        self.lr, self.momentum, self.max_grad_norm, self.pgd_clip = float(lr), float(momentum), float(max_grad_norm), float(pgd_clip)
        # the PGD ball is measured and projected over the model parameters [0, n_pgd) only (layout.n_vote): BatchNorm running
        # statistics behind them are not in the reference's parameters_to_vector() (src/agent.py:54-60)
        self.n_pgd = int(n if n_pgd is None else n_pgd)
        self.norms = torch.zeros(2, dtype=torch.float64, device=device)  # [||g||^2, ||w-w0||^2]

    def step(self, w, g, m, w0=None, w_bf16=None, w_in=None):
        """``w_in``: first local step of a round fused with the hand-off -- parameters are read from ``w_in`` (the round's global
        parameters, i.e. the broadcast buffer) instead of ``w`` and the momentum counts as zero, so no separate
        ``w <- w_global, m <- 0`` pass is needed; coordinates ``>= n_pgd`` (BatchNorm running statistics already updated in ``w``
        by this step's forward pass) keep their value."""
        if w.is_cuda:
            e = ext()
            e.memset_zero(self.norms)
            e.sqnorm(g, self.norms[0:1])
            pgd = self.pgd_clip > 0
            e.sgd_step(w, g, m, w0 if pgd else None, w_bf16, self.lr, self.momentum, self.max_grad_norm,
                       self.norms[0:1], self.norms[1:2] if pgd else None, self.n_pgd, w_in)
            if pgd:
                e.pgd_project(w, w0, w_bf16, self.pgd_clip, self.norms[1:2], self.n_pgd)
            return
        gn = g.double().norm()
        coef = min(1.0, self.max_grad_norm / (float(gn) + 1e-6)) if self.max_grad_norm > 0 else 1.0
        if w_in is not None:
            k = self.n_pgd
            m.zero_()
            m[:k].add_(g[:k], alpha=coef)
            w[:k].copy_(w_in[:k] - self.lr * m[:k])
        else:
            m.mul_(self.momentum).add_(g, alpha=coef)
            w.add_(m, alpha=-self.lr)
        if self.pgd_clip > 0:
            k = self.n_pgd
            d = w[:k] - w0[:k]
            denom = max(1.0, float(d.double().norm()) / self.pgd_clip)
            if denom > 1.0:
                w[:k].copy_(w0[:k] + d / denom)
        if w_bf16 is not None:
            w_bf16.copy_(w.to(torch.bfloat16))


# =====================================================================================================================
# loss / evaluation
# =====================================================================================================================
def softmax_xent(logits, labels, want_grad=True, loss_sum=None, correct=None, dlogits=None):
    """Fused softmax cross-entropy (mean reduction) forward + backward: returns ``(loss_sum_tensor, dlogits)`` where
    ``dlogits = (softmax - onehot) / B``.  ``dlogits``: optional pre-allocated output of the logits' dtype and
    shape (the native executor passes its gradient buffer, so the loss kernel writes the head's gradient in place)."""
    B = logits.shape[0]
    if logits.is_cuda:
        dl = (dlogits if dlogits is not None else torch.empty_like(logits)) if want_grad else None
        if loss_sum is None:
            loss_sum = torch.zeros(1, dtype=torch.float32, device=logits.device)
        ext().softmax_xent(logits.contiguous(), labels, dl, loss_sum, correct, 1.0 / B)
        return loss_sum, dl
    lf = logits.float()
    lsm = torch.log_softmax(lf, dim=1)
    loss = -lsm.gather(1, labels[:, None]).sum()
    dl = None
    if want_grad:
        dl = (lsm.exp() - torch.nn.functional.one_hot(labels, lf.shape[1]).float()) / B
        dl = dl.to(logits.dtype)
        if dlogits is not None:
            dlogits.copy_(dl)
            dl = dlogits
    if loss_sum is None:
        loss_sum = torch.zeros(1)
    loss_sum += loss
    if correct is not None:
        correct += (lf.argmax(1) == labels).sum().to(correct.dtype)
    return loss_sum, dl


def eval_metrics(logits, labels, loss_sum, confusion):
    """Accumulate the summed per-sample loss and the confusion matrix ``confusion[true, pred]`` on the device
    (replaces the per-sample host loop of src/utils.py:144-152)."""
    if logits.is_cuda:
        ext().eval_metrics(logits.contiguous(), labels, loss_sum, confusion)
        return
    lf = logits.float()
    loss_sum += torch.nn.functional.cross_entropy(lf, labels, reduction="sum").double()
    C = lf.shape[1]
    pred = lf.argmax(1)
    confusion.view(-1).index_add_(0, labels * C + pred, torch.ones_like(labels))


from .nn import (avgpool_bwd, avgpool_fwd, bn_bwd, bn_fwd, conv2d_dgrad_sm100, conv2d_fwd_sm100, conv2d_wgrad_sm100, conv_supported,  # noqa: E402,F401
                 dropout_bwd, dropout_fwd, linear_bwd, linear_fused_dropout_ok, linear_fwd, maxpool2_bwd, maxpool2_fwd, relu_bwd_, scratch, stem_geometry, STAT_SLOTS)
