"""CPU check of the HOST-SIDE orchestration in ops/nn.py (tap tables, parity planes, padding, scratch reuse, opt-in code paths)
against PyTorch convolutions, with the CUDA extension replaced by a small fp32 emulator of each kernel's documented contract
(ops/csrc/gemm_binding.cpp).  The kernels themselves are tested on the GPU (tests/test_gpu_native.py, test_gpu_experimental.py);
this suite makes sure that what Python hands them is right -- in particular for paths that could not be run on hardware yet."""
import pytest
import torch
import torch.nn.functional as F

import rlr_b200  # noqa: F401
from rlr_b200.ops import nn


def _gather(x, n_idx, hh, ww):
    """x[n_idx, hh, ww, :] with zeros outside the image; hh / ww are [Ho, Wo] index grids."""
    NB_tot, H, W, C = x.shape
    ok = (hh >= 0) & (hh < H) & (ww >= 0) & (ww < W)
    v = x[n_idx][:, hh.clamp(0, H - 1), ww.clamp(0, W - 1), :]          # [N, Ho, Wo, C]
    return v * ok[None, :, :, None]


class FakeExt:
    """fp32 emulation of the extension entry points used by the convolution wrappers."""

    def __init__(self):
        self.calls = []

    # ---- layout helpers ------------------------------------------------------------------------------------------------
    def space_to_depth(self, x, x4):
        self.calls.append("space_to_depth")
        NB = x.shape[0]
        for ph in range(2):
            for pw in range(2):
                pl = ph * 2 + pw
                x4[pl * NB:(pl + 1) * NB] = x[:, ph::2, pw::2]

    def depth_to_space(self, x4, y, accumulate, mask):
        self.calls.append("depth_to_space")
        NB = y.shape[0]
        if not accumulate:
            y.zero_()
        for ph in range(2):
            for pw in range(2):
                pl = ph * 2 + pw
                if mask >> pl & 1:
                    y[:, ph::2, pw::2] += x4[pl * NB:(pl + 1) * NB]

    def filter_transpose(self, w, wt, Cout, T, Cin):
        self.calls.append("filter_transpose")
        w3 = w.reshape(Cout, T, Cin)
        wt.copy_(w3.flip(1).permute(2, 1, 0).reshape(wt.shape))          # wt[ci][T-1-t][co] = w[co][t][ci]

    def filter_gather_transpose(self, w, wt, Cout, T, Cin, taps):
        self.calls.append("filter_gather_transpose")
        w3 = w.reshape(Cout, T, Cin)
        wt.copy_(w3[:, list(taps), :].permute(2, 1, 0).reshape(wt.shape))  # wt[ci][i][co] = w[co][taps[i]][ci]

    # ---- implicit-GEMM convolution -------------------------------------------------------------------------------------
    def _conv(self, x, w, out, NB, dh, dw, dplane, bias, relu, accumulate, wtap, w_taps_total, in_stride, out_stride, ph, pw):
        Cin = x.shape[3]
        Cout = out.shape[3]
        Ho, Wo = out.shape[1] // out_stride, out.shape[2] // out_stride
        hh, ww = torch.meshgrid(torch.arange(Ho), torch.arange(Wo), indexing="ij")
        acc = torch.zeros(NB, Ho, Wo, Cout)
        n_idx = torch.arange(NB)
        for t in range(len(dh)):
            v = _gather(x, n_idx + dplane[t] * NB, in_stride * hh + dh[t], in_stride * ww + dw[t])
            if wtap:       # MN-major B: w is [K = Cin_here][w_taps_total * Cout_here]
                Wt = w[:, wtap[t] * Cout:(wtap[t] + 1) * Cout]           # [Cin][Cout]
                acc += v @ Wt
            else:          # K-major B: w is [Cout][T * Cin]
                Wt = w[:, t * Cin:(t + 1) * Cin]                          # [Cout][Cin]
                acc += v @ Wt.t()
        if bias is not None:
            acc += bias
        if relu:
            acc = acc.clamp_min(0)
        view = out[:, ph::out_stride, pw::out_stride]
        if accumulate:
            view += acc
        else:
            view.copy_(acc)

    def conv_bf16(self, x, w, out, NB, planes, dh, dw, dplane, bias, relu, accumulate, stats, wtap, w_taps_total):
        self.calls.append("conv_bf16")
        assert x.shape[0] == planes * NB and out.shape[0] == NB and stats is None
        self._conv(x, w, out, NB, dh, dw, dplane, bias, relu, accumulate, list(wtap), w_taps_total, 1, 1, 0, 0)

    def conv_bf16_strided(self, x, w, out, dh, dw, bias, relu, accumulate, wtap, w_taps_total, in_stride, out_stride, ph, pw):
        self.calls.append("conv_bf16_strided")
        assert x.shape[0] == out.shape[0]
        self._conv(x, w, out, x.shape[0], dh, dw, [0] * len(dh), bias, relu, accumulate, list(wtap), w_taps_total, in_stride, out_stride, ph, pw)

    def conv3x3_halo_bf16(self, x, w, out, bias, relu, accumulate, stats, bo_mode, dbg):
        self.calls.append("conv3x3_halo_bf16")
        dh = [d - 1 for d in range(3) for _ in range(3)]
        dw = [d - 1 for _ in range(3) for d in range(3)]
        self._conv(x, w, out, x.shape[0], dh, dw, [0] * 9, bias, relu, accumulate, [], 0, 1, 1, 0, 0)

    def conv3x3_halo3_bf16(self, x, w, out, bias, relu, accumulate):
        self.calls.append("conv3x3_halo3_bf16")
        self.conv3x3_halo_bf16(x, w, out, bias, relu, accumulate, None, 0, None)
        self.calls.pop()

    # ---- weight gradients ---------------------------------------------------------------------------------------------
    def _wgrad(self, dy, x, dW, NB, cin_valid, dh, dw, dplane, in_stride):
        Cout = dy.shape[3]
        Ho, Wo = dy.shape[1], dy.shape[2]
        hh, ww = torch.meshgrid(torch.arange(Ho), torch.arange(Wo), indexing="ij")
        g = dW.view(Cout, len(dh), cin_valid)
        n_idx = torch.arange(NB)
        for t in range(len(dh)):
            v = _gather(x, n_idx + dplane[t] * NB, in_stride * hh + dh[t], in_stride * ww + dw[t])[..., :cin_valid]
            g[:, t, :] += torch.einsum("nhwo,nhwi->oi", dy, v)

    def conv_wgrad_bf16(self, dy, x, dW, NB, planes, cin_valid, dh, dw, dplane):
        self.calls.append("conv_wgrad_bf16")
        assert x.shape[0] == planes * NB
        self._wgrad(dy, x, dW, NB, cin_valid, dh, dw, dplane, 1)

    def conv_wgrad_bf16_strided(self, dy, x, dW, cin_valid, dh, dw, in_stride):
        self.calls.append("conv_wgrad_bf16_strided")
        self._wgrad(dy, x, dW, x.shape[0], cin_valid, dh, dw, [0] * len(dh), in_stride)

    def conv_wgrad_halo_bf16(self, dy, x, dW, cin_valid):
        self.calls.append("conv_wgrad_halo_bf16")
        dh = [d - 1 for d in range(3) for _ in range(3)]
        dw = [d - 1 for _ in range(3) for d in range(3)]
        self._wgrad(dy, x, dW, x.shape[0], cin_valid, dh, dw, [0] * 9, 1)

    # ---- GEMM-shaped helpers (stem path) -------------------------------------------------------------------------------
    def im2col_small(self, x, A, k, pad):
        self.calls.append("im2col_small")
        NB, H, W, C = x.shape
        Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
        hh, ww = torch.meshgrid(torch.arange(Ho), torch.arange(Wo), indexing="ij")
        A.zero_()
        A3 = A.view(NB, Ho, Wo, 64)
        for t in range(k * k):
            A3[..., t * C:(t + 1) * C] = _gather(x, torch.arange(NB), hh + t // k - pad, ww + t % k - pad)

    def gemm_bf16(self, A, B, out, bias, relu, accumulate, stats):
        self.calls.append("gemm_bf16")
        r = A @ B.t()
        if bias is not None:
            r = r + bias
        if relu:
            r = r.clamp_min(0)
        if accumulate:
            out += r
        else:
            out.copy_(r)

    def linear_wgrad_bf16(self, dy, x, dW):
        self.calls.append("linear_wgrad_bf16")
        dW += dy.t() @ x

    def channel_stats(self, x, st):
        self.calls.append("channel_stats")
        xf = x.reshape(-1, x.shape[-1])
        st[0] += xf.sum(0)
        st[1] += (xf * xf).sum(0)


@pytest.fixture
def fake(monkeypatch):
    ext = FakeExt()
    monkeypatch.setattr(nn, "_ext", lambda: ext)
    monkeypatch.setattr(nn, "_scratch", {})
    monkeypatch.setattr(nn, "_s2d_done", {})
    for flag in ("USE_STRIDED_TMA", "USE_IM2COL_STEM", "USE_HALO3"):
        monkeypatch.setattr(nn, flag, False)
    return ext


def _reference(x, w, bias, dy, s, p):
    xn = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    wn = w.permute(0, 3, 1, 2).clone().requires_grad_(True)
    b = bias.clone().requires_grad_(True)
    y = F.relu(F.conv2d(xn, wn, b, s, p))
    yl = F.conv2d(xn, wn, None, s, p)                                  # the backward entry points act on the pre-activation
    yl.backward(dy.permute(0, 3, 1, 2))
    return y.detach().permute(0, 2, 3, 1), xn.grad.permute(0, 2, 3, 1), wn.grad.permute(0, 2, 3, 1), dy.sum((0, 1, 2))


CASES = [  # B, H, W, Cin, Cout, k, stride, pad
    (2, 8, 8, 64, 128, 3, 2, 1), (2, 8, 8, 64, 128, 1, 2, 0), (3, 4, 8, 128, 64, 3, 2, 1), (2, 16, 8, 64, 64, 3, 1, 1), (2, 6, 6, 64, 128, 3, 1, 0),
    (2, 8, 8, 128, 128, 3, 1, 1), (2, 8, 8, 3, 64, 3, 1, 1), (2, 10, 10, 1, 32, 3, 1, 0), (2, 8, 8, 128, 64, 1, 1, 0),
]


@pytest.mark.parametrize("mode", ["default", "strided", "stem", "halo3"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", CASES)
def test_conv_wrappers_hand_the_kernels_the_right_problem(fake, monkeypatch, mode, B, H, W, Cin, Cout, k, s, p):
    monkeypatch.setattr(nn, "USE_STRIDED_TMA", mode == "strided")
    monkeypatch.setattr(nn, "USE_IM2COL_STEM", mode == "stem")
    monkeypatch.setattr(nn, "USE_HALO3", mode == "halo3")
    torch.manual_seed(B + H + Cin + k)
    x = torch.randn(B, H, W, Cin)
    w = torch.randn(Cout, k, k, Cin) / (k * k * Cin) ** 0.5
    bias = torch.randn(Cout) * 0.1
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    dy = torch.randn(B, Ho, Wo, Cout)
    y_ref, dx_ref, gw_ref, gb_ref = _reference(x, w, bias, dy, s, p)

    y = torch.full((B, Ho, Wo, Cout), 7.0)
    nn.conv2d_fwd_sm100(x, w, bias, y, s, p, True, None, tag=("t", mode), s2d_epoch=1)
    torch.testing.assert_close(y, y_ref, rtol=1e-4, atol=1e-4)

    gw, gb = torch.zeros(Cout, k, k, Cin), torch.zeros(Cout)
    nn.conv2d_wgrad_sm100(x, dy, gw, gb, s, p, tag=("t", mode))
    torch.testing.assert_close(gw, gw_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gb, gb_ref, rtol=1e-4, atol=1e-4)

    if Cin % 8 == 0 and Cout % 64 == 0:                                  # conv_supported(..., "dgrad")
        base = torch.randn(B, H, W, Cin)
        dx0, dx1 = torch.full_like(base, 3.0), base.clone()
        nn.conv2d_dgrad_sm100(dy, w, dx0, s, p, False)
        nn.conv2d_dgrad_sm100(dy, w, dx1, s, p, True)
        torch.testing.assert_close(dx0, dx_ref, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(dx1, dx_ref + base, rtol=1e-4, atol=1e-4)

    # the mode really took the path it names
    if mode == "strided" and s == 2:
        assert "conv_bf16_strided" in fake.calls and "conv_wgrad_bf16_strided" in fake.calls
        assert "space_to_depth" not in fake.calls and "depth_to_space" not in fake.calls
    if mode == "default" and s == 2:
        assert "space_to_depth" in fake.calls
    if mode == "stem" and Cin * k * k <= 64 and s == 1:
        assert fake.calls.count("im2col_small") == 1 and "gemm_bf16" in fake.calls and "linear_wgrad_bf16" in fake.calls
    if mode == "halo3" and nn._halo_ok(k, s, p, Cin, H, W):
        assert "conv3x3_halo3_bf16" in fake.calls and "conv3x3_halo_bf16" not in fake.calls


def test_shared_parity_copy_is_made_once_per_forward(fake):
    """The 3x3/s2 conv and the 1x1/s2 shortcut of a ResNet block read the same input: one space_to_depth per forward epoch."""
    x = torch.randn(2, 8, 8, 64)
    w3, w1 = torch.randn(128, 3, 3, 64), torch.randn(128, 1, 1, 64)
    y = torch.empty(2, 4, 4, 128)
    for epoch in (1, 2):
        nn.conv2d_fwd_sm100(x, w3, None, y, 2, 1, False, None, tag="a", s2d_epoch=epoch)
        nn.conv2d_fwd_sm100(x, w1, None, y, 2, 0, False, None, tag="b", s2d_epoch=epoch)
    assert fake.calls.count("space_to_depth") == 2
