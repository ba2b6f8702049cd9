"""Launch the hot kernels of the DEFAULT code path on representative shapes (for `ncu --set full -k regex:...`): conv forward / data
gradient / weight gradient on the ResNet-18 layer shapes, BatchNorm forward / backward, batch assembly + stem GEMM, the fused server step
and the optimizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200 import ops  # noqa: E402
from rlr_b200.data import make_synthetic  # noqa: E402

DEV, BF = "cuda:0", torch.bfloat16
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def conv_case(B, H, C, Cout, tag):
    x = torch.randn(B, H, H, C, device=DEV).to(BF)
    w = (torch.randn(Cout, 3, 3, C, device=DEV) * 0.05).to(BF)
    y = torch.empty(B, H, H, Cout, device=DEV, dtype=BF)
    dy = torch.randn(B, H, H, Cout, device=DEV).to(BF)
    gw = torch.zeros(Cout, 3, 3, C, device=DEV)
    dx = torch.empty_like(x)
    for _ in range(2):
        if which in ("all", "fwd"):
            ops.conv2d_fwd_sm100(x, w, None, y, 1, 1, False, None, tag=tag)
        if which in ("all", "wgrad"):
            ops.conv2d_wgrad_sm100(x, dy, gw, None, 1, 1, tag=tag)
        if which in ("all", "dgrad"):
            ops.conv2d_dgrad_sm100(dy, w, dx, 1, 1, False)
    torch.cuda.synchronize()


def bn_case(M, C):
    x = torch.randn(M, 1, 1, C, device=DEV).to(BF); r = torch.randn_like(x); y = torch.empty_like(x)
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    rm, rv, mr = torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros(2, C, device=DEV)
    dy = torch.randn_like(x); dx = torch.empty_like(x); dres = torch.empty_like(x)
    dg, db, ds = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(2, C, device=DEV)
    for _ in range(2):
        ops.bn_fwd(x, y, r, g, b, rm, rv, None, mr, M, 1e-5, 0.1, True, True, "sm100")
        ops.bn_bwd(dy, y, x, g, mr, ds, dx, dres, dg, db, True, "sm100")
        ops.bn_fwd(x, y, None, g, b, rm, rv, None, mr, M, 1e-5, 0.1, True, True, "sm100")
        ops.bn_bwd(dy, y, x, g, mr, ds, dx, None, dg, db, True, "sm100", beta=b)
    torch.cuda.synchronize()


if which in ("all", "bn"):
    bn_case(256 * 32 * 32, 64)
    bn_case(256 * 8 * 8, 256)
if which in ("all", "fwd", "wgrad", "dgrad"):
    conv_case(256, 32, 64, 64, "l1")
    conv_case(256, 16, 128, 128, "l2")
    conv_case(256, 8, 256, 256, "l3")
if which in ("all", "stem"):
    tr, _ = make_synthetic("cifar10", 1024)
    d = tr.to(DEV)
    perm = torch.randperm(1024, device=DEV)
    A = torch.zeros(256 * 32 * 32, 64, device=DEV, dtype=BF)
    w = (torch.randn(64, 27, device=DEV) * 0.1).to(BF)
    y = torch.empty(256 * 32 * 32, 64, device=DEV, dtype=BF)
    for _ in range(2):
        ops.gather_im2col(d.data, perm, d.meta.mean, d.meta.std, 3, 1, A, batch=256)
        ops.ext().stem_gemm_bf16(A, w, y, None, False, None, 0, 0, 0, None)
    torch.cuda.synchronize()
if which in ("all", "agg"):
    n = 11190272
    g = torch.randn(n, device=DEV)
    ws = [g + 0.01 * torch.randn(n, device=DEV) for _ in range(8)]
    out = torch.empty_like(g); sh = torch.empty(n, device=DEV, dtype=BF)
    for mode in ("avg", "comed"):
        for _ in range(2):
            ops.fused_aggregate(g, ws, [1.0] * 8, mode, 4, 1.0, n_vote=n - 12288, out=out, out_bf16=sh)
    ws40 = ws + [g + 0.01 * torch.randn(n, device=DEV) for _ in range(32)]
    ops.fused_aggregate(g, ws40, [1.0] * 40, "comed", 10, 1.0, n_vote=n - 12288, out=out, out_bf16=sh)
    opt = ops.FlatSGD(n, DEV, 0.1, 0.9, 10.0, 0.0, n_pgd=n - 12288)
    m = torch.zeros_like(g)
    for _ in range(2):
        opt.step(out, ws[0], m, w_bf16=sh)
    torch.cuda.synchronize()
print("ok")
