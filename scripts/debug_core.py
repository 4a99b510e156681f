"""Per-direction check of xp_ss2d_core against the op-level scan (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200 import ss2d as S
from xpoint_b200.selective_scan import scan_forward

def run(H, W, D, N, dtype, B=2):
    g = torch.Generator().manual_seed(1)
    K, L, R = 4, H * W, 3
    x = torch.randn(B, D, H, W, generator=g)
    xx = torch.stack([x.reshape(B, D, L), x.transpose(2, 3).reshape(B, D, L)], 1).to(dtype).cuda()
    delta = (0.5 * torch.rand(B, K, D, L, generator=g) - 0.3).to(dtype).cuda()
    x_dbl0 = torch.randn(B, K, R + 2 * N, L, generator=g).to(dtype).cuda()
    A = (-0.5 * torch.rand(K * D, N, generator=g) - 0.01).cuda()
    Ds0 = torch.randn(K * D, generator=g).cuda()
    bias = (0.5 * torch.rand(K * D, generator=g)).cuda()
    for q in (0, 1, 2, 3, None):
        x_dbl, Ds = x_dbl0.clone(), Ds0.clone()
        if q is not None:
            for k in range(4):
                if k != q:
                    x_dbl[:, k, R + N:] = 0
                    Ds[k * D:(k + 1) * D] = 0
        Bs, Cs = x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:]
        ys, _ = scan_forward(xx.view(B, 2 * D, L), delta.view(B, K * D, L), A, Bs, Cs, Ds, None, bias, True, True,
                             u_group_div=2, reverse_group_mask=S.REVERSE_MASK)
        ys = ys.view(B, K, D, L)
        want = (ys[:, 0] + ys[:, 1]).view(B, D, H, W) + (ys[:, 2] + ys[:, 3]).view(B, D, W, H).transpose(2, 3)
        got = S.ss2d_core(xx, delta, A, Bs, Cs, Ds, bias, H, W, True)
        torch.cuda.synchronize()
        err = (got - want).abs()
        rel = (err.norm() / want.norm()).item()
        bad = (err > 1e-3 * want.abs().max()).float()
        msg = f"{H}x{W} D={D} N={N} {str(dtype)[6:]} dir={q}: rel_l2={rel:.2e}"
        if rel > 1e-4:
            # where are the bad elements? per batch/channel fraction, first bad rows/cols
            idx = torch.nonzero(bad)
            msg += f" bad={int(bad.sum())}/{bad.numel()} first={idx[:3].tolist()} last={idx[-2:].tolist()}"
            hb = bad.sum(dim=(0, 1, 3)); wb = bad.sum(dim=(0, 1, 2))
            msg += f" rows_bad={int((hb>0).sum())}/{H} cols_bad={int((wb>0).sum())}/{W}"
        print(msg, flush=True)

for dt in (torch.float32, torch.float16):
    run(8, 8, 2, 1, dt)
    run(16, 32, 2, 1, dt)       # exactly one / two full blocks
    run(32, 32, 2, 1, dt)
    run(128, 160, 2, 1, dt)
    run(32, 40, 12, 2, dt)
