"""Fused dt_proj (xp_scan_args.dt_weight) vs dt_proj kernel/GEMM + scan at the stage shapes of preset E, 64 pairs of
512x640 (B = 128 images): time per call, CUDA events, 20 warm-up + 50 timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X
from xpoint_b200 import ss2d as S
from xpoint_b200.selective_scan import scan_forward, algorithmic_bytes

def timeit(f, n=50, w=20):
    for _ in range(w): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dt = torch.float16
for (D, L, R) in ((96, 20480, 6), (192, 5120, 12), (384, 1280, 24)):
    B, K, N = 128, 4, 1
    g = torch.Generator(device="cuda").manual_seed(0)
    xx = torch.randn(B, 2, D, L, generator=g, device="cuda").to(dt)
    x_dbl = (0.3 * torch.randn(B, K, R + 2 * N, L, generator=g, device="cuda")).to(dt)
    wdt = (torch.randn(K, D, R, generator=g, device="cuda") * R ** -0.5)
    A = -0.5 * torch.rand(K * D, N, generator=g, device="cuda")
    Ds = torch.randn(K * D, generator=g, device="cuda"); bias = 0.5 * torch.rand(K * D, generator=g, device="cuda")
    w16 = wdt.reshape(K * D, R).to(dt).contiguous()
    wb = wdt.to(dt).reshape(1, 2, 2, D, R).expand(B, -1, -1, -1, -1).contiguous()
    def dtproj():
        if S.dt_proj_supported(R, L, dt, B * K):
            return S.ss2d_dt_proj(x_dbl[:, :, :R], wdt)
        return torch.matmul(wb, x_dbl.view(B, 2, 2, R + 2 * N, L)[:, :, :, :R])
    def plain():
        d = dtproj()
        return scan_forward(xx.view(B, 2 * D, L), d.view(B, K * D, L), A, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], Ds, None, bias,
                            True, True, u_group_div=2, reverse_group_mask=0b1010)[0]
    d0 = dtproj()
    def scan_only():
        return scan_forward(xx.view(B, 2 * D, L), d0.view(B, K * D, L), A, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], Ds, None, bias,
                            True, True, u_group_div=2, reverse_group_mask=0b1010)[0]
    def fused():
        return scan_forward(xx.view(B, 2 * D, L), x_dbl[:, :, :R], A, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], Ds, None, bias,
                            True, True, u_group_div=2, reverse_group_mask=0b1010, dt_weight=w16)[0]
    a, b = plain(), fused()
    err = float((a - b).norm() / a.norm())
    tp, ts, tf = timeit(plain), timeit(scan_only), timeit(fused)
    gb = algorithmic_bytes(B, K * D, K, N, L, 2, 4, R) / 1e9
    print(f"D={D} L={L} R={R}: dt_proj+scan {tp:.3f} ms (scan alone {ts:.3f}), fused {tf:.3f} ms = {gb / tf * 1e3:.0f} GB/s "
          f"on {gb:.2f} GB; rel diff {err:.2e}")
    del xx, x_dbl, wb, d0, a, b
    torch.cuda.empty_cache()
