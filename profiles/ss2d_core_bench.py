"""SS2D core at the preset-E stage shapes (B = 128 images, fp16): the fused core (xp_ss2d_core + xp_ss2d_plane_norm) against the
op-level path it replaces (xp_selective_scan_fwd writing four fp32 planes + xp_ss2d_merge_norm).  Inputs as the model produces
them (xx, materialised delta, B/C views of x_dbl).   python profiles/ss2d_core_bench.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200 import ss2d as S
from xpoint_b200.selective_scan import scan_forward

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dt = torch.float16
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (H, W, D, R) in ((128, 160, 96, 6), (64, 80, 192, 12), (32, 40, 384, 24), (16, 20, 768, 48)):
    L, K, N = H * W, 4, 1
    g = torch.Generator(device="cuda").manual_seed(0)
    xx = torch.randn(B, 2, D, L, device="cuda", generator=g).to(dt)
    delta = (0.5 * torch.rand(B, K, D, L, device="cuda", generator=g)).to(dt)
    x_dbl = torch.randn(B, K, R + 2 * N, L, device="cuda", generator=g).to(dt)
    Bs, Cs = x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:]
    A = -0.5 * torch.rand(K * D, N, device="cuda", generator=g)
    Ds = torch.randn(K * D, device="cuda", generator=g)
    bias = 0.5 * torch.rand(K * D, device="cuda", generator=g)
    gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    ch = S.core_channels(D, N, H, W, dt)
    ys = None
    def old_scan():
        global ys
        ys, _ = scan_forward(xx.view(B, 2 * D, L), delta.view(B, K * D, L), A, Bs, Cs, Ds, None, bias, True, True,
                             u_group_div=2, reverse_group_mask=S.REVERSE_MASK)
    t_scan = timeit(old_scan)
    t_merge = timeit(lambda: S.ss2d_merge_norm(ys.view(B, K, D, L), H, W, gam, bet, None, 1e-5, out_dtype=dt))
    line = f"{H}x{W}x{D}: scan {t_scan:.3f} + merge_norm {t_merge:.3f} = {t_scan + t_merge:.3f} ms"
    if ch > 0:
        y = None
        def core():
            global y
            y = S.ss2d_core(xx, delta, A, Bs, Cs, Ds, bias, H, W, True)
        t_core = timeit(core)
        t_norm = timeit(lambda: S.ss2d_plane_norm(y, gam, bet, None, 1e-5, out_dtype=dt))
        alg = B * L * (D * 2 * 2 + 4 * D * 2 + 8 * N * 2 + D * 4)     # xx + delta + B/C + merged y
        line += f" | core (CH={ch}) {t_core:.3f} + plane_norm {t_norm:.3f} = {t_core + t_norm:.3f} ms  (core {alg / t_core / 1e6:.0f} GB/s of its {alg / 1e9:.2f} GB)"
    print(line, flush=True)
    del xx, delta, x_dbl, ys
    torch.cuda.empty_cache()
