"""fc1 + GELU of the VSSBlock MLP at the preset-E stage shapes (B = 128 images): tcgen05 fused kernel vs cuBLAS + torch GELU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200.cross_scan import linear_act
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (L, C) in ((20480, 96), (5120, 192), (1280, 384), (320, 768)):
    M, K, N = 128 * L, C, 4 * C
    x = torch.randn(M, K, device="cuda", dtype=torch.float16)
    w = torch.randn(N, K, device="cuda", dtype=torch.float16) / K ** 0.5
    b = torch.randn(N, device="cuda")
    bh = b.half()
    t_f = timeit(lambda: linear_act(x, w, b, gelu=True))
    t_p = timeit(lambda: linear_act(x, w, b, gelu=False))
    t_l = timeit(lambda: torch.nn.functional.linear(x, w, bh))
    t_g = timeit(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, bh)))
    nb = M * (K + N) * 2
    print(f"M={M} K={K} N={N}: fused+gelu {t_f:.3f} ms ({nb/t_f/1e6:.0f} GB/s)  fused plain {t_p:.3f}  cublas linear {t_l:.3f}  cublas+gelu {t_g:.3f}", flush=True)
