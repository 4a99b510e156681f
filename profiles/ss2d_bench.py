"""Microbench of the copy-free SS2D kernels at the preset-E stage shapes (B = 128 images, fp16 autocast layout)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200 import ss2d

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = 128
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (H, W, D) in ((128, 160, 96), (64, 80, 192), (32, 40, 384), (16, 20, 768)):
    L = H * W
    if which in ("all", "merge"):
        ys = torch.randn(B, 4, D, L, device="cuda")
        g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
        ms = timeit(lambda: ss2d.ss2d_merge_norm(ys, H, W, g, b, None, 1e-5, out_dtype=torch.float16))
        nb = B * D * L * 18
        print(f"merge_norm {H}x{W}x{D}: {ms:.3f} ms  {nb/ms/1e6:.0f} GB/s", flush=True)
        del ys
    if which in ("all", "dwconv"):
        x = torch.randn(B, H, W, D, device="cuda", dtype=torch.float16)
        w = torch.randn(D, 1, 3, 3, device="cuda")
        ms = timeit(lambda: ss2d.ss2d_dwconv_pack(x, D, w, None, True))
        nb = B * D * L * 6
        print(f"dwconv_pack {H}x{W}x{D}: {ms:.3f} ms  {nb/ms/1e6:.0f} GB/s", flush=True)
        del x
