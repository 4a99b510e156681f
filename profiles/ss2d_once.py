"""Launch the copy-free SS2D kernels once at the stage-0 shape of preset E (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200 import ss2d
B, H, W, D = 128, 128, 160, 96
L = H * W
ys = torch.randn(B, 4, D, L, device="cuda")
g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
x = torch.randn(B, H, W, D, device="cuda", dtype=torch.float16)
w = torch.randn(D, 1, 3, 3, device="cuda")
for _ in range(2):
    ss2d.ss2d_merge_norm(ys, H, W, g, b, None, 1e-5, out_dtype=torch.float16)
    ss2d.ss2d_dwconv_pack(x, D, w, None, True)
torch.cuda.synchronize()
