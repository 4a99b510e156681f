"""Launch xp_ss2d_core once per stage shape (for ncu captures).  python profiles/core_once.py [B] [stage]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200 import ss2d as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H, W, D, R = ((128, 160, 96, 6), (64, 80, 192, 12), (32, 40, 384, 24), (16, 20, 768, 48))[stage]
L, K, N, dt = H * W, 4, 1, torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
xx = torch.randn(B, 2, D, L, device="cuda", generator=g).to(dt)
delta = (0.5 * torch.rand(B, K, D, L, device="cuda", generator=g)).to(dt)
x_dbl = torch.randn(B, K, R + 2 * N, L, device="cuda", generator=g).to(dt)
A = -0.5 * torch.rand(K * D, N, device="cuda", generator=g)
Ds = torch.randn(K * D, device="cuda", generator=g)
bias = 0.5 * torch.rand(K * D, device="cuda", generator=g)
for _ in range(2):
    y = S.ss2d_core(xx, delta, A, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], Ds, bias, H, W, True)
torch.cuda.synchronize()
print("done", y.shape)
