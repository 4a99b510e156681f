"""One stage-0 launch shape of the fused-dt_proj scan (B=128 images, K*D=384, N=1, L=20480, R=6, fp16 -> fp32) for
`ncu -k regex:scan_lanes --set full` captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200.selective_scan import scan_forward
B, K, N, D, L, R, dt = 128, 4, 1, 96, 20480, 6, torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
xx = torch.randn(B, 2, D, L, generator=g, device="cuda").to(dt)
x_dbl = (0.3 * torch.randn(B, K, R + 2 * N, L, generator=g, device="cuda")).to(dt)
w16 = (torch.randn(K * D, R, generator=g, device="cuda") * R ** -0.5).to(dt)
A = -0.5 * torch.rand(K * D, N, generator=g, device="cuda")
Ds = torch.randn(K * D, generator=g, device="cuda"); bias = 0.5 * torch.rand(K * D, generator=g, device="cuda")
for _ in range(3):
    y = scan_forward(xx.view(B, 2 * D, L), x_dbl[:, :, :R], A, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], Ds, None, bias,
                     True, True, u_group_div=2, reverse_group_mask=0b1010, dt_weight=w16)[0]
torch.cuda.synchronize()
print("done")
