"""One launch each of linear_act(GELU), linear_res_ln and mlp_res_ln at the stage-0 shape (for ncu)."""
import torch
from xpoint_b200.cross_scan import linear_act, linear_res_ln, mlp_res_ln
M, C, dt = 128 * 128 * 80, 96, torch.float16
x = torch.randn(M, C, device="cuda").to(dt)
W1 = (torch.randn(4 * C, C, device="cuda") / C ** 0.5).to(dt); W2 = (torch.randn(C, 4 * C, device="cuda") / (4 * C) ** 0.5).to(dt)
b1 = torch.randn(4 * C, device="cuda"); b2 = torch.randn(C, device="cuda")
res = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
for _ in range(2):
    h = linear_act(x, W1, b1, gelu=True)
    linear_res_ln(h, W2, b2, res, g, b)
    mlp_res_ln(x, W1, b1, W2, b2, res, g, b)
torch.cuda.synchronize()
