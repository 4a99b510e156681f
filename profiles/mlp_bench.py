"""Fused Mlp branch (xp_mlp_res_ln) against linear_act + linear_res_ln at the stage-0 / stage-1 shapes of configs[1].
Run: python profiles/mlp_bench.py"""
import torch
from xpoint_b200.cross_scan import linear_act, linear_res_ln, mlp_res_ln

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for M, C in [(128 * 128 * 80, 96), (128 * 64 * 40, 192)]:
    dt = torch.float16
    x = torch.randn(M, C, device="cuda").to(dt)
    W1 = (torch.randn(4 * C, C, device="cuda") / C ** 0.5).to(dt); W2 = (torch.randn(C, 4 * C, device="cuda") / (4 * C) ** 0.5).to(dt)
    b1 = torch.randn(4 * C, device="cuda"); b2 = torch.randn(C, device="cuda")
    res = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    two = t(lambda: linear_res_ln(linear_act(x, W1, b1, gelu=True), W2, b2, res, g, b))
    one = t(lambda: mlp_res_ln(x, W1, b1, W2, b2, res, g, b))
    nosum = t(lambda: mlp_res_ln(x, W1, b1, W2, b2, res, g, b, want_sum=False))
    noln = t(lambda: mlp_res_ln(x, W1, b1, W2, b2, res, None, None, want_sum=False))
    s1, y1 = mlp_res_ln(x, W1, b1, W2, b2, res, g, b)
    s2, y2 = linear_res_ln(linear_act(x, W1, b1, gelu=True), W2, b2, res, g, b)
    alg = M * C * (2 + 4 + 4 + 2)
    print(f"M={M} C={C}: fc1+gelu , fc2+res+ln = {two:.3f} ms ; fused = {one:.3f} ms ({alg / one / 1e6:.0f} GB/s algorithmic) ; "
          f"without the fp32 sum {nosum:.3f} ms, 16-bit sum only {noln:.3f} ms ; max|ds|={float((s1 - s2).abs().max()):.2e}")
