"""Launch each selective-scan microbench shape a few times (for `ncu -k regex:scan_ ...` captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {"n16_fp32": (32, 768, 4, 16, 20480, torch.float32), "n1_fp32": (64, 384, 4, 1, 20480, torch.float32),
         "n1_fp16": (128, 384, 4, 1, 20480, torch.float16), "n16_bf16": (32, 768, 4, 16, 20480, torch.bfloat16)}
for name, (Bt, KD, K, N, L, dt) in cases.items():
    if which not in ("all", name):
        continue
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randn(Bt, KD, L, generator=g, device="cuda").to(dt)
    dl = (0.5 * torch.rand(Bt, KD, L, generator=g, device="cuda")).to(dt)
    A = -0.5 * torch.rand(KD, N, generator=g, device="cuda")
    Bm = torch.randn(Bt, K, N, L, generator=g, device="cuda").to(dt)
    Cm = torch.randn(Bt, K, N, L, generator=g, device="cuda").to(dt)
    D = torch.randn(KD, generator=g, device="cuda")
    bias = 0.5 * torch.rand(KD, generator=g, device="cuda")
    for _ in range(3):
        y = X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, True)
    torch.cuda.synchronize()
    print(name, "done")
