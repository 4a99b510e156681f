"""Selective-scan microbench only (BASELINE configs[1] shapes): prints ms / GB/s / fraction of measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X
from xpoint_b200.selective_scan import algorithmic_bytes

peak = 6543.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
cases = [("n16 fp32", 32, 768, 4, 16, 20480, torch.float32, True), ("n16 bf16->f32", 32, 768, 4, 16, 20480, torch.bfloat16, True),
         ("n1 fp32 B64", 64, 384, 4, 1, 20480, torch.float32, True), ("n1 fp16->f32 B128", 128, 384, 4, 1, 20480, torch.float16, True),
         ("n1 fp16 s1", 128, 768, 4, 1, 5120, torch.float16, True), ("n1 fp16 s2", 128, 1536, 4, 1, 1280, torch.float16, True),
         ("n1 fp16 s3", 128, 3072, 4, 1, 320, torch.float16, True), ("n1 fp32 B1", 1, 384, 4, 1, 20480, torch.float32, True)]
sel = sys.argv[1:] 
for name, Bt, KD, K, N, L, dt, oflex in cases:
    if sel and not any(s in name for s in sel):
        continue
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randn(Bt, KD, L, generator=g, device="cuda").to(dt)
    dl = (0.5 * torch.rand(Bt, KD, L, generator=g, device="cuda")).to(dt)
    A = -0.5 * torch.rand(KD, N, generator=g, device="cuda")
    Bm = torch.randn(Bt, K, N, L, generator=g, device="cuda").to(dt)
    Cm = torch.randn(Bt, K, N, L, generator=g, device="cuda").to(dt)
    D = torch.randn(KD, generator=g, device="cuda")
    bias = 0.5 * torch.rand(KD, generator=g, device="cuda")
    for _ in range(10):
        y = X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, oflex)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30):
        y = X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, oflex)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    nb = algorithmic_bytes(Bt, KD, K, N, L, u.element_size(), y.element_size())
    print(f"{name:22s} {ms:8.4f} ms  {nb/ms/1e6:8.1f} GB/s  {nb/ms/1e6/peak:6.3f} of measured HBM peak", flush=True)
    del u, dl, Bm, Cm, y
