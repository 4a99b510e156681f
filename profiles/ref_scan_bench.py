"""The reference's own CUDA kernel (selective_scan_cuda_oflex.fwd, compiled for sm_100a by oracle/build_ref.py) timed beside
xp_selective_scan_fwd on the same box, same tensors: "the kernel to beat" (SURVEY 8c / BASELINE.md 3).
    python profiles/ref_scan_bench.py [--json out.json]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refstage as R  # noqa: E402
import xpoint_b200 as X  # noqa: E402
from xpoint_b200.selective_scan import algorithmic_bytes  # noqa: E402

CASES = [("config2 fp32->fp32 N=16", 32, 768, 4, 16, 20480, torch.float32, True),
         ("config2 bf16->fp32 N=16", 32, 768, 4, 16, 20480, torch.bfloat16, True),
         ("config2 bf16->bf16 N=16", 32, 768, 4, 16, 20480, torch.bfloat16, False),
         ("xpoint-actual fp32 N=1 (B=64,K*D=384)", 64, 384, 4, 1, 20480, torch.float32, True),
         ("xpoint-actual fp16->fp32 N=1 (B=128,K*D=384)", 128, 384, 4, 1, 20480, torch.float16, True),
         ("E stage1 fp16 (B=128,K*D=768,L=5120)", 128, 768, 4, 1, 5120, torch.float16, True),
         ("E stage2 fp16 (B=128,K*D=1536,L=1280)", 128, 1536, 4, 1, 1280, torch.float16, True),
         ("E stage3 fp16 (B=128,K*D=3072,L=320)", 128, 3072, 4, 1, 320, torch.float16, True)]


def timeit(fn, warm=10, iters=30):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ext = R.load_ext()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    for name, Bt, KD, K, N, L, dt, oflex in CASES:
        g = torch.Generator(device="cuda").manual_seed(0)
        u = torch.randn(Bt, KD, L, generator=g, device="cuda").to(dt)
        dl = (0.5 * torch.rand(Bt, KD, L, generator=g, device="cuda")).to(dt)
        A = -0.5 * torch.rand(KD, N, generator=g, device="cuda")
        Bm = torch.randn(Bt, K, N, L, generator=g, device="cuda").to(dt)
        Cm = torch.randn(Bt, K, N, L, generator=g, device="cuda").to(dt)
        D = torch.randn(KD, generator=g, device="cuda")
        bias = 0.5 * torch.rand(KD, generator=g, device="cuda")
        ms_ref = timeit(lambda: ext.fwd(u, dl, A, Bm, Cm, D, bias, True, 1, oflex))
        ms_ours = timeit(lambda: X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, oflex))
        nbytes = algorithmic_bytes(Bt, KD, K, N, L, u.element_size(), 4 if oflex else u.element_size())
        row = {"case": name, "ref_kernel_ms": round(ms_ref, 4), "ours_ms": round(ms_ours, 4), "speedup": round(ms_ref / ms_ours, 2),
               "ref_frac_of_hbm_peak": round(nbytes / ms_ref / 1e6 / peak, 4), "ours_frac_of_hbm_peak": round(nbytes / ms_ours / 1e6 / peak, 4)}
        print(json.dumps(row), flush=True)
        rows.append(row)
        del u, dl, Bm, Cm
        torch.cuda.empty_cache()
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        json.dump(rows, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
