"""The UNMODIFIED reference on the same B200, through its own code path (context for bench.py, not a bench line):
XPoint.forward (its CUDA scan + Triton CrossScan/Merge + cuDNN/cuBLAS, fp16 autocast as params.yaml: mixed_precision) ->
box_nms (configs/cipdp.yaml: cpu_nms true | GPU) -> per-sample nonzero / interpolate_descriptors -> cv2 BFMatcher, i.e. the loop of
xpoint/utils/evaluation.py:229-301 without the metrics.
    python profiles/ref_pipeline_bench.py [--pairs 8] [--height 512 --width 640] [--topk 4096] [--cpu-nms]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refstage as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--topk", type=int, default=4096)
    ap.add_argument("--preset", default="E")
    ap.add_argument("--cpu-nms", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    ns = R.load()
    try:
        ns.RC.cross_scan_fn(torch.randn(1, 4, 8, 8, device="cuda"))
        cross = "triton"
    except Exception as e:
        R.patch_cross_torch_path(ns)
        cross = f"torch ({type(e).__name__})"
    net = R.build_xpoint(ns, a.preset, mixed_precision=True, height=a.height, width=a.width).cuda()
    g = torch.Generator().manual_seed(0)
    o = torch.rand(a.pairs, 1, a.height, a.width, generator=g).cuda()
    t = torch.rand(a.pairs, 1, a.height, a.width, generator=g).cuda()
    U = ns.utils

    def step():
        tm = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            po, pt, _ = net({"optical": {"image": o}, "thermal": {"image": t}})
        torch.cuda.synchronize()
        tm["forward"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        nms_o = U.box_nms(po["prob"], 8, 0.015, keep_top_k=a.topk, on_cpu=a.cpu_nms)
        nms_t = U.box_nms(pt["prob"], 8, 0.015, keep_top_k=a.topk, on_cpu=a.cpu_nms)
        torch.cuda.synchronize()
        tm["nms"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        nm = 0
        t_match = 0.0
        for b in range(a.pairs):
            kp_o = torch.nonzero(nms_o[b].squeeze() > 0.015)
            kp_t = torch.nonzero(nms_t[b].squeeze() > 0.015)
            d_o = U.interpolate_descriptors(kp_o.to(po["desc"].device), po["desc"][b], a.height, a.width).cpu().numpy()
            d_t = U.interpolate_descriptors(kp_t.to(pt["desc"].device), pt["desc"][b], a.height, a.width).cpu().numpy()
            t1 = time.perf_counter()
            m = U.get_matches(d_o, d_t, "bfmatcher", False, crossCheck=True)
            t_match += time.perf_counter() - t1
            nm += len(m)
        tm["sample+match"] = time.perf_counter() - t0
        tm["match_only"] = t_match
        tm["matches"] = nm
        return tm

    step()
    rows = [step() for _ in range(a.steps)]
    tot = [r["forward"] + r["nms"] + r["sample+match"] for r in rows]
    best = min(range(len(rows)), key=lambda i: tot[i])
    out = {"what": "unmodified reference on this GPU (own CUDA scan, CrossScan path: %s)" % cross, "preset": a.preset,
           "pairs": a.pairs, "size": [a.height, a.width], "topk": a.topk, "cpu_nms": a.cpu_nms,
           "pairs_per_s": a.pairs / tot[best], "pairs_per_s_forward_only": a.pairs / rows[best]["forward"],
           "seconds": {k: round(v, 4) for k, v in rows[best].items()}, "torch_threads": torch.get_num_threads()}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
