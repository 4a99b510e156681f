#!/usr/bin/env python
"""bench.py -- XPoint pair-inference throughput on B200 + selective-scan roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--preset E|V] [--batch 64]
                    [--height 512 --width 640] [--topk 4096] [--dtype fp16|fp32] [--no-microbench] [--no-cpu-baseline]

A "step" is one pass of the hot path over one batch of synthetic image pairs (BASELINE.json configs[2]/[3]:
full XPoint inference at 512x640, batch 64 pairs per GPU: VMamba encoder + heads + NMS top-4096 + descriptor
sampling + mutual-NN matching), random-init weights (seed 0, identical on every rank), synthetic images
(seed = rank).  Pairs are batch-sharded over ranks with no collective on the data path; the only distributed
calls are the barrier and the MAX-reduction of the device time after the timed region.

One JSON line is printed by rank 0:
  value      pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        pairs/s through the public API (PairPipeline) from pinned HOST images, H2D + D2H inside the region
  roofline   selective-scan launches inside the timed region: algorithmic bytes / CUDA-event time vs measured HBM peak
  scan_microbench  BASELINE configs[1] (B=32, K*D=768, N=16, L=20480; fp32 and bf16) and the XPoint-actual N=1 shape
  cpu_baseline     the reference's own CPU path on the host cores on a bounded sample (kind "reference": the UNMODIFIED
                   reference staged by oracle/build_ref.py -- XPoint.forward + box_nms(on_cpu) + interpolate_descriptors +
                   cv2 BFMatcher; kind "port": the oracle restatement when the reference is not staged), N=1 only; the GPU
                   tail is checked bit-exactly against that run's keypoints / matches (parity_check)
  ss2d_core        SURVEY 8d secondary figure: one stage-0 SS2D core (dwconv_pack .. out_norm) against its 1.17 GB
  other_configs    BASELINE configs[4] (1024x1280, top-16384), preset V, single-pair latency (benchmark.py:151-164)
--impl reference times the CPU path as the reference arm (rank 0 only; one pair per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image pairs/sec (512x640 XPoint inference)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="E", choices=["E", "V"])
    ap.add_argument("--batch", type=int, default=64, help="image pairs per GPU per step")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--topk", type=int, default=4096)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32-match", action="store_true", help="use the exact CUDA-core matcher instead of tcgen05")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--ref-port", action="store_true", help="CPU arm: time the oracle port even when the reference is staged")
    ap.add_argument("--no-extras", action="store_true", help="skip the other_configs / ss2d_core measurements")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pair_run(args, n_pairs=1, threads=None):
    """kind "port": the oracle restatement of the whole path (oracle/model.py + oracle/xp_oracle.c) on the host cores."""
    import torch
    from oracle import model as M
    from oracle import oracle as O
    import xpoint_b200 as X
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    O.set_num_threads(threads)
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": True, "use_attention": {"preset": args.preset}}).eval()   # weights only (CPU tensors)
    sd = {k: v.float() for k, v in net.state_dict().items()}
    kw = X.PRESETS[args.preset]
    cfg = dict(depths=kw["depths"], downsample_version=kw["downsample_version"], patchembed_version=kw["patchembed_version"],
               forward_type=kw["forward_type"], patch_size=kw["patch_size"])
    g = torch.Generator().manual_seed(0)
    opt = torch.rand(n_pairs, 1, args.height, args.width, generator=g)
    thr = torch.rand(n_pairs, 1, args.height, args.width, generator=g)

    def step():
        t0 = time.perf_counter()
        res = M.pair_inference(opt, thr, sd, cfg, topk=args.topk)
        return time.perf_counter() - t0, res
    return step, threads


def reference_staged():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refstage
    return refstage if refstage.available(need_ext=False) else None


def cpu_reference_run(args, n_pairs=1, threads=None):
    """kind "reference": the UNMODIFIED reference (baseline/_ref, staged by oracle/build_ref.py) on the host cores, driven
    like its evaluation loop (xpoint/utils/evaluation.py:229-301 with configs/cipdp.yaml: nms 8, detection_threshold 0.015,
    cpu_nms true): XPoint.forward (fp32, selective_scan_torch -- csms6s.py:25-68 -- because no CUDA extension is imported
    in this process before the package) -> box_nms(on_cpu=True, keep_top_k) -> nonzero -> interpolate_descriptors ->
    get_matches('bfmatcher', crossCheck=True).  The one patch is SURVEY 0.8's CPU CrossScan rebind."""
    import torch
    R = reference_staged()
    ns = R.load(with_ext=False)
    assert not ns.RS.WITH_SELECTIVESCAN_OFLEX, "the CPU arm must run the reference's torch path"
    R.patch_cross_torch_path(ns)
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    net = R.build_xpoint(ns, args.preset, mixed_precision=False, height=args.height, width=args.width, seed=0)
    g = torch.Generator().manual_seed(0)
    opt = torch.rand(n_pairs, 1, args.height, args.width, generator=g)
    thr = torch.rand(n_pairs, 1, args.height, args.width, generator=g)
    U = ns.utils

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            po, pt, _ = net({"optical": {"image": opt}, "thermal": {"image": thr}})
        nms_o = U.box_nms(po["prob"], 8, 0.015, keep_top_k=args.topk, on_cpu=True)
        nms_t = U.box_nms(pt["prob"], 8, 0.015, keep_top_k=args.topk, on_cpu=True)
        out = []
        for b in range(n_pairs):
            kp_o = torch.nonzero(nms_o[b].squeeze() > 0.015)
            kp_t = torch.nonzero(nms_t[b].squeeze() > 0.015)
            d_o = U.interpolate_descriptors(kp_o, po["desc"][b], args.height, args.width).numpy()
            d_t = U.interpolate_descriptors(kp_t, pt["desc"][b], args.height, args.width).numpy()
            m = U.get_matches(d_o, d_t, "bfmatcher", False, crossCheck=True) if len(kp_o) and len(kp_t) else []
            out.append(dict(kp_o=kp_o, kp_t=kp_t, pairs=[(x.queryIdx, x.trainIdx) for x in m], d_o=d_o, d_t=d_t))
        dt = time.perf_counter() - t0
        return dt, dict(prob_o=po["prob"], prob_t=pt["prob"], desc_o=po["desc"], desc_t=pt["desc"], tail=out,
                        state_dict=net.state_dict())
    return step, threads


def cpu_arm(args, n_pairs):
    """(step, threads, kind, description) of the CPU baseline: the staged reference when present, else the oracle port."""
    if reference_staged() is not None and not args.ref_port:
        step, threads = cpu_reference_run(args, n_pairs)
        return step, threads, "reference", ("unmodified reference (baseline/_ref): XPoint.forward fp32 via selective_scan_torch + "
                                            "box_nms(on_cpu) + interpolate_descriptors + cv2.BFMatcher(crossCheck)")
    step, threads = cpu_pair_run(args, n_pairs)
    return step, threads, "port", "oracle/model.py + xp_oracle.c (torch CPU for dense layers): the reference is not staged here"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    NP = 1 if (reference_staged() is not None and not args.ref_port) else 8      # pairs per step: a bounded sample
    step, threads, kind, what = cpu_arm(args, NP)
    for _ in range(args.warmup):
        step()
    times = [step()[0] for _ in range(args.steps)]
    per = sum(times) / len(times)
    v = NP / per
    line = {"metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"XPoint preset {args.preset} pair inference {args.height}x{args.width}, top-{args.topk} keypoints, "
                                   f"MNN matching on the host cores; {what}", "pairs_per_step": NP},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind,
                             "sample": f"{NP} pair(s) per step x {args.steps} steps; {what}"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def scan_microbench(peak):
    """BASELINE configs[1]: SS2D selective-scan microbench, 30 warm-up + 100 timed launches."""
    import torch
    import xpoint_b200 as X
    from xpoint_b200.selective_scan import algorithmic_bytes
    out = []
    dev = "cuda"
    # the reference's own CUDA kernel (selective_scan_cuda_oflex.fwd) compiled for sm_100a by oracle/build_ref.py, when staged:
    # "the kernel to beat on the same box" (SURVEY 8c); timed on the same tensors, never on the product path
    ref_ext = None
    try:
        R = reference_staged()
        if R is not None and R.ext_path():
            ref_ext = R.load_ext()
    except Exception:
        ref_ext = None
    cases = [("config2 fp32->fp32 N=16", 32, 768, 4, 16, 20480, torch.float32, True),
             ("config2 bf16->fp32 N=16", 32, 768, 4, 16, 20480, torch.bfloat16, True),
             ("config2 bf16->bf16 N=16", 32, 768, 4, 16, 20480, torch.bfloat16, False),
             ("xpoint-actual fp32 N=1 (B=64,K*D=384)", 64, 384, 4, 1, 20480, torch.float32, True),
             ("xpoint-actual fp16->fp32 N=1 (B=128,K*D=384)", 128, 384, 4, 1, 20480, torch.float16, True)]
    for name, Bt, KD, K, N, L, dt, oflex in cases:
        g = torch.Generator(device=dev).manual_seed(0)
        u = torch.randn(Bt, KD, L, generator=g, device=dev).to(dt)
        dl = (0.5 * torch.rand(Bt, KD, L, generator=g, device=dev)).to(dt)
        A = -0.5 * torch.rand(KD, N, generator=g, device=dev)
        Bm = torch.randn(Bt, K, N, L, generator=g, device=dev).to(dt)
        Cm = torch.randn(Bt, K, N, L, generator=g, device=dev).to(dt)
        D = torch.randn(KD, generator=g, device=dev)
        bias = 0.5 * torch.rand(KD, generator=g, device=dev)
        for _ in range(30):
            y = X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, oflex)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            y = X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, oflex)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        nbytes = algorithmic_bytes(Bt, KD, K, N, L, u.element_size(), y.element_size())
        gbs = nbytes / ms / 1e6
        row = {"case": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3), "achieved_GBs": round(gbs, 1),
               "frac_of_hbm_peak": round(gbs / peak, 4)}
        if ref_ext is not None:
            for _ in range(5):
                ref_ext.fwd(u, dl, A, Bm, Cm, D, bias, True, 1, oflex)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                ref_ext.fwd(u, dl, A, Bm, Cm, D, bias, True, 1, oflex)
            e1.record()
            torch.cuda.synchronize()
            ref_ms = e0.elapsed_time(e1) / 20
            row["ref_kernel_ms"] = round(ref_ms, 4)
            row["ref_frac_of_hbm_peak"] = round(nbytes / ref_ms / 1e6 / peak, 4)
            row["speedup_vs_ref_kernel"] = round(ref_ms / ms, 2)
        out.append(row)
        del u, dl, Bm, Cm, y
        torch.cuda.empty_cache()
    return out


def ss2d_core_figure(peak, B=128):
    """SURVEY 8d secondary figure: ONE stage-0 SS2D core of preset E (x after in_proj -> y before out_proj: depth-wise conv +
    SiLU + CrossScan -> x_proj / dt_proj -> selective scan -> CrossMerge -> out_norm) at B images of 128x160 tokens, fp16.
    Algorithmic bytes B*L*[(D + K(R+2N)) s_in + D s_out] = 448 B per token; times are CUDA events around each kernel."""
    import torch
    import xpoint_b200 as X
    from xpoint_b200 import ss2d as S
    from xpoint_b200.selective_scan import scan_forward
    H, W, C = 128, 160, 96
    L = H * W
    torch.manual_seed(0)
    m = X.SS2D(d_model=C, d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False).cuda().eval().half()
    D, K, N, R = m.d_inner, 4, 1, m.dt_rank
    x = torch.randn(B, H, W, D, device="cuda", dtype=torch.float16)
    w = m._fused_weights(B, torch.float16)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = {}

    def timed(name, fn, n=10):
        for _ in range(3):
            r = fn()
        a, b = ev(), ev()
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        times[name] = a.elapsed_time(b) / n
        return r
    with torch.no_grad():
        xx = timed("dwconv_pack", lambda: S.ss2d_dwconv_pack(x, D, w["conv_w"], w["conv_b"], True))
        x_dbl = timed("x_proj (cuBLAS)", lambda: torch.matmul(w["wx"], xx)).view(B, K, R + 2 * N, L)
        dts = timed("dt_proj", lambda: S.ss2d_dt_proj(x_dbl[:, :, :R], w["wdt32"]))
        ys = timed("selective_scan (4 fp32 planes)", lambda: scan_forward(
            xx.view(B, 2 * D, L), dts.view(B, K * D, L), w["A"], x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], w["Ds"], None,
            w["dt_bias"], True, True, u_group_div=2, reverse_group_mask=S.REVERSE_MASK)[0])
        timed("merge_norm", lambda: S.ss2d_merge_norm(ys.view(B, K, D, L), H, W, w["norm_w"], w["norm_b"], None, 1e-5,
                                                      out_dtype=torch.float16))
        fused = None
        if S.core_channels(D, N, H, W, torch.float16) > 0:
            y = timed("fused core (xp_ss2d_core, opt-in)", lambda: S.ss2d_core(
                xx, dts.view(B, K, D, L), w["A"], x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], w["Ds"], w["dt_bias"], H, W, True))
            timed("plane_norm (opt-in)", lambda: S.ss2d_plane_norm(y, w["norm_w"], w["norm_b"], None, 1e-5, out_dtype=torch.float16))
            fused = times["dwconv_pack"] + times["x_proj (cuBLAS)"] + times["dt_proj"] + times["fused core (xp_ss2d_core, opt-in)"] \
                + times["plane_norm (opt-in)"]
    default = sum(times[k] for k in ("dwconv_pack", "x_proj (cuBLAS)", "dt_proj", "selective_scan (4 fp32 planes)", "merge_norm"))
    alg = B * L * ((D + K * (R + 2 * N)) * 2 + D * 2)
    es = 2
    # DRAM bytes each kernel must move (inputs read once + outputs written once), from the tensor shapes
    moved = {"dwconv_pack": B * L * D * es * 3, "x_proj (cuBLAS)": B * L * (2 * D + K * (R + 2 * N)) * es,
             "dt_proj": B * L * (K * R + K * D) * es, "selective_scan (4 fp32 planes)": B * L * (4 * D * es + K * D * es + 2 * K * N * es + K * D * 4),
             "merge_norm": B * L * (K * D * 4 + D * es),
             "fused core (xp_ss2d_core, opt-in)": B * L * (2 * D * es + K * D * es + 2 * K * N * es + D * 4),
             "plane_norm (opt-in)": B * L * (D * 4 + D * es)}
    out = {"shape": f"B{B} x 128x160 tokens, d_inner 96, N 1, K 4, R {R}, fp16", "algorithmic_GB": round(alg / 1e9, 3),
           "default_path_ms": round(default, 3), "default_effective_GBs": round(alg / default / 1e6, 1),
           "default_frac_of_hbm_peak": round(alg / default / 1e6 / peak, 4),
           "default_tensor_traffic_GB": round(sum(moved[k] for k in list(moved)[:5]) / 1e9, 2),
           "ncu_dram_GB_r1": 17.2, "kernels_ms": {k: round(v, 3) for k, v in times.items()},
           "kernel_tensor_traffic_GB": {k: round(v / 1e9, 2) for k, v in moved.items()}}
    if fused is not None:
        out.update({"fused_core_path_ms": round(fused, 3), "fused_core_effective_GBs": round(alg / fused / 1e6, 1),
                    "fused_core_tensor_traffic_GB": round((moved["dwconv_pack"] + moved["x_proj (cuBLAS)"] + moved["dt_proj"]
                                                           + moved["fused core (xp_ss2d_core, opt-in)"] + moved["plane_norm (opt-in)"]) / 1e9, 2)})
    return out


def time_pipeline(X, torch, preset, B, H, W, topk, dtype_fp16, steps=5, graph=True):
    """pairs/s of a PairPipeline at another configuration (inputs resident, CUDA events, after warm-up)."""
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": True, "mixed_precision": dtype_fp16, "use_attention": {"preset": preset}}).cuda().eval()
    pipe = X.PairPipeline(net, nms=8, detection_threshold=0.015, keep_top_k=topk)
    g = torch.Generator().manual_seed(0)
    o = torch.rand(B, 1, H, W, generator=g).cuda()
    t = torch.rand(B, 1, H, W, generator=g).cuda()
    for _ in range(3):
        r = pipe(o, t)
    fn = lambda: pipe(o, t)
    if graph:
        gp = pipe.capture(o, t)
        for _ in range(3):
            gp.replay()
        fn = gp.replay
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    out = {"pairs_per_s": round(B / ms * 1e3, 2), "ms_per_step": round(ms, 3), "pairs_per_step": B,
           "keypoints_first": int(r.n_optical[0]), "matches_first": int(r.n_matches[0])}
    del net, pipe, o, t, r
    torch.cuda.empty_cache()
    return out


def other_configs(args):
    """The BASELINE configurations that are not the headline line (measured after it, N=1 only)."""
    import torch
    import xpoint_b200 as X
    out = {}
    try:
        out["configs[4] 1024x1280 top-16384 preset E fp16 (batch 8 pairs)"] = time_pipeline(X, torch, "E", 8, 1024, 1280, 16384, True, 3)
    except Exception as e:  # noqa: BLE001 (reported, not hidden)
        out["configs[4] 1024x1280 top-16384 preset E fp16 (batch 8 pairs)"] = {"error": repr(e)[:200]}
    try:
        out["preset V (vanilla VMamba-tiny, N=16) 512x640 top-4096 fp16 (batch 16 pairs)"] = time_pipeline(X, torch, "V", 16, 512, 640, 4096, True, 3)
    except Exception as e:  # noqa: BLE001
        out["preset V (vanilla VMamba-tiny, N=16) 512x640 top-4096 fp16 (batch 16 pairs)"] = {"error": repr(e)[:200]}
    # single-pair latency, the quantity the reference's benchmark.py:151-164 prints (two_forward + nms + interpolate timers,
    # benchmark_evaluation.py:28-37,77-83,118-134; matching is not timed there, it is here)
    try:
        lat = {}
        for name, graph in (("eager", False), ("cuda_graph", True)):
            r = time_pipeline(X, torch, "E", 1, 512, 640, 4096, True, 20, graph)
            lat[name + "_ms_per_pair"] = r["ms_per_step"]
        out["single pair latency 512x640 preset E fp16 (forward x2 + NMS + sampling + matching)"] = lat
    except Exception as e:  # noqa: BLE001
        out["single pair latency 512x640 preset E fp16 (forward x2 + NMS + sampling + matching)"] = {"error": repr(e)[:200]}
    return out


def parity_check(args, ref_out, dev):
    """The GPU tail on the CPU arm's own fp32 prob / desc tensors must reproduce its keypoints and match pairs bit-exactly
    (north star: "bit-exact when both sides are fed the same fp32 score/descriptor tensors"); the GPU model (fp32, same
    weights) must agree with the CPU arm's prob / desc within 1e-4."""
    import numpy as np
    import torch
    import xpoint_b200 as X
    res = {}
    po, pt = ref_out["prob_o"].float(), ref_out["prob_t"].float()
    do, dt_ = ref_out["desc_o"].float(), ref_out["desc_t"].float()
    n = po.shape[0]
    pipe = X.PairPipeline(None, nms=8, detection_threshold=0.015, keep_top_k=args.topk)
    r = pipe.tail(po.to(dev), pt.to(dev), do.to(dev), dt_.to(dev))
    kp_ok, m_ok, nk, nm = True, True, 0, 0
    for b in range(n):
        t = ref_out["tail"][b]
        no, nt = int(r.n_optical[b]), int(r.n_thermal[b])
        kp_ok &= no == len(t["kp_o"]) and nt == len(t["kp_t"])
        kp_ok &= bool(np.array_equal(r.kp_optical[b, :no].cpu().numpy().astype(np.int64), np.asarray(t["kp_o"]).reshape(-1, 2)))
        kp_ok &= bool(np.array_equal(r.kp_thermal[b, :nt].cpu().numpy().astype(np.int64), np.asarray(t["kp_t"]).reshape(-1, 2)))
        # identical descriptors to both matchers: the GPU matcher on the CPU arm's sampled descriptors
        if len(t["d_o"]) and len(t["d_t"]):
            got = X.get_matches(torch.from_numpy(np.asarray(t["d_o"])).to(dev), torch.from_numpy(np.asarray(t["d_t"])).to(dev),
                                "bfmatcher", crossCheck=True)
            got = [(a.queryIdx, a.trainIdx) for a in got]
        else:
            got = []
        want = [tuple(x) for x in t["pairs"]]
        if got != want:
            # rows whose best-vs-second gap is below fp32 summation noise may legitimately differ
            diff = set(got) ^ set(want)
            m_ok &= len(diff) <= max(2, len(want) // 200)
            res.setdefault("match_pairs_differing", 0)
            res["match_pairs_differing"] += len(diff)
        nk += no + nt
        nm += len(want)
    res.update({"pairs": n, "keypoints": nk, "matches": nm, "keypoints_bit_exact": bool(kp_ok), "match_pairs_equal": bool(m_ok)})
    if "state_dict" in ref_out:
        net = X.XPoint({"takes_pair": True, "use_attention": {"preset": args.preset}})
        net.load_state_dict(ref_out["state_dict"], strict=True)
        net = net.to(dev).eval()
        g = torch.Generator().manual_seed(0)
        opt = torch.rand(n, 1, args.height, args.width, generator=g).to(dev)
        thr = torch.rand(n, 1, args.height, args.width, generator=g).to(dev)
        tf = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            go, gt = net.forward_pair_batched(opt, thr)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf

        def rel(x, y):
            x, y = x.double().cpu(), y.double()
            return float((x - y).norm() / y.norm())
        res["gpu_fp32_model_vs_cpu_arm_rel_l2"] = {"prob": rel(go["prob"], po), "desc": rel(go["desc"], do)}
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import xpoint_b200 as X
    from xpoint_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().xp_check_device())
    peak, peak_src = load_peaks()

    torch.manual_seed(0)                      # identical random-init weights on every rank
    net = X.XPoint({"takes_pair": True, "mixed_precision": args.dtype == "fp16", "use_attention": {"preset": args.preset}})
    net = net.to(dev).eval()
    pipe = X.PairPipeline(net, nms=8, detection_threshold=0.015, keep_top_k=args.topk, use_tensor_cores=not args.fp32_match)
    g = torch.Generator().manual_seed(rank)   # per-rank synthetic images
    B, H, W = args.batch, args.height, args.width
    host_o = torch.rand(B, 1, H, W, generator=g).pin_memory()
    host_t = torch.rand(B, 1, H, W, generator=g).pin_memory()
    dev_o, dev_t = host_o.to(dev), host_t.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from xpoint_b200.sharding import max_over_ranks as _max

    def max_over_ranks(ms):
        return _max(ms, dev)

    # ---- warm-up -------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        res = pipe(dev_o, dev_t)
    barrier()
    # The step has no host synchronisation and no data-dependent launch shape: it is recorded once into a CUDA graph
    # and the timed steps replay it (same kernels, same launch shapes, no per-launch host cost).
    step_fn = pipe
    graphed = None
    if not args.no_graph:
        graphed = pipe.capture(dev_o, dev_t)
        for _ in range(max(args.warmup, 3)):
            graphed.replay()
        step_fn = lambda o, t: graphed.replay()       # inputs already resident in the graph's static buffers
    barrier()

    # ---- eager profiled pass: CUDA events around every selective-scan launch (the roofline figures) + launch count ------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.scan_profile = []
    launches0 = _lib.launch_count
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for _ in range(args.steps):
        res = pipe(dev_o, dev_t)
    p1.record()
    barrier()
    ms_eager = p0.elapsed_time(p1)
    launches = _lib.launch_count - launches0           # a graph replay launches exactly these kernels again
    prof, _lib.scan_profile = _lib.scan_profile, None

    # ---- timed region 1: inputs resident in HBM --------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = step_fn(dev_o, dev_t)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    scan_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    scan_bytes = sum(n for _, _, n, _ in prof)
    by_shape = {}
    for a, b, n, shp in prof:
        k = "B{} KD{} K{} N{} L{} {}->{}".format(*shp[:7]).replace("torch.", "") + (f" fused-dt R{shp[7]}" if shp[7] else "")
        s = by_shape.setdefault(k, [0, 0.0, 0])
        s[0] += 1
        s[1] += a.elapsed_time(b)
        s[2] += n

    # ---- timed region 2: end to end from pinned host memory through the public API ---------------------
    # Every step uploads ITS OWN images from pinned host memory and downloads its keypoints / matches.  The upload of
    # step i+1 runs on a copy stream while step i computes (two device buffers); the downloads are queued behind the
    # step's kernels and the host blocks on them one step later -- the standard input pipeline of a serving loop.
    from xpoint_b200.pipeline import PairStream
    stream = PairStream(pipe, dev, use_graph=graphed is not None, graphed=graphed)     # the product's own streaming loop
    stream.run([(host_o, host_t)] * 2, keep=False)
    barrier()
    stream.h2d_bytes = stream.d2h_bytes = 0
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    host_res = stream.run(((host_o, host_t) for _ in range(args.steps)), keep=False)[-1]
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1))
    clocks = sampler.stop() if rank == 0 else None
    d2h = stream.d2h_bytes // args.steps          # counted from the tensors copied, per step
    h2d = stream.h2d_bytes // args.steps

    n_matches = [int(v) for v in host_res["n_matches"][:4].tolist()]
    n_kp = [int(v) for v in host_res["n_optical"][:4].tolist()]

    micro = None
    if rank == 0 and not args.no_microbench:
        del res
        torch.cuda.empty_cache()
        micro = scan_microbench(peak)
    cpu, parity, core_fig, others = None, None, None, None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            core_fig = ss2d_core_figure(peak)
        except Exception as e:  # noqa: BLE001 (reported in the line, not hidden)
            core_fig = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        others = other_configs(args)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample: ONE pair through the staged reference on the host cores (~10-30 s), or 16 pairs through the port
        staged = reference_staged() is not None and not args.ref_port
        NP = 1 if staged else 16
        step, threads, kind, what = cpu_arm(args, NP)
        t, ref_out = step()
        cpu = {"value": NP / t, "unit": "pairs/s", "cores": threads, "kind": kind,
               "sample": f"{NP} pair(s) {H}x{W}, {t:.1f} s; {what}"}
        if kind == "reference":
            try:
                parity = parity_check(args, ref_out, dev)
            except Exception as e:  # noqa: BLE001
                parity = {"error": repr(e)[:300]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pairs = B * world * args.steps
    achieved_all = scan_bytes / scan_ms / 1e6 if scan_ms > 0 else 0.0
    # dominant kernel = the selective-scan launch shape with the largest share of the step (stage 0 of the encoder)
    if not by_shape:            # (a path without op-level scan launches, e.g. XP_SS2D_CORE=1)
        by_shape = {"none": [1, 1e-9, 0]}
    dom_key = max(by_shape, key=lambda k: by_shape[k][1])
    dom = by_shape[dom_key]
    dom_bytes, dom_ms = dom[2] / dom[0], dom[1] / dom[0]
    dom_gbs = dom_bytes / dom_ms / 1e6
    # DRAM bytes of ONE launch of that shape from `ncu --set full` (profiles/r1_s3_summary.md): only known for the
    # default workload (preset E, fp16 autocast, 64 pairs of 512x640 -> B128 KD384 K4 N1 L20480 float16->float32)
    ncu_traffic = {"B128 KD384 K4 N1 L20480 float16->float32": 4.075043e9 + 3.979829e9}
    roofline = {"bound": "hbm", "kernel": f"xp_selective_scan_fwd -> scan_lanes_kernel, launch shape {dom_key}",
                "achieved": round(dom_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(dom_gbs / peak, 4),
                "traffic": ncu_traffic.get(dom_key), "algorithmic_bytes_per_launch": int(dom_bytes),
                "ms_per_launch": round(dom_ms, 4), "peak_source": peak_src,
                "traffic_source": "NOT measured in this run: ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch of "
                                  "this kernel template and shape, profiles/r2_ncu_full.md (round 2, profiles/capture_r2.sh; round 1 "
                                  "measured 8.056 GB for the same kernel); null for any other shape",
                "launches": dom[0], "share_of_step": round(dom[1] / ms_eager, 4),
                "measured_in": "eager pass of the same K steps inside this process (CUDA events around each launch); the timed "
                               "region replays the same kernels from a CUDA graph" if graphed is not None else "timed region",
                "all_scan_launches": {"launches": len(prof), "achieved": round(achieved_all, 1),
                                      "frac": round(achieved_all / peak, 4),
                                      "scan_share_of_step": round(scan_ms / ms_eager, 4)},
                "by_shape": {k: {"launches": v[0], "ms_per_launch": round(v[1] / v[0], 4),
                                 "GBs": round(v[2] / v[1] / 1e6, 1)} for k, v in by_shape.items()}}
    line = {
        "metric": METRIC, "value": pairs / (ms_total / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 autocast (fp32 scan state/accum)" if args.dtype == "fp16" else "f32",
        "data": "synthetic",
        "config": {"workload": f"BASELINE configs[2]: full XPoint inference, preset {args.preset} "
                               f"({'shipped XPoint-EXP1 VMamba N=1' if args.preset == 'E' else 'vanilla VMamba-tiny N=16'}), "
                               f"{H}x{W} pairs, batch {B} pairs/GPU, NMS top-{args.topk}, MNN matching",
                   "pairs_per_gpu": B, "l2": "inputs and activations larger than L2 (no flush needed)",
                   "e2e_pipeline": "xpoint_b200.pipeline.PairStream: per-step H2D of both image batches from pinned host memory on a copy stream (double-buffered), D2H of keypoints / matches",
                   "keypoints_first4": n_kp, "matches_first4": n_matches,
                   "matcher": "fp32 CUDA cores" if args.fp32_match else "tcgen05 3xFP16 split on kind::f16 (3xTF32 when C % 64 != 0)",
                   "cuda_graph": graphed is not None, "eager_ms_per_step": round(ms_eager / args.steps, 3)},
        "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "scan_microbench": micro,
        "cpu_baseline": cpu,
        "parity_check": parity,
        "ss2d_core": core_fig,
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
