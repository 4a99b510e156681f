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
  cpu_baseline     the oracle port of the same path on the host cores, 1 pair (N=1 only)
--impl reference times that CPU path as the reference arm (rank 0 only).
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
    """The oracle port of the whole path (oracle/model.py + oracle/xp_oracle.c) on the host cores."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    NP = 8                                   # pairs per step: a bounded sample (~5 s of host work per step)
    step, threads = cpu_pair_run(args, NP)
    for _ in range(min(args.warmup, 1)):
        step()
    times = [step()[0] for _ in range(args.steps)]
    per = sum(times) / len(times)
    v = NP / per
    line = {"metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"XPoint preset {args.preset} pair inference {args.height}x{args.width}, top-{args.topk} keypoints, "
                                   "MNN matching; CPU oracle port of the reference path", "pairs_per_step": NP},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{NP} pairs per step x {args.steps} steps (oracle/model.py + xp_oracle.c, torch CPU for dense layers)"},
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
        out.append({"case": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3), "achieved_GBs": round(gbs, 1),
                    "frac_of_hbm_peak": round(gbs / peak, 4)})
        del u, dl, Bm, Cm, y
        torch.cuda.empty_cache()
    return out


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
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [(torch.empty_like(dev_o), torch.empty_like(dev_t)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    out_names = ("kp_optical", "kp_thermal", "n_optical", "n_thermal", "match_idx", "match_dist", "n_matches")
    host_out = None

    def upload(i):
        o, t = dev_bufs[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the step that last read this buffer pair has finished
            o.copy_(host_o, non_blocking=True)
            t.copy_(host_t, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        nonlocal host_out
        main = torch.cuda.current_stream(dev)
        for ev in consumed:
            ev.record(main)
        upload(0)
        pending = None
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            main.wait_event(ready[i % 2])
            o, t = dev_bufs[i % 2]
            if graphed is not None:
                graphed.load(o, t)                           # device-to-device into the graph's static inputs
                consumed[i % 2].record(main)
                r = graphed.replay()
            else:
                r = pipe(o, t)
                consumed[i % 2].record(main)
            outs = [getattr(r, k) for k in out_names]
            if host_out is None:
                host_out = [[torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in outs] for _ in range(2)]
            for h, x in zip(host_out[i % 2], outs):
                h.copy_(x, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending.synchronize()                        # results of step i-1 are on the host
            pending = done
        pending.synchronize()
        return host_out[(n - 1) % 2]

    e2e_loop(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    host_res = e2e_loop(args.steps)
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1))
    clocks = sampler.stop() if rank == 0 else None
    d2h = sum(x.numel() * x.element_size() for x in host_res)
    h2d = host_o.numel() * 4 * 2

    n_matches = [int(v) for v in host_res[-1][:4].tolist()]
    n_kp = [int(v) for v in host_res[2][:4].tolist()]

    micro = None
    if rank == 0 and not args.no_microbench:
        del res
        torch.cuda.empty_cache()
        micro = scan_microbench(peak)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        NP = 16                              # bounded sample: ~10-20 s of host work
        step, threads = cpu_pair_run(args, NP)
        t, _ = step()
        cpu = {"value": NP / t, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": f"{NP} pairs {H}x{W} through oracle/model.py + xp_oracle.c (torch CPU for dense layers), {t:.1f} s"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pairs = B * world * args.steps
    achieved_all = scan_bytes / scan_ms / 1e6 if scan_ms > 0 else 0.0
    # dominant kernel = the selective-scan launch shape with the largest share of the step (stage 0 of the encoder)
    dom_key = max(by_shape, key=lambda k: by_shape[k][1])
    dom = by_shape[dom_key]
    dom_bytes, dom_ms = dom[2] / dom[0], dom[1] / dom[0]
    dom_gbs = dom_bytes / dom_ms / 1e6
    # DRAM bytes of ONE launch of that shape from `ncu --set full` (profiles/r1_s3_summary.md): only known for the
    # default workload (preset E, fp16 autocast, 64 pairs of 512x640 -> B128 KD384 K4 N1 L20480 float16->float32)
    ncu_traffic = {"B128 KD384 K4 N1 L20480 float16->float32": 4.075503e9 + 3.980529e9}
    roofline = {"bound": "hbm", "kernel": f"xp_selective_scan_fwd -> scan_lanes_kernel, launch shape {dom_key}",
                "achieved": round(dom_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(dom_gbs / peak, 4),
                "traffic": ncu_traffic.get(dom_key), "algorithmic_bytes_per_launch": int(dom_bytes),
                "ms_per_launch": round(dom_ms, 4), "peak_source": peak_src,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_s3_summary.md",
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
                   "e2e_pipeline": "per-step H2D of both image batches on a copy stream (double-buffered), D2H of keypoints/matches",
                   "keypoints_first4": n_kp, "matches_first4": n_matches,
                   "matcher": "fp32 CUDA cores" if args.fp32_match else "tcgen05 3xTF32",
                   "cuda_graph": graphed is not None, "eager_ms_per_step": round(ms_eager / args.steps, 3)},
        "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "scan_microbench": micro,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
