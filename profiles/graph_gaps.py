"""How much of a CUDA-graph-replayed step is idle?  Profiles ONE replay of the bench step (preset E, 64 pairs of 512x640, fp16)
with the torch profiler (CUPTI kernel records) and reports: span (first kernel start -> last kernel end), sum of kernel
durations, and the gaps between consecutive kernels.   python profiles/graph_gaps.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).cuda().eval()
pipe = X.PairPipeline(net, keep_top_k=4096)
g = torch.Generator().manual_seed(0)
o, t = torch.rand(64, 1, 512, 640, generator=g).cuda(), torch.rand(64, 1, 512, 640, generator=g).cuda()
gp = pipe.capture(o, t)
for _ in range(5):
    gp.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gp.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph replay, CUDA events: {e0.elapsed_time(e1) / 10:.3f} ms per step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gp.replay()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time > 0 and "Memcpy" not in e.name and "Memset" not in e.name],
            key=lambda e: e.time_range.start)
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(e.device_time for e in ev)
gaps = [max(0.0, b.time_range.start - a.time_range.end) for a, b in zip(ev, ev[1:])]
print(f"one profiled replay: {len(ev)} kernels, span {span / 1e3:.3f} ms, kernel time {busy / 1e3:.3f} ms, gaps {sum(gaps) / 1e3:.3f} ms "
      f"(mean {sum(gaps) / max(len(gaps), 1):.2f} us, max {max(gaps):.1f} us)")
big = sorted(zip(gaps, ev, ev[1:]), key=lambda x: -x[0])[:6]
for gdur, a, b in big:
    print(f"  gap {gdur:7.1f} us after {a.name[:60]} -> before {b.name[:60]}")
