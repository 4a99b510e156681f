"""Per-kernel time table of one bench step (torch.profiler / CUPTI) -- a cheap complement to the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import xpoint_b200 as X

preset = sys.argv[1] if len(sys.argv) > 1 else "E"
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
torch.manual_seed(0)
net = X.XPoint({"takes_pair": True, "mixed_precision": dtype == "fp16", "use_attention": {"preset": preset}}).cuda().eval()
pipe = X.PairPipeline(net, keep_top_k=4096)
g = torch.Generator().manual_seed(0)
o = torch.rand(B, 1, 512, 640, generator=g).cuda(); t = torch.rand(B, 1, 512, 640, generator=g).cuda()
for _ in range(3):
    pipe(o, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pipe(o, t)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith("aten::")
      and str(e.device_type).endswith("CUDA")]   # kernels only (CPU-side op rows would double count)
tot = sum(e.device_time_total for e in ev)
print(f"total device time {tot/1e3:.2f} ms, {sum(e.count for e in ev)} launches")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:45]:
    print(f"{e.device_time_total/1e3:8.3f} ms {100*e.device_time_total/tot:5.1f}%  n={e.count:4d}  {e.key[:110]}")
