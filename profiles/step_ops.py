"""Which ATen ops (with input shapes) still launch kernels in one bench step -- finds leftover copies / elementwise passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import xpoint_b200 as X

torch.manual_seed(0)
net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).cuda().eval()
pipe = X.PairPipeline(net, keep_top_k=4096)
g = torch.Generator().manual_seed(0)
B = 64
o = torch.rand(B, 1, 512, 640, generator=g).cuda(); t = torch.rand(B, 1, 512, 640, generator=g).cuda()
for _ in range(3):
    pipe(o, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    pipe(o, t)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0 and e.key.startswith("aten::")]
for e in sorted(rows, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"{e.self_device_time_total/1e3:8.3f} ms n={e.count:3d} {e.key:28s} {str(e.input_shapes)[:150]}")
