"""Single-pair latency (the quantity benchmark.py:151-164 of the reference prints), CUDA-graph replay and eager."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X
torch.manual_seed(0)
net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).cuda().eval()
pipe = X.PairPipeline(net, keep_top_k=4096)
g = torch.Generator().manual_seed(0)
o = torch.rand(1, 1, 512, 640, generator=g).cuda(); t = torch.rand(1, 1, 512, 640, generator=g).cuda()
for _ in range(3): pipe(o, t)
gp = pipe.capture(o, t)
def timeit(fn, n=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(f"single pair: graph {timeit(gp.replay):.3f} ms, eager {timeit(lambda: pipe(o, t)):.3f} ms, matches {int(gp.result.n_matches[0])}")
