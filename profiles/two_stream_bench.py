"""Experiment: one 64-pair step as ONE graph vs. two concurrent 32-pair half-steps on two streams inside one graph
(kernel tails / launch gaps of one half overlap the other half's kernels).
Result on B200 (round 1): 34.35 ms for the single graph, 37.07 ms for two halves, 37.73 ms for four quarters -- the kernels
are bandwidth-bound, concurrency only costs efficiency; the pipeline stays one stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X

torch.manual_seed(0)
net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).cuda().eval()
pipe = X.PairPipeline(net, keep_top_k=4096)
g = torch.Generator().manual_seed(0)
B = 64
o = torch.rand(B, 1, 512, 640, generator=g).cuda(); t = torch.rand(B, 1, 512, 640, generator=g).cuda()

def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for _ in range(3): pipe(o, t)
gp = pipe.capture(o, t)
print(f"one graph, 64 pairs: {timeit(gp.replay):.2f} ms")

for parts in (2, 4):
    hb = B // parts
    for _ in range(3): pipe(o[:hb], t[:hb])
    streams = [torch.cuda.Stream() for _ in range(parts)]
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cur = torch.cuda.current_stream()
        res = []
        for i, s in enumerate(streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                res.append(pipe(o[i * hb:(i + 1) * hb], t[i * hb:(i + 1) * hb]))
        for s in streams:
            cur.wait_stream(s)
    print(f"{parts} concurrent part-steps of {hb} pairs in one graph: {timeit(graph.replay):.2f} ms", [int(r.n_matches[0]) for r in res])
