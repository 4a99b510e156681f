"""Full-size sanity check (GPU): preset E at 512x640, compare fp32 / fp16, fused / unfused SS2D paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import xpoint_b200 as X

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator().manual_seed(0)
o = torch.rand(B, 1, 512, 640, generator=g).cuda()
t = torch.rand(B, 1, 512, 640, generator=g).cuda()


def run(mixed, fused):
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": True, "mixed_precision": mixed, "use_attention": {"preset": "E"}}).cuda().eval()
    for m in net.modules():
        if isinstance(m, X.SS2D):
            m.disable_fused = not fused
    with torch.no_grad():
        po, pt = net.forward_pair_batched(o, t)
    return {k: po[k].float().cpu().numpy() for k in ("encoder_output", "prob", "desc")}


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())), float(np.abs(a - b).max() / np.abs(b).max())


ref = run(False, False)
for name, (mixed, fused) in {"fp32 fused": (False, True), "fp16 unfused": (True, False), "fp16 fused": (True, True)}.items():
    r = run(mixed, fused)
    print(name, {k: tuple(f"{v:.2e}" for v in rel(r[k], ref[k])) for k in ref})
