"""patch_embed_stem_kernel at the configs[1] shape (128 images 512x640 -> 256x320x48, LayerNorm + GELU, fp16 out)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xpoint_b200.cross_scan import patch_embed_stem
x = torch.rand(128, 1, 512, 640, device="cuda")
w = torch.randn(48, 1, 3, 3, device="cuda") * 0.3; b = torch.randn(48, device="cuda") * 0.1
g = torch.ones(48, device="cuda"); be = torch.zeros(48, device="cuda")
f = lambda: patch_embed_stem(x, w, b, g, be, 1e-5, out_dtype=torch.float16, gelu=True)
for _ in range(3): f()
torch.cuda.synchronize()
a, e = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(20): f()
e.record(); torch.cuda.synchronize()
print(f"patch_embed_stem 128x512x640 -> 48ch: {a.elapsed_time(e) / 20:.3f} ms")
