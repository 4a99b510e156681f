"""One call of the tcgen05 matcher at the bench shape (64 pairs x 4096 x 4096 x 256) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X
g = torch.Generator(device="cuda").manual_seed(0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d1 = torch.nn.functional.normalize(torch.randn(P, 4096, 256, generator=g, device="cuda"), dim=-1)
d2 = torch.nn.functional.normalize(torch.randn(P, 4096, 256, generator=g, device="cuda"), dim=-1)
for _ in range(2):
    X.mnn_match(d1, d2, use_tensor_cores=True)
torch.cuda.synchronize()
