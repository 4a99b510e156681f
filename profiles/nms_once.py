"""Run xp_box_nms on model-like score maps (softmax of random logits, pixel-shuffled), 128 images of 512x640."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xpoint_b200 as X
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(0)
lg = 0.3 * torch.randn(B, 65, 64, 80, generator=g, device="cuda")
prob = X.detector_post(lg)[:, 0].contiguous()
for kind, p in (("model-like", prob), ("rand**6", torch.rand(B, 512, 640, generator=g, device="cuda") ** 6)):
    for _ in range(2):
        r = X.nms_keypoints(p, 8, 0.015, keep_top_k=4096, capacity=4096, want_map=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        r = X.nms_keypoints(p, 8, 0.015, keep_top_k=4096, capacity=4096, want_map=False)
    torch.cuda.synchronize()
    print(kind, f"{(time.perf_counter()-t0)/3*1e3:.2f} ms", "candidates/img", int((p > 0.015).sum()) // B, "kept", r.count[:4].tolist())
