#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What it imports from the reference (nothing is copied into this repo):
  * xpoint.models.vmamba_src.csms6s.selective_scan_torch           (csms6s.py:25-68)
  * selective_scan_ref, exec'd from the kernel test file's source   (test_selective_scan.py:168-234)
  * xpoint.models.vmamba_src.csm_triton.cross_scan_fwd / cross_merge_fwd / *1b1*  (csm_triton.py:22-179)
  * xpoint.models.vmamba_src.VMamba.SS2D / VSSM                     (VMamba.py:1107, :1243)
  * xpoint.models.XPoint.XPoint                                     (XPoint.py:27)
  * xpoint.utils.utils.box_nms / interpolate_descriptors            (utils.py:148, :229)
  * xpoint.utils.matching.get_matches / NNMatcher                   (matching.py:4, :38)
Import shims for timm / fvcore / yacs live in tests/golden/_shims (those packages are not in
the image).  The only behavioural patch is the one SURVEY 0.8 documents: `cross_scan_fn`
wraps its call in torch.cuda.device(x.device), which raises on CPU, so VMamba's
cross_scan_fn / cross_merge_fn are rebound to the reference's own CrossScanF / CrossMergeF.
"""
import ast
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("XPOINT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import xpoint.models as xmodels  # noqa: E402
import xpoint.utils as xutils  # noqa: E402
from xpoint.models.vmamba_src import VMamba as RV  # noqa: E402
from xpoint.models.vmamba_src import csm_triton as RC  # noqa: E402
from xpoint.models.vmamba_src import csms6s as RS  # noqa: E402

torch.set_num_threads(8)


def load_selective_scan_ref():
    path = os.path.join(REF, "xpoint/models/vmamba_src/kernels/selective_scan/test_selective_scan.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "selective_scan_ref":
            code = ast.get_source_segment(src, node)
            ns = {}
            import torch.nn.functional as F
            from einops import rearrange, repeat
            ns.update(torch=torch, F=F, rearrange=rearrange, repeat=repeat)
            exec(code, ns)
            return ns["selective_scan_ref"]
    raise RuntimeError("selective_scan_ref not found")


selective_scan_ref = load_selective_scan_ref()


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().float().cpu().numpy() if v.is_floating_point() else v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {name}.npz  {os.path.getsize(path)/1024:.1f} KiB")


# ------------------------------------------------------------------------------ selective scan
def scan_inputs(seed, Bt, KD, K, N, L, KD1=None, dtype=torch.float32, three_d=False):
    """Input distributions of test_selective_scan.py:414-444."""
    g = torch.Generator().manual_seed(seed)
    KD1 = KD1 or KD
    A = -0.5 * torch.rand(KD, N, generator=g)
    shape = (Bt, N, L) if three_d else (Bt, K, N, L)
    B = torch.randn(*shape, generator=g).to(dtype)
    C = torch.randn(*shape, generator=g).to(dtype)
    D = torch.randn(KD, generator=g)
    z = torch.randn(Bt, KD, L, generator=g).to(dtype)
    bias = 0.5 * torch.rand(KD1, generator=g)
    u = torch.randn(Bt, KD, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(Bt, KD1, L, generator=g)).to(dtype)
    return u, delta, A, B, C, D, z, bias


def gen_scan():
    # XPoint-style grouped call through the importable API (csms6s.py:112 -> selective_scan_torch)
    for name, (Bt, KD, K, N, L) in dict(scan_n16_k4=(2, 32, 4, 16, 300), scan_n1_k4=(2, 24, 4, 1, 257),
                                        scan_n8_k2=(1, 12, 2, 8, 64)).items():
        u, delta, A, B, C, D, z, bias = scan_inputs(0, Bt, KD, K, N, L)
        out = RS.selective_scan_fn(u, delta, A, B, C, D, bias, True, True, backend="torch")
        out_t = RS.selective_scan_torch(u, delta, A, B, C, D, bias, True, True)
        out_r, last = selective_scan_ref(u, delta, A, B, C, D, None, bias, True, return_last_state=True)
        assert torch.equal(out, out_t)
        print(name, "selective_scan_torch vs selective_scan_ref max|diff| =", (out_t - out_r).abs().max().item())
        save(name, u=u, delta=delta, A=A, B=B, C=C, D=D, delta_bias=bias, out=out_t, out_ref=out_r, last_state=last,
             delta_softplus=1)
    # mamba-style extras: z gate + last state, no softplus, no bias, no D  (test_selective_scan.py:168-234)
    u, delta, A, B, C, D, z, bias = scan_inputs(1, 2, 16, 2, 8, 130)
    out, last = selective_scan_ref(u, delta, A, B, C, None, z, None, False, return_last_state=True)
    save("scan_z_last", u=u, delta=delta, A=A, B=B, C=C, z=z, out=out, last_state=last, delta_softplus=0)
    # delta groups: dim1 != dim, reference repeats delta rows (test_selective_scan.py:446-450)
    u, delta, A, B, C, D, z, bias = scan_inputs(2, 2, 24, 2, 4, 96, KD1=6)
    delta_rep = delta.unsqueeze(2).repeat(1, 1, 4, 1).contiguous().flatten(1, 2)
    bias_rep = bias.unsqueeze(1).repeat(1, 4).view(-1)
    out, last = selective_scan_ref(u, delta_rep, A, B, C, D, None, bias_rep, True, return_last_state=True)
    save("scan_dgroups", u=u, delta=delta, A=A, B=B, C=C, D=D, delta_bias=bias, out=out, last_state=last,
         delta_softplus=1)
    # 3-D B/C (single group)
    u, delta, A, B, C, D, z, bias = scan_inputs(3, 2, 8, 1, 4, 70, three_d=True)
    out, last = selective_scan_ref(u, delta, A, B, C, D, z, bias, True, return_last_state=True)
    save("scan_3d", u=u, delta=delta, A=A, B=B, C=C, D=D, z=z, delta_bias=bias, out=out, last_state=last,
         delta_softplus=1)
    # 16-bit inputs: both sides get the same pre-rounded tensors; oflex=True -> fp32 out (csms6s.py:68)
    for tag, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        u, delta, A, B, C, D, z, bias = scan_inputs(4, 2, 16, 4, 16, 200, dtype=dt)
        out = RS.selective_scan_torch(u, delta, A, B, C, D, bias, True, True)
        out_in = RS.selective_scan_torch(u, delta, A, B, C, D, bias, True, False)
        assert out.dtype == torch.float32 and out_in.dtype == dt
        save("scan_" + tag, u=u, delta=delta, A=A, B=B, C=C, D=D, delta_bias=bias, out=out, out_indtype=out_in,
             delta_softplus=1)


# ------------------------------------------------------------------------------ cross scan/merge
def gen_cross():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 5, 7, generator=g)
    x4 = torch.randn(2, 4, 3, 5, 7, generator=g)
    ys = torch.randn(2, 4, 3, 5, 7, generator=g)
    arrs = dict(x=x, x4=x4, ys=ys)
    for s in (0, 1, 2):
        arrs[f"scan_s{s}"] = RC.cross_scan_fwd(x, True, True, s)
        arrs[f"scan_s{s}_cl"] = RC.cross_scan_fwd(x.permute(0, 2, 3, 1).contiguous(), False, False, s)
        arrs[f"scan_s{s}_cf2cl"] = RC.cross_scan_fwd(x, True, False, s)
        arrs[f"merge_s{s}"] = RC.cross_merge_fwd(ys, True, True, s)
        arrs[f"merge_s{s}_cl"] = RC.cross_merge_fwd(ys.permute(0, 3, 4, 1, 2).contiguous(), False, False, s)
        arrs[f"scan1b1_s{s}"] = RC.cross_scan1b1_fwd(x4, True, True, s)
        arrs[f"merge1b1_s{s}"] = RC.cross_merge1b1_fwd(ys, True, True, s)
    # the autograd Function wrappers return (B,4,C,L) / (B,C,L) like the functional forms
    arrs["F_scan"] = RC.CrossScanF.apply(x, True, True, False, 0)
    arrs["F_merge"] = RC.CrossMergeF.apply(ys, True, True, False, 0)
    save("cross", **arrs)


# ------------------------------------------------------------------------------ SS2D block
def patch_cross_for_cpu():
    RV.cross_scan_fn = lambda x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False: \
        RC.CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)
    RV.cross_merge_fn = lambda y, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False: \
        RC.CrossMergeF.apply(y, in_channel_first, out_channel_first, one_by_one, scans)


def gen_ss2d():
    patch_cross_for_cpu()
    for name, kw in dict(
        ss2d_v0=dict(d_model=16, d_state=4, ssm_ratio=2.0, forward_type="v0"),
        ss2d_v05_noz=dict(d_model=16, d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False),
        ss2d_v05=dict(d_model=16, d_state=2, ssm_ratio=2.0, forward_type="v05"),
    ).items():
        torch.manual_seed(0)
        m = RV.SS2D(**kw).eval()
        with torch.no_grad():
            # make A / D / norm non-trivial so the test does not pass on initialisation structure
            m.A_logs.copy_(torch.log(0.5 * torch.rand_like(m.A_logs) + 0.05))
            m.Ds.copy_(torch.randn_like(m.Ds))
            m.out_norm.weight.copy_(1 + 0.1 * torch.randn_like(m.out_norm.weight))
            m.out_norm.bias.copy_(0.1 * torch.randn_like(m.out_norm.bias))
            x = torch.randn(2, 6, 9, 16)
            y = m(x)
            # a second input whose H, W are multiples of 4: the shape class the copy-free SS2D path covers
            x4 = torch.randn(2, 12, 20, 16)
            y4 = m(x4)
        arrs = {"sd." + k: v for k, v in m.state_dict().items()}
        save(name, x=x, y=y, x4=x4, y4=y4, **arrs)


# ------------------------------------------------------------------------------ heads + tail
def gen_tail():
    g = torch.Generator().manual_seed(0)
    logits = 3.0 * torch.randn(2, 65, 6, 8, generator=g)
    prob = torch.nn.PixelShuffle(8)(torch.nn.Softmax2d()(logits)[:, :-1])  # XPoint.py:356-357
    draw = torch.randn(2, 256, 6, 8, generator=g)
    dnorm = torch.nn.functional.normalize(draw, p=2, dim=1)  # XPoint.py:366
    save("heads", logits=logits, prob=prob, desc_raw=draw, desc=dnorm)

    p1 = torch.rand(96, 120, generator=g) ** 6
    pb = torch.rand(3, 1, 64, 80, generator=g) ** 6
    arrs = dict(prob=p1, prob_batched=pb)
    arrs["nms8"] = xutils.box_nms(p1, 8, 0.015)
    arrs["nms4"] = xutils.box_nms(p1, 4, 0.015)
    arrs["nms8_top50"] = xutils.box_nms(p1, 8, 0.015, keep_top_k=50)
    arrs["nms8_iou03"] = xutils.box_nms(p1, 8, 0.015, iou=0.3)
    arrs["nms8_batched_top40"] = xutils.box_nms(pb, 8, 0.015, keep_top_k=40, on_cpu=True)
    arrs["nms8_batched"] = xutils.box_nms(pb, 8, 0.015)
    kp = torch.nonzero(arrs["nms8"] > 0.015)  # evaluation.py:281-282
    arrs["kp8"] = kp
    desc_low = torch.nn.functional.normalize(torch.randn(256, 12, 15, generator=g), p=2, dim=0)
    arrs["desc_low"] = desc_low
    arrs["desc_kp"] = xutils.interpolate_descriptors(kp, desc_low, 96, 120)
    save("tail", **arrs)


def gen_match():
    seed = 0
    while True:
        g = torch.Generator().manual_seed(seed)
        d1 = torch.nn.functional.normalize(torch.randn(300, 256, generator=g), dim=1)
        # second set: half noisy copies of d1 rows (so many mutual matches), half random
        perm = torch.randperm(300, generator=g)[:140]
        d2a = torch.nn.functional.normalize(d1[perm] + 0.08 * torch.randn(140, 256, generator=g), dim=1)
        d2b = torch.nn.functional.normalize(torch.randn(140, 256, generator=g), dim=1)
        d2 = torch.cat([d2a, d2b])[torch.randperm(280, generator=g)]
        dm = torch.cdist(d1.double(), d2.double())
        s1 = dm.sort(1).values
        s2 = dm.sort(0).values
        gap = min((s1[:, 1] - s1[:, 0]).min().item(), (s2[1] - s2[0]).min().item())
        if gap > 1e-5:
            break
        seed += 1
    n1, n2 = d1.numpy(), d2.numpy()
    m = xutils.get_matches(n1, n2, "bfmatcher", False, crossCheck=True)  # matching.py:7,34
    bf = np.array([(x.queryIdx, x.trainIdx) for x in m], dtype=np.int32)
    bfd = np.array([x.distance for x in m], dtype=np.float32)
    m2 = xutils.get_matches(n1, n2, "nnmatcher", False, threshold=10.0)  # matching.py:38-75
    nn_ = np.array([(x.queryIdx, x.trainIdx) for x in m2], dtype=np.int32)
    nnd = np.array([x.distance for x in m2], dtype=np.float32)
    print("matches: bf", len(bf), "nn", len(nn_), "min gap", gap, "seed", seed)
    save("match", d1=d1, d2=d2, bf_pairs=bf, bf_dist=bfd, nn_pairs=nn_, nn_dist=nnd, min_gap=gap)


# ------------------------------------------------------------------------------ whole model (config 1, tiny dims)
TINY = dict(
    E=dict(DEPTHS=[1, 1, 1, 1], DOWNSAMPLE="v3", EMBED_DIM=16, MLP_RATIO=4.0, PATCHEMBED="v2", SSM_CONV=3,
           SSM_CONV_BIAS=False, SSM_DT_RANK="auto", SSM_D_STATE=1, SSM_FORWARDTYPE="v05_noz", SSM_RATIO=1.0),
    V=dict(DEPTHS=[1, 1, 2, 1], DOWNSAMPLE="v1", EMBED_DIM=16, MLP_RATIO=0.0, PATCHEMBED="v1", SSM_CONV=3,
           SSM_CONV_BIAS=True, SSM_DT_RANK="auto", SSM_D_STATE=4, SSM_FORWARDTYPE="v0", SSM_RATIO=2.0,
           SSM_INIT="v0", NORM_LAYER="ln"),
)


def gen_model():
    import yaml
    patch_cross_for_cpu()
    for tag, vssm in TINY.items():
        with tempfile.TemporaryDirectory() as td:
            ypath = os.path.join(td, "vssm_tiny.yaml")
            with open(ypath, "w") as f:
                yaml.safe_dump({"MODEL": {"TYPE": "vssm", "NAME": "tiny", "DROP_PATH_RATE": 0.2, "VSSM": vssm}}, f)
            cfg = dict(
                multispectral=False, descriptor_head=True, descriptor_size=256, normalize_descriptors=True,
                final_batchnorm=True, reflection_pad=True, bn_first=False, mixed_precision=False, takes_pair=True,
                homography_regression_head=dict(check=False, type="RegNet"),
                use_attention=dict(check=True, type="VMamba", height=64, width=96,
                                   pretrained=dict(check=True, yaml_file=ypath), model_parameters={}),
            )
            torch.manual_seed(0)
            net = xmodels.XPoint(cfg).eval()
        with torch.no_grad():
            # BatchNorm running stats / SSM params away from their init values
            for mod in net.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_mean.normal_(0, 0.1)
                    mod.running_var.uniform_(0.5, 1.5)
                    mod.weight.normal_(1.0, 0.1)
                    mod.bias.normal_(0, 0.1)
            for n_, p in net.named_parameters():
                if n_.endswith("A_logs"):
                    p.copy_(torch.log(0.5 * torch.rand_like(p) + 0.05))
                if n_.endswith("Ds"):
                    p.normal_(1.0, 0.3)
            g = torch.Generator().manual_seed(1)
            data = {"optical": {"image": torch.rand(2, 1, 64, 96, generator=g)},
                    "thermal": {"image": torch.rand(2, 1, 64, 96, generator=g)}}
            po, pt, hm = net(data)
        assert hm is None and po["logits"] is None
        arrs = {"sd." + k: v for k, v in net.state_dict().items()}
        save("xpoint_tiny_" + tag, img_optical=data["optical"]["image"], img_thermal=data["thermal"]["image"],
             prob_optical=po["prob"], desc_optical=po["desc"], enc_optical=po["encoder_output"],
             prob_thermal=pt["prob"], desc_thermal=pt["desc"], enc_thermal=pt["encoder_output"], **arrs)
        print(tag, "params", sum(p.numel() for p in net.parameters()))


def gen_homography():
    """The reference's homography step, exactly as evaluation.py:359-378 calls it: cv2.findHomography(optical_pts, thermal_pts,
    USAC_MAGSAC, ransacReprojThreshold=3, confidence=0.9999, maxIters=10000) on (x, y) float32 points shaped (-1, 1, 2).
    Synthetic matches: integer keypoints of a 512x640 pair related by a known homography (+ rounding noise) and a share of
    wrong matches; the fixture keeps the keypoints, the match list, the ground truth, OpenCV's H and its inlier mask."""
    import cv2
    rng = np.random.default_rng(0)
    Hh, Ww, k = 512, 640, 1500
    cases = {}
    for name, outlier_share, n_match in (("easy", 0.1, 1200), ("hard", 0.6, 900), ("few", 0.3, 12)):
        ang, sc = rng.uniform(-0.2, 0.2), rng.uniform(0.9, 1.1)
        Hgt = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-30, 30)],
                        [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-30, 30)],
                        [rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4), 1.0]])
        kp1 = np.stack([rng.integers(0, Hh, k), rng.integers(0, Ww, k)], 1).astype(np.int32)       # (y, x)
        xy1 = np.concatenate([kp1[:, ::-1].astype(np.float64), np.ones((k, 1))], 1)
        w = xy1 @ Hgt.T
        xy2 = w[:, :2] / w[:, 2:]
        kp2 = np.round(xy2[:, ::-1]).astype(np.int32)                                             # (y, x), rounded to pixels
        perm = rng.permutation(k)
        kp2s = kp2[perm]                                    # image-2 keypoints in their own order
        inv = np.argsort(perm)                              # kp1 row i <-> kp2s row inv[i]
        match_idx = np.full(k, -1, np.int32)
        rows = rng.choice(k, n_match, replace=False)
        match_idx[rows] = inv[rows]
        wrong = rows[rng.random(n_match) < outlier_share]
        match_idx[wrong] = rng.integers(0, k, wrong.size)
        inb = (kp2s[:, 0] >= 0) & (kp2s[:, 0] < Hh) & (kp2s[:, 1] >= 0) & (kp2s[:, 1] < Ww)
        match_idx[(match_idx >= 0) & ~inb[np.maximum(match_idx, 0)]] = -1
        q = np.nonzero(match_idx >= 0)[0]
        optical_pts = np.float32(kp1[q][:, ::-1]).reshape(-1, 1, 2)
        thermal_pts = np.float32(kp2s[match_idx[q]][:, ::-1]).reshape(-1, 1, 2)
        cv2.setRNGSeed(0)
        H_cv, mask = cv2.findHomography(optical_pts, thermal_pts, method=cv2.USAC_MAGSAC, ransacReprojThreshold=3.0,
                                        confidence=0.9999, maxIters=10000)
        cases.update({f"{name}_kp1": kp1, f"{name}_kp2": kp2s, f"{name}_match_idx": match_idx, f"{name}_H_gt": Hgt,
                      f"{name}_H_cv": H_cv, f"{name}_mask_cv": mask.ravel().astype(np.uint8), f"{name}_query": q.astype(np.int32)})
        print(name, "matches", q.size, "cv2 inliers", int(mask.sum()))
    save("homography", height=np.int32(Hh), width=np.int32(Ww), cv2_version=np.array(cv2.__version__), **cases)


def gen_metrics():
    """The evaluation driver's per-sample metrics through the reference's own functions
    (benchmark_evaluation.py:396-467 compute_repeatability_for_sample, :588-750 compute_descriptor_for_sample;
    homographies.py:479-526 warp_keypoints / filter_points) on a synthetic batch of 3 pairs at 96x128: sparse score maps
    (already non-maximum-suppressed), unit descriptor maps, identity / affine / projective ground-truth homographies."""
    from xpoint.utils import benchmark_evaluation as BE
    from xpoint.utils import homographies as HM
    g = torch.Generator().manual_seed(0)
    Bn, Hh, Ww = 3, 96, 128
    Hs = torch.eye(3).repeat(Bn, 1, 1)
    Ht = torch.stack([torch.eye(3),
                      torch.tensor([[0.98, -0.05, 3.0], [0.04, 1.01, -2.0], [0.0, 0.0, 1.0]]),
                      torch.tensor([[1.02, 0.03, -4.0], [-0.02, 0.97, 5.0], [1e-4, -2e-4, 1.0]])])
    prob_o = torch.zeros(Bn, 1, Hh, Ww)
    prob_t = torch.zeros(Bn, 1, Hh, Ww)
    for b in range(Bn):
        n = 160
        ys, xs = torch.randint(0, Hh, (n,), generator=g), torch.randint(0, Ww, (n,), generator=g)
        prob_o[b, 0, ys, xs] = 0.1 + 0.8 * torch.rand(n, generator=g)
        # thermal keypoints: the optical ones moved by the ground truth (+ jitter), plus clutter
        kp = torch.nonzero(prob_o[b, 0] > 0.015)
        w = HM.warp_keypoints(kp.numpy(), Ht[b].numpy(), float)
        w = np.round(w + np.random.default_rng(b).normal(0, 0.7, w.shape)).astype(int)
        ok = (w[:, 0] >= 0) & (w[:, 0] < Hh) & (w[:, 1] >= 0) & (w[:, 1] < Ww)
        w = w[ok][: int(0.8 * ok.sum())]
        prob_t[b, 0, w[:, 0], w[:, 1]] = 0.1 + 0.8 * torch.rand(len(w), generator=g)
        ys, xs = torch.randint(0, Hh, (40,), generator=g), torch.randint(0, Ww, (40,), generator=g)
        prob_t[b, 0, ys, xs] = 0.1 + 0.8 * torch.rand(40, generator=g)
    base = torch.randn(Bn, 256, Hh // 8, Ww // 8, generator=g)
    desc_o = torch.nn.functional.normalize(base, dim=1)
    desc_t = torch.nn.functional.normalize(base + 0.35 * torch.randn(base.shape, generator=g), dim=1)
    data = {"optical": {"image": torch.zeros(Bn, 1, Hh, Ww), "valid_mask": torch.ones(Bn, 1, Hh, Ww), "homography": Hs},
            "thermal": {"image": torch.zeros(Bn, 1, Hh, Ww), "valid_mask": torch.ones(Bn, 1, Hh, Ww), "homography": Ht}}
    thr_rep, thr_kp = [1, 3], [2, 4]
    rep = {th: [] for th in thr_rep}
    nko, nkt = [], []
    for b in range(Bn):          # the reference function overwrites its per-threshold dict per sample: call it sample by sample
        d1 = {k: {kk: vv[b:b + 1] for kk, vv in v.items()} for k, v in data.items()}
        r, a, c = BE.compute_repeatability_for_sample({"prob": prob_o[b:b + 1]}, {"prob": prob_t[b:b + 1]}, d1, Hs[b:b + 1],
                                                      Ht[b:b + 1], 0.015, thr_rep)
        for th in thr_rep:
            rep[th].append(r[th][0] if r[th] else float("nan"))
        nko += a
        nkt += c
    cfg = {"prediction": {"matching": {"method": "bfmatcher", "knn_matches": False, "method_kwargs": {"crossCheck": True}}}}
    arrs = dict(prob_o=prob_o, prob_t=prob_t, desc_o=desc_o, desc_t=desc_t, H_o=Hs, H_t=Ht, thr_rep=np.array(thr_rep, float),
                thr_kp=np.array(thr_kp, float), n_kp_o=np.array(nko), n_kp_t=np.array(nkt),
                repeatability=np.array([rep[th] for th in thr_rep]).T)
    ms_o, ms_t, nm = [], [], []
    for b in range(Bn):
        d1 = {k: {kk: vv[b:b + 1] for kk, vv in v.items()} for k, v in data.items()}
        dd = BE.compute_descriptor_for_sample(prob_o[b:b + 1], prob_t[b:b + 1], desc_o[b:b + 1], desc_t[b:b + 1], d1, cfg, 0.015, thr_kp)
        ms_o.append([dd[th]["m_score_optical"][0] for th in thr_kp])
        ms_t.append([dd[th]["m_score_thermal"][0] for th in thr_kp])
        nm.append([int(sum(dd[th]["tp_optical"])) for th in thr_kp])
    arrs.update(m_score_optical=np.array(ms_o), m_score_thermal=np.array(ms_t), n_correct_optical=np.array(nm))
    # warp_keypoints itself, int and float flavours
    kp = torch.nonzero(prob_o[2, 0] > 0.015).numpy()
    arrs.update(wk_in=kp, wk_int=HM.warp_keypoints(kp, Ht[2].numpy()), wk_float=HM.warp_keypoints(kp.astype(np.float32), Ht[2].numpy(), float))
    save("metrics", **arrs)
    print("metrics: repeatability", arrs["repeatability"].round(3).tolist(), "m_score_o", np.array(ms_o).round(3).tolist())


if __name__ == "__main__":
    which = sys.argv[1:] or ["scan", "cross", "ss2d", "tail", "match", "model", "homography"]
    for w in which:
        globals()["gen_" + w]()
