"""GPU parity against the UNMODIFIED reference running on the same B200 (VERDICT r1 "next round" item 1).  -m gpu.

What runs on the other side (staged by oracle/build_ref.py, loaded by tests/refstage.py; skipped when not staged):
  * the reference's own CUDA kernel `selective_scan_cuda_oflex.fwd` (selective_scan_oflex.cpp:143-231, kernel
    selective_scan_fwd_kernel_oflex.cuh:67-212) compiled for sm_100a from the reference's sources;
  * the reference's `XPoint.forward` (XPoint.py:181-214) at FULL width -- presets E (shipped EXP1) and V
    (vanilla_vmamba_tiny) -- with its own CUDA scan and Triton CrossScan/CrossMerge, fp32 and fp16 autocast;
  * the reference's tail: `box_nms` (torchvision), `nonzero`, `interpolate_descriptors` (grid_sample),
    `get_matches('bfmatcher')` (cv2) -- utils.py:148-238, matching.py:4-36, evaluation.py:281-301.
Tolerances (BASELINE north_star): fp32 rel 1e-4, 16-bit rel 1e-2 (rel-L2 and max/max); keypoints and match pairs
bit-exact on identical fp32 inputs.
"""
import numpy as np
import pytest
import torch

import refstage as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason=R.why_missing())]
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture(scope="module")
def ref():
    ns = R.load()
    assert ns.RS.WITH_SELECTIVESCAN_OFLEX, "the staged reference did not pick up its CUDA extension"
    # the reference's Triton CrossScan must JIT on this box; if it cannot, fall back to its own torch Functions
    try:
        x = torch.randn(1, 4, 8, 8, device=DEV)
        ns.RC.cross_scan_fn(x)
        torch.cuda.synchronize()
        ns.cross_path = "triton"
    except Exception as e:  # pragma: no cover - depends on the box
        R.patch_cross_torch_path(ns)
        ns.cross_path = f"torch ({type(e).__name__})"
    print("reference CrossScan path:", ns.cross_path)
    # SS2D.forwardv0 (preset V) hard-wires backend="mamba" (VMamba.py:312), i.e. the third-party mamba_ssm module
    # `selective_scan_cuda`, which is neither in the reference tree nor in this image.  Its fwd(u, delta, A, B, C, D, z,
    # delta_bias, delta_softplus) is served here by the reference's OWN oflex kernel (same algorithm, csms6s.py:81-85; v0
    # feeds fp32, so the output dtype is the same): a stand-in for an absent dependency, not a change of the reference.
    if not ns.RS.WITH_SELECTIVESCAN_MAMBA:
        import types
        ext = ns.ext
        ns.RS.selective_scan_cuda = types.SimpleNamespace(
            fwd=lambda u, delta, A, B, C, D, z, bias, sp: ext.fwd(u, delta, A, B, C, D, bias, sp, 1, u.dtype == torch.float32))
    return ns


def gpu_rel(x, r):
    """rel-L2 and max-abs/max-ref evaluated on the device in float64, chunked over dim 0."""
    num = den = 0.0
    mx = mr = 0.0
    for xs, rs in zip(x.split(4), r.split(4)):
        d = xs.double() - rs.double()
        num += float((d * d).sum())
        den += float((rs.double() ** 2).sum())
        mx = max(mx, float(d.abs().max()))
        mr = max(mr, float(rs.abs().max()))
    return (num / max(den, 1e-300)) ** 0.5, mx / max(mr, 1e-300)


def scan_inputs(Bt, KD, K, N, L, dtype, seed=0):
    """Input distributions of test_selective_scan.py:414-444 / SURVEY 8d config 2, generated on the device."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    u = torch.randn(Bt, KD, L, device=DEV, generator=g).to(dtype)
    delta = (0.5 * torch.rand(Bt, KD, L, device=DEV, generator=g)).to(dtype)
    A = -0.5 * torch.rand(KD, N, device=DEV, generator=g)
    Bm = torch.randn(Bt, K, N, L, device=DEV, generator=g).to(dtype)
    Cm = torch.randn(Bt, K, N, L, device=DEV, generator=g).to(dtype)
    D = torch.randn(KD, device=DEV, generator=g)
    bias = 0.5 * torch.rand(KD, device=DEV, generator=g)
    return u, delta, A, Bm, Cm, D, bias


SCAN_CASES = [
    # BASELINE configs[1], every row: B 32, K*D 768, N 16, L 20480
    ("config2_fp32", 32, 768, 4, 16, 20480, torch.float32, True, 1e-4),
    ("config2_bf16_f32out", 32, 768, 4, 16, 20480, torch.bfloat16, True, 1e-2),
    ("config2_bf16_bf16out", 32, 768, 4, 16, 20480, torch.bfloat16, False, 1e-2),
    ("config2_fp16_f32out", 32, 768, 4, 16, 20480, torch.float16, True, 1e-2),
    # preset E stage shapes at 512x640, N = 1
    ("E_stage0_fp32", 8, 384, 4, 1, 20480, torch.float32, True, 1e-4),
    ("E_stage1_fp32", 8, 768, 4, 1, 5120, torch.float32, True, 1e-4),
    ("E_stage2_fp32", 8, 1536, 4, 1, 1280, torch.float32, True, 1e-4),
    ("E_stage3_fp32", 8, 3072, 4, 1, 320, torch.float32, True, 1e-4),
    ("E_stage0_fp16", 8, 384, 4, 1, 20480, torch.float16, True, 1e-2),
    ("E_stage1_fp16", 8, 768, 4, 1, 5120, torch.float16, True, 1e-2),
    ("E_stage2_fp16", 8, 1536, 4, 1, 1280, torch.float16, True, 1e-2),
    ("E_stage3_fp16", 8, 3072, 4, 1, 320, torch.float16, True, 1e-2),
    ("E_stage0_bf16", 8, 384, 4, 1, 20480, torch.bfloat16, True, 1e-2),
    # preset V stage shapes (N = 16, d_inner = 2C); v0 feeds the scan fp32 (VMamba.py:341)
    ("V_stage1_fp32", 4, 1536, 4, 16, 5120, torch.float32, True, 1e-4),
    ("V_stage2_fp32", 4, 3072, 4, 16, 1280, torch.float32, True, 1e-4),
    ("V_stage3_fp32", 4, 6144, 4, 16, 320, torch.float32, True, 1e-4),
    # high-res stage 0 (config 5): L = 81 920
    ("Q_stage0_fp16", 2, 384, 4, 1, 81920, torch.float16, True, 1e-2),
    # other state sizes / odd lengths the reference kernel accepts
    ("n8_L1011", 2, 64, 2, 8, 1011, torch.float32, True, 1e-4),
    ("n4_L4096", 2, 96, 4, 4, 4096, torch.float32, True, 1e-4),
    ("n2_L777", 3, 48, 4, 2, 777, torch.float32, True, 1e-4),
]


@pytest.mark.parametrize("name,Bt,KD,K,N,L,dtype,oflex,tol", SCAN_CASES, ids=[c[0] for c in SCAN_CASES])
def test_scan_vs_reference_cuda_kernel(ref, name, Bt, KD, K, N, L, dtype, oflex, tol):
    """xp_selective_scan_fwd against selective_scan_cuda_oflex.fwd on identical device tensors, every row."""
    import xpoint_b200 as X
    u, delta, A, Bm, Cm, D, bias = scan_inputs(Bt, KD, K, N, L, dtype)
    out_ref, _x = ref.ext.fwd(u, delta, A, Bm, Cm, D, bias, True, 1, oflex)
    got = X.selective_scan_fn(u, delta, A, Bm, Cm, D, bias, True, oflex)
    torch.cuda.synchronize()
    assert got.dtype == out_ref.dtype and got.shape == out_ref.shape
    l2, mx = gpu_rel(got, out_ref)
    assert l2 <= tol and mx <= tol, f"{name}: rel_l2={l2:.3e} max/max={mx:.3e} > {tol:.0e}"
    # the reference's own allclose bounds (test_selective_scan.py:401-403)
    rtol, atol = {torch.float32: (6e-4, 2e-3), torch.float16: (3e-3, 5e-3), torch.bfloat16: (3e-2, 5e-2)}[dtype]
    if out_ref.dtype == torch.float32:
        bad = ((got - out_ref).abs() > atol + rtol * out_ref.abs()).float().mean().item()
        assert bad <= 1e-6, f"{name}: {bad:.2e} of the elements outside the reference's rtol/atol"


def test_scan_via_reference_dispatcher(ref):
    """The reference's own dispatcher (csms6s.py:112-126) with OUR module bound where it imports its extension
     gives the same result as with the reference's kernel."""
    import xpoint_b200 as X
    u, delta, A, Bm, Cm, D, bias = scan_inputs(2, 96, 4, 16, 2048, torch.float32, seed=3)
    RS = ref.RS
    y_ref = RS.selective_scan_fn(u, delta, A, Bm, Cm, D, bias, True, True)
    saved = RS.selective_scan_cuda_oflex
    try:
        RS.selective_scan_cuda_oflex = X.selective_scan_cuda_oflex
        y_ours = RS.selective_scan_fn(u, delta, A, Bm, Cm, D, bias, True, True)
    finally:
        RS.selective_scan_cuda_oflex = saved
    l2, mx = gpu_rel(y_ours, y_ref)
    assert l2 <= 1e-4 and mx <= 1e-4, (l2, mx)


# ------------------------------------------------------------------------------------------ whole model, full width
def _pair(Bn, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(Bn, 1, H, W, generator=g).to(DEV), torch.rand(Bn, 1, H, W, generator=g).to(DEV)


def _run_ref(net, o, t):
    with torch.no_grad():
        po, pt, hm = net({"optical": {"image": o}, "thermal": {"image": t}})
    assert hm is None
    return po, pt


MODEL_CASES = [("E", 256, 256, 2), ("V", 256, 256, 2), ("E", 512, 640, 2), ("E", 1024, 1280, 1)]   # configs[0], [2], [4]


@pytest.mark.parametrize("preset,H,W,Bn", MODEL_CASES, ids=[f"{c[0]}_{c[1]}x{c[2]}" for c in MODEL_CASES])
def test_full_width_model_vs_reference(ref, preset, H, W, Bn):
    """BASELINE configs[0] (256x256, presets V and E) and configs[2] (512x640, preset E), FULL width: this repo's XPoint
    with the reference's state_dict against the reference's XPoint.forward on the same GPU.  fp32 at 1e-4; fp16 autocast
    (the reference's mixed_precision path) at 1e-2; both on encoder_output / prob / desc."""
    import xpoint_b200 as X
    o, t = _pair(Bn, H, W, seed=7)
    ref32 = None
    for mixed, tol in ((False, 1e-4), (True, 1e-2)):
        rnet = R.randomise_stats(R.build_xpoint(ref, preset, mixed_precision=mixed, height=H, width=W)).to(DEV)
        ours = X.XPoint({"takes_pair": True, "mixed_precision": mixed, "use_attention": {"preset": preset}})
        ours.load_state_dict(rnet.state_dict(), strict=True)
        ours = ours.to(DEV).eval()
        rpo, rpt = _run_ref(rnet, o, t)
        if not mixed:
            ref32 = (rpo, rpt)
        with torch.no_grad():
            po, pt, hm = ours({"optical": {"image": o}, "thermal": {"image": t}})
            bo, bt = ours.forward_pair_batched(o, t)
        assert hm is None and po["logits"] is None
        for mine, bat, theirs, truth, s in ((po, bo, rpo, ref32[0], "optical"), (pt, bt, rpt, ref32[1], "thermal")):
            for k in ("encoder_output", "prob", "desc"):
                assert mine[k].shape == theirs[k].shape, (k, mine[k].shape, theirs[k].shape)
                # the reference's fp32 result is the anchor for both precisions ("agree with the reference's torch path
                # within fp32 rel 1e-4, 16-bit rel 1e-2"); the reference's own fp16 result is compared as well whenever it is
                # itself within tolerance of its fp32 result (on B200 cuDNN's fp16 depth-wise convolution, which the
                # reference calls at VMamba.py:651, is wrong for some shapes -- DESIGN.md section 2 -- so it may not be)
                ref_l2, ref_mx = gpu_rel(theirs[k].float(), truth[k].float())
                ref_ok = max(ref_l2, ref_mx) <= tol
                if mixed:
                    print(f"{preset} {H}x{W} {k} {s}: reference fp16 vs reference fp32 rel_l2={ref_l2:.3e} max/max={ref_mx:.3e}"
                          + ("" if ref_ok else "  <-- the reference's own fp16 path is off on this box"))
                for tag, val in (("forward", mine[k]), ("batched", bat[k])):
                    l2, mx = gpu_rel(val.float(), truth[k].float())
                    assert l2 <= tol and mx <= tol, \
                        f"{preset} {H}x{W} mixed={mixed} {tag} {k} {s} vs reference fp32: rel_l2={l2:.3e} max/max={mx:.3e} > {tol:.0e}"
                    if mixed and ref_ok:
                        l2, mx = gpu_rel(val.float(), theirs[k].float())
                        assert l2 <= tol and mx <= tol, \
                            f"{preset} {H}x{W} {tag} {k} {s} vs reference fp16: rel_l2={l2:.3e} max/max={mx:.3e} > {tol:.0e}"
        del rnet, ours
        torch.cuda.empty_cache()


def test_vssm_encoder_blocks_vs_reference(ref):
    """One full-width SS2D block per preset and stage shape of 512x640 against the reference's SS2D module on the GPU
    (its CUDA scan + Triton CrossScan/Merge): fp32 1e-4."""
    import xpoint_b200 as X
    for kw, C, H, W in ((dict(d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False), 96, 128, 160),
                        (dict(d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False), 768, 16, 20),
                        (dict(d_state=16, ssm_ratio=2.0, forward_type="v0"), 96, 128, 160),
                        (dict(d_state=16, ssm_ratio=2.0, forward_type="v0"), 384, 32, 40)):
        torch.manual_seed(0)
        rm = ref.RV.SS2D(d_model=C, **kw).eval()
        with torch.no_grad():
            rm.A_logs.copy_(torch.log(0.5 * torch.rand_like(rm.A_logs) + 0.05))
            rm.Ds.copy_(torch.randn_like(rm.Ds))
        m = X.SS2D(d_model=C, **kw)
        m.load_state_dict(rm.state_dict(), strict=True)
        rm, m = rm.to(DEV), m.to(DEV).eval()
        x = torch.randn(2, H, W, C, device=DEV)
        with torch.no_grad():
            y_ref, y = rm(x), m(x)
        l2, mx = gpu_rel(y, y_ref)
        assert l2 <= 1e-4 and mx <= 1e-4, f"SS2D {kw['forward_type']} C={C} {H}x{W}: rel_l2={l2:.3e} max/max={mx:.3e}"


# ------------------------------------------------------------------------------------------ tail on the reference's tensors
def _distinct_scores(Bn, H, W, seed):
    """Score maps with pairwise distinct values (a permutation of a grid, ** 6): greedy NMS has no ties to break, so
    torchvision's unspecified tie order cannot matter."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(Bn):
        p = (torch.randperm(H * W, generator=g).double() + 0.5) / (H * W)
        out.append((p ** 6).float().reshape(1, H, W))
    s = torch.stack(out)
    assert all(torch.unique(s[b]).numel() == H * W for b in range(Bn))
    return s


def _ref_tail(ref, prob, desc, k, thr=0.015):
    """evaluation.py:250-301 for a batch: box_nms -> nonzero -> interpolate_descriptors; per image lists."""
    H, W = prob.shape[-2:]
    nms = ref.utils.box_nms(prob, 8, thr, keep_top_k=k, on_cpu=False)
    kps, descs = [], []
    for b in range(prob.shape[0]):
        kp = torch.nonzero(nms[b].squeeze() > thr)
        kps.append(kp)
        descs.append(ref.utils.interpolate_descriptors(kp, desc[b], H, W))
    return nms, kps, descs


@pytest.mark.parametrize("H,W,k", [(256, 256, 512), (512, 640, 4096), (1024, 1280, 16384)])   # configs[0], [2], [4]
def test_tail_on_reference_tensors_bit_exact(ref, H, W, k):
    """Keypoints and match pairs bit-exact at top-k when both sides are fed the same fp32 score / descriptor tensors:
    (i) the reference model's own prob/desc for a pair (full-width preset E, fp32), (ii) distinct-valued synthetic score
    maps that leave exactly k survivors.  Reference side: box_nms (torchvision on the GPU), nonzero,
    interpolate_descriptors (grid_sample), cv2.BFMatcher(crossCheck)."""
    import xpoint_b200 as X
    Bn = 2 if H < 1024 else 1            # cv2.BFMatcher on 16 384 x 16 384 descriptors takes ~10 s per pair
    rnet = R.randomise_stats(R.build_xpoint(ref, "E", height=H, width=W)).to(DEV)
    o, t = _pair(Bn, H, W, seed=11)
    rpo, rpt = _run_ref(rnet, o, t)
    g = torch.Generator().manual_seed(5)
    syn_prob = _distinct_scores(2 * Bn, H, W, seed=5).to(DEV)
    syn_desc = torch.nn.functional.normalize(torch.randn(2 * Bn, 256, H // 8, W // 8, generator=g), dim=1).to(DEV)
    cases = {"model": (torch.cat([rpo["prob"], rpt["prob"]]).float(), torch.cat([rpo["desc"], rpt["desc"]]).float()),
             "synthetic": (syn_prob, syn_desc)}
    for name, (prob, desc) in cases.items():
        # torchvision's GPU nms needs N^2 / 8 bytes of mask for N boxes above the threshold: at 1024x1280 the shipped 0.015 lets
        # half of a random-weight score map through (650 k boxes, 53 GB), so that size runs at the map's 0.9 quantile instead
        thr = 0.015 if H < 1024 else float(torch.quantile(prob[0].flatten().float(), 0.9))
        pipe = X.PairPipeline(None, nms=8, detection_threshold=thr, keep_top_k=k, use_tensor_cores=True)
        r = pipe.tail(prob[:Bn], prob[Bn:], desc[:Bn], desc[Bn:])
        nms_ref, kps_ref, d_ref = _ref_tail(ref, prob, desc, k, thr)
        # the dense NMS map itself
        nms_ours = X.box_nms(prob, 8, thr, keep_top_k=k)
        assert torch.equal(nms_ours, nms_ref), f"{name}: box_nms map differs from the reference's"
        n_all = torch.cat([r.n_optical, r.n_thermal]).tolist()
        kp_all = torch.cat([r.kp_optical, r.kp_thermal])
        d_all = torch.cat([r.desc_optical, r.desc_thermal])
        for b in range(2 * Bn):
            n = len(kps_ref[b])
            assert n_all[b] == n, f"{name}: image {b} has {n_all[b]} keypoints, reference {n}"
            if name == "synthetic" and H < 1024:
                assert n == k
            assert torch.equal(kp_all[b, :n].long(), kps_ref[b]), f"{name}: keypoints of image {b} differ"
            np.testing.assert_allclose(d_all[b, :n].cpu().numpy(), d_ref[b].cpu().numpy(), rtol=0, atol=5e-6 if H < 1024 else 1e-5)   # ATen grid_sample on the GPU contracts differently from its CPU kernel (2e-6 there); 1 of 4.2 M values at 5.3e-6 for 1280-wide maps
        for b in range(Bn):
            n1, n2 = n_all[b], n_all[Bn + b]
            d1, d2 = d_all[b, :n1].cpu().numpy(), d_all[Bn + b, :n2].cpu().numpy()
            # identical fp32 descriptors to both matchers (matching.py:4-36)
            m = ref.utils.get_matches(d1, d2, "bfmatcher", False, crossCheck=True)
            want = [(x.queryIdx, x.trainIdx) for x in m]
            idx = r.match_idx[b].cpu().numpy()
            got = [(i, int(j)) for i, j in enumerate(idx[:n1]) if j >= 0]
            # SURVEY C.13 guard: rows whose best-vs-second gap is below fp32 summation noise may legitimately differ
            dm = torch.cdist(torch.from_numpy(d1).double().to(DEV), torch.from_numpy(d2).double().to(DEV))
            if min(n1, n2) >= 2:
                s1, s2 = dm.topk(2, 1, largest=False).values, dm.topk(2, 0, largest=False).values
                amb_q = set(torch.nonzero((s1[:, 1] - s1[:, 0]) < 1e-5).flatten().tolist())
                amb_t = set(torch.nonzero((s2[1] - s2[0]) < 1e-5).flatten().tolist())
            else:
                amb_q, amb_t = set(), set()
            f = lambda pairs: [(i, j) for i, j in pairs if i not in amb_q and j not in amb_t]  # noqa: E731
            assert len(amb_q) + len(amb_t) <= 0.01 * max(n1 + n2, 1)
            assert f(got) == f(want), f"{name}: match pairs of pair {b} differ from cv2.BFMatcher"
            assert int(r.n_matches[b]) == len(got)
