"""SS2D block, XPoint model and pair pipeline vs vectors produced by the real reference (tests/golden).  -m gpu."""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
FP32_REL = 1e-4

TINY = dict(
    E=dict(DEPTHS=[1, 1, 1, 1], DOWNSAMPLE="v3", EMBED_DIM=16, MLP_RATIO=4.0, PATCHEMBED="v2", SSM_CONV=3,
           SSM_CONV_BIAS=False, SSM_DT_RANK="auto", SSM_D_STATE=1, SSM_FORWARDTYPE="v05_noz", SSM_RATIO=1.0),
    V=dict(DEPTHS=[1, 1, 2, 1], DOWNSAMPLE="v1", EMBED_DIM=16, MLP_RATIO=0.0, PATCHEMBED="v1", SSM_CONV=3,
           SSM_CONV_BIAS=True, SSM_DT_RANK="auto", SSM_D_STATE=4, SSM_FORWARDTYPE="v0", SSM_RATIO=2.0,
           SSM_INIT="v0", NORM_LAYER="ln"),
)   # same dicts as tests/golden/make_golden.py


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _sd(g):
    return {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}


@pytest.mark.parametrize("name,kw", [
    ("ss2d_v0", dict(d_model=16, d_state=4, ssm_ratio=2.0, forward_type="v0")),
    ("ss2d_v05_noz", dict(d_model=16, d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False)),
    ("ss2d_v05", dict(d_model=16, d_state=2, ssm_ratio=2.0, forward_type="v05")),
])
def test_ss2d_block_golden(name, kw):
    import xpoint_b200 as X
    g = golden(name)
    m = X.SS2D(**kw)
    m.load_state_dict(_sd(g), strict=True)      # reference state_dict loads unchanged
    m = m.to(DEV).eval()
    with torch.no_grad():
        y = m(torch.from_numpy(g["x"]).to(DEV))
    assert_close(y.cpu().numpy(), g["y"], FP32_REL, name)
    # 12 x 20 tokens: the shape class of the copy-free path (H, W multiples of 4; d_state 1, 2, 4, 8, 16)
    assert m._use_fused(12, 20, torch.float32)
    with torch.no_grad():
        y4 = m(torch.from_numpy(g["x4"]).to(DEV))
        m.disable_fused = True
        y4u = m(torch.from_numpy(g["x4"]).to(DEV))
    assert_close(y4.cpu().numpy(), g["y4"], FP32_REL, name + " 12x20 (fused where supported)")
    assert_close(y4u.cpu().numpy(), g["y4"], FP32_REL, name + " 12x20 (CrossScan/CrossMerge kernels)")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("ft,N,ratio", [("v05_noz", 1, 1.0), ("v05", 2, 2.0), ("v05_nozact", 1, 2.0), ("v05_oact", 1, 1.0)])
def test_ss2d_fused_matches_unfused(dtype, ft, N, ratio):
    """Copy-free path vs the CrossScan -> scan -> CrossMerge path of the same module at an XPoint stage shape."""
    import xpoint_b200 as X
    torch.manual_seed(0)
    m = X.SS2D(d_model=48, d_state=N, ssm_ratio=ratio, forward_type=ft, conv_bias=(ft != "v05_noz")).to(DEV).eval()
    x = torch.randn(2, 32, 40, 48, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
        assert m._use_fused(32, 40, dtype)
        y = m(x)
        m.disable_fused = True
        yu = m(x)
    assert y.dtype == yu.dtype
    assert_close(y.float().cpu().numpy(), yu.float().cpu().numpy(), 2e-5 if dtype == torch.float32 else 1e-2, f"{ft} {dtype}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("C", [96, 192, 384])
def test_ss2d_fused_dt_proj_matches_materialised_delta(dtype, C):
    """dt ranks 6 / 12 form delta inside the scan (xp_scan_args.dt_weight), rank 24 keeps the
    materialised delta; both must agree with the dt_proj -> scan path of the same module."""
    import xpoint_b200 as X
    torch.manual_seed(1)
    m = X.SS2D(d_model=C, d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False).to(DEV).eval()
    x = torch.randn(2, 32, 40, C, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
        m.fuse_dt_proj = True
        y = m(x)
        m.fuse_dt_proj = False
        ym = m(x)
    assert_close(y.float().cpu().numpy(), ym.float().cpu().numpy(), 2e-5 if dtype == torch.float32 else 1e-2, f"C={C} {dtype}")


@pytest.mark.parametrize("H,W,C", [(128, 160, 96), (64, 80, 192), (32, 40, 384), (16, 20, 768)])
def test_ss2d_fp16_autocast_at_xpoint_stage_shapes(H, W, C):
    """The four stage shapes of preset E at 512x640: fp16 autocast (the reference's
    mixed_precision path, XPoint.py:182) against fp32, on both SS2D paths.  Catches shape-dependent library
    behaviour as well: cuDNN's fp16 depth-wise conv was wrong at (16, 20, 768), which is why SS2D runs its own."""
    import xpoint_b200 as X
    torch.manual_seed(0)
    m = X.SS2D(d_model=C, d_state=1, ssm_ratio=1.0, forward_type="v05_noz", conv_bias=False).to(DEV).eval()
    x = torch.randn(2, H, W, C, device=DEV)
    with torch.no_grad():
        ref = m(x)
        for fused in (True, False):
            m.disable_fused = not fused
            with torch.autocast("cuda", dtype=torch.float16):
                y = m(x)
            assert y.dtype == torch.float16
            assert_close(y.float().cpu().numpy(), ref.cpu().numpy(), 1e-2, f"fp16 fused={fused} {H}x{W}x{C}")


@pytest.mark.parametrize("tag", ["E", "V"])
def test_xpoint_tiny_golden(tag):
    """BASELINE config 1 at reduced width: the reference XPoint forward on a synthetic pair, CPU fp32."""
    import xpoint_b200 as X
    g = golden("xpoint_tiny_" + tag)
    net = X.XPoint({"takes_pair": True,
                    "use_attention": {"model_parameters": {"MODEL": {"DROP_PATH_RATE": 0.2, "VSSM": TINY[tag]}}}})
    net.load_state_dict(_sd(g), strict=True)
    net = net.to(DEV).eval()
    data = {"optical": {"image": torch.from_numpy(g["img_optical"]).to(DEV)},
            "thermal": {"image": torch.from_numpy(g["img_thermal"]).to(DEV)}}
    with torch.no_grad():
        po, pt, hm = net(data)
        bo, bt = net.forward_pair_batched(data["optical"]["image"], data["thermal"]["image"])
    assert hm is None and po["logits"] is None
    for pred, bat, s in ((po, bo, "optical"), (pt, bt, "thermal")):
        assert_close(pred["encoder_output"].cpu().numpy(), g["enc_" + s], FP32_REL, f"{tag} enc {s}")
        assert_close(pred["prob"].cpu().numpy(), g["prob_" + s], FP32_REL, f"{tag} prob {s}")
        assert_close(pred["desc"].cpu().numpy(), g["desc_" + s], FP32_REL, f"{tag} desc {s}")
        assert_close(bat["prob"].cpu().numpy(), g["prob_" + s], FP32_REL, f"{tag} batched prob {s}")
        assert pred["prob"].shape == (2, 1, 64, 96) and pred["desc"].shape == (2, 256, 8, 12)


def test_pipeline_tail_vs_oracle():
    """PairPipeline.tail on synthetic score maps / descriptor maps (rand**6 maps give >= k survivors)
    against the oracle run per image: keypoints bit-exact, descriptors 2e-6, matches bit-exact."""
    import xpoint_b200 as X
    from oracle import oracle as O
    g = torch.Generator().manual_seed(0)
    B, H, W, k = 2, 256, 320, 512
    prob = torch.rand(2 * B, 1, H, W, generator=g) ** 6
    desc = torch.nn.functional.normalize(torch.randn(2 * B, 256, H // 8, W // 8, generator=g), dim=1)
    pipe = X.PairPipeline(None, nms=8, detection_threshold=0.015, keep_top_k=k, use_tensor_cores=True)
    r = pipe.tail(prob[:B].to(DEV), prob[B:].to(DEV), desc[:B].to(DEV), desc[B:].to(DEV))
    for b in range(B):
        kps, ds = [], []
        for img, kp_t, n_t, d_t in ((b, r.kp_optical, r.n_optical, r.desc_optical), (B + b, r.kp_thermal, r.n_thermal, r.desc_thermal)):
            nms = O.box_nms(prob[img, 0].numpy(), 8, 0.015, keep_top_k=k)
            kp = O.extract_keypoints(nms, 0.015)
            assert int(n_t[b]) == len(kp) == k
            assert np.array_equal(kp_t[b, :k].cpu().numpy().astype(np.int64), kp)
            d = O.interpolate_descriptors(kp, desc[img].numpy(), H, W)
            np.testing.assert_allclose(d_t[b, :k].cpu().numpy(), d, rtol=0, atol=2e-6)
            kps.append(kp)
            ds.append(d_t[b, :k].cpu().numpy())       # match on the SAME descriptors the GPU matched
        q, t, _, gap = O.mnn_match(ds[0], ds[1], return_gap=True)
        idx = r.match_idx[b].cpu().numpy()
        got = [(i, int(j)) for i, j in enumerate(idx) if j >= 0]
        if gap.min() > 1e-5:
            assert got == list(zip(q.tolist(), t.tolist()))
        assert int(r.n_matches[b]) == len(got)


def test_config5_highres_pair_runs_and_agrees():
    """BASELINE configs[4]: one 1024x1280 pair (stage-0 scan length 81 920), preset E, top-16384 keypoints.  fp16 autocast
    on the copy-free path against fp32 on the CrossScan/CrossMerge path; the tail must return full keypoint sets."""
    import xpoint_b200 as X
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    o = torch.rand(1, 1, 1024, 1280, generator=g).to(DEV)
    t = torch.rand(1, 1, 1024, 1280, generator=g).to(DEV)
    outs = {}
    for mixed in (False, True):
        torch.manual_seed(0)
        net = X.XPoint({"takes_pair": True, "mixed_precision": mixed, "use_attention": {"preset": "E"}}).to(DEV).eval()
        for m in net.modules():
            if isinstance(m, X.SS2D):
                m.disable_fused = not mixed
        with torch.no_grad():
            po, pt = net.forward_pair_batched(o, t)
        outs[mixed] = po
        if mixed:
            pipe = X.PairPipeline(net, keep_top_k=16384)
            r = pipe(o, t)
            assert r.kp_optical.shape == (1, 16384, 2) and 0 < int(r.n_optical[0]) <= 16384
            assert 0 <= int(r.n_matches[0]) <= int(r.n_optical[0])
    for k in ("encoder_output", "prob", "desc"):
        assert outs[True][k].shape == outs[False][k].shape
        assert_close(outs[True][k].float().cpu().numpy(), outs[False][k].float().cpu().numpy(), 1e-2, "config 5 " + k)
    assert outs[True]["prob"].shape == (1, 1, 1024, 1280) and outs[True]["desc"].shape == (1, 256, 128, 160)


def test_graphed_pipeline_replays_the_eager_step():
    """PairPipeline.capture: the CUDA-graph replay returns exactly what the eager step returns, also
    for new images loaded into the graph's static inputs."""
    import xpoint_b200 as X
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).to(DEV).eval()
    pipe = X.PairPipeline(net, keep_top_k=512, estimate_homography=True)
    g = torch.Generator().manual_seed(1)
    o1, t1 = torch.rand(2, 1, 128, 160, generator=g).to(DEV), torch.rand(2, 1, 128, 160, generator=g).to(DEV)
    o2, t2 = torch.rand(2, 1, 128, 160, generator=g).to(DEV), torch.rand(2, 1, 128, 160, generator=g).to(DEV)
    graphed = pipe.capture(o1, t1)
    assert isinstance(graphed, X.GraphedPairPipeline)
    for o, t in ((o1, t1), (o2, t2), (o1, t1)):
        ref = pipe(o, t)
        got = graphed(o, t)
        torch.cuda.synchronize()
        for name, a, b in zip(ref._fields, ref, got):
            if name.startswith("kp_"):          # keypoint rows past the count are undefined (uninitialised memory)
                n = ref.n_optical if name == "kp_optical" else ref.n_thermal
                for i in range(a.shape[0]):
                    assert torch.equal(a[i, : int(n[i])], b[i, : int(n[i])]), f"graph replay differs in {name}"
            else:
                assert torch.equal(a, b), f"graph replay differs from the eager step in {name}"
    with pytest.raises(RuntimeError):
        pipe.capture(o1.cpu(), t1.cpu())


@pytest.mark.parametrize("reflection", [True, False])
def test_heads_respect_reflection_pad(reflection):
    """with reflection_pad=False the reference builds nn.ZeroPad2d (XPoint.py:101-104); the fused eval heads
    must pad the same way as the module-by-module path."""
    import xpoint_b200 as X
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": False, "reflection_pad": reflection, "use_attention": {"preset": "E"}}).to(DEV).eval()
    img = torch.rand(2, 1, 64, 96, device=DEV)
    with torch.no_grad():
        assert net._heads_foldable()
        fused = net({"image": img})
        net._heads_foldable = lambda: False
        plain = net({"image": img})
    assert type(net.detector_head_convolutions[0]).__name__ == ("ReflectionPad2d" if reflection else "ZeroPad2d")
    for k in ("prob", "desc", "encoder_output"):
        assert_close(fused[k].cpu().numpy(), plain[k].cpu().numpy(), 2e-5, f"reflection_pad={reflection} {k}")


def test_pipeline_topk_zero_means_no_cap():
    """keep_top_k = 0 is the reference's "no cap" (utils.py:179 `if keep_top_k > 0`, configs/cipdp.yaml:55 topk: 0)."""
    import xpoint_b200 as X
    from oracle import oracle as O
    g = torch.Generator().manual_seed(3)
    B, H, W = 1, 128, 160
    prob = torch.rand(2 * B, 1, H, W, generator=g) ** 6
    desc = torch.nn.functional.normalize(torch.randn(2 * B, 256, H // 8, W // 8, generator=g), dim=1)
    mask = (torch.rand(2 * B, 1, H, W, generator=g) > 0.2).float()
    pipe = X.PairPipeline(None, nms=8, detection_threshold=0.015, keep_top_k=0)
    for m in (None, mask):
        r = pipe.tail(prob[:B].to(DEV), prob[B:].to(DEV), desc[:B].to(DEV), desc[B:].to(DEV),
                      valid_mask_o=None if m is None else m[:B].to(DEV), valid_mask_t=None if m is None else m[B:].to(DEV))
        for img, kp_t, n_t in ((0, r.kp_optical, r.n_optical), (1, r.kp_thermal, r.n_thermal)):
            p = prob[img, 0] if m is None else prob[img, 0] * m[img, 0]
            kp = O.extract_keypoints(O.box_nms(p.numpy(), 8, 0.015, keep_top_k=0), 0.015)
            assert len(kp) > 100 and int(n_t[0]) == len(kp) <= pipe.capacity(H, W)
            assert np.array_equal(kp_t[0, :len(kp)].cpu().numpy().astype(np.int64), kp)
        assert int(r.n_matches[0]) > 0
    with pytest.raises(ValueError):
        X.PairPipeline(None, keep_top_k=-1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_ss2d_fused_core_matches_default_path(dtype):
    """SS2D with the fused core (xp_ss2d_core + xp_ss2d_plane_norm, `use_core`) against the default copy-free path
    (xp_selective_scan_fwd + xp_ss2d_merge_norm) and the CrossScan/CrossMerge kernels, at two stage shapes."""
    import xpoint_b200 as X
    for C, H, W, N, ft in ((96, 64, 80, 1, "v05_noz"), (48, 32, 40, 2, "v05")):
        torch.manual_seed(0)
        m = X.SS2D(d_model=C, d_state=N, ssm_ratio=1.0, forward_type=ft, conv_bias=False).to(DEV).eval()
        x = torch.randn(2, H, W, C, device=DEV)
        with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
            y0 = m(x)
            m.use_core = True
            y1 = m(x)
            m.use_core = False
            m.disable_fused = True
            y2 = m(x)
        tol = 2e-5 if dtype == torch.float32 else 1e-2
        assert_close(y1.float().cpu().numpy(), y0.float().cpu().numpy(), tol, f"fused core vs default C={C} {dtype}")
        assert_close(y1.float().cpu().numpy(), y2.float().cpu().numpy(), tol, f"fused core vs CrossScan path C={C} {dtype}")


def test_sharded_pair_pipeline_matches_direct_calls():
    """ShardedPairPipeline: a host list of pairs streamed through pinned double buffers and a CUDA graph,
    padded tail batch, results gathered on the host in input order == the same pairs through PairPipeline directly."""
    import xpoint_b200 as X
    from xpoint_b200.pipeline import ShardedPairPipeline
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).to(DEV).eval()
    g = torch.Generator().manual_seed(2)
    o = torch.rand(7, 1, 128, 160, generator=g).pin_memory()
    t = torch.rand(7, 1, 128, 160, generator=g).pin_memory()
    sp = ShardedPairPipeline(net, devices=[torch.device("cuda", 0)], batch=3, keep_top_k=300)
    res = sp.run(o, t)
    pipe = X.PairPipeline(net, keep_top_k=300)
    for lo in range(0, 7, 3):
        r = pipe(o[lo:lo + 3].to(DEV), t[lo:lo + 3].to(DEV))
        n = min(3, 7 - lo)
        assert torch.equal(res["n_optical"][lo:lo + n], r.n_optical.cpu()[:n])
        assert torch.equal(res["n_matches"][lo:lo + n], r.n_matches.cpu()[:n])
        for b in range(n):
            k = int(r.n_optical[b])
            assert torch.equal(res["kp_optical"][lo + b, :k], r.kp_optical[b, :k].cpu())
            assert torch.equal(res["match_idx"][lo + b, :k], r.match_idx[b, :k].cpu())
    assert res["kp_optical"].shape[0] == 7


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_pair_pipeline_two_devices_in_process():
    """ShardedPairPipeline over two GPUs of one process (worker thread + weight replica + CUDA graph per device) returns the
    same keypoints / matches as a single device, in input order."""
    import xpoint_b200 as X
    from xpoint_b200.pipeline import ShardedPairPipeline
    torch.manual_seed(0)
    net = X.XPoint({"takes_pair": True, "mixed_precision": True, "use_attention": {"preset": "E"}}).to("cuda:0").eval()
    g = torch.Generator().manual_seed(3)
    o = torch.rand(10, 1, 128, 160, generator=g).pin_memory()
    t = torch.rand(10, 1, 128, 160, generator=g).pin_memory()
    one = ShardedPairPipeline(net, devices=[torch.device("cuda", 0)], batch=4, keep_top_k=256).run(o, t)
    two = ShardedPairPipeline(net, devices=[torch.device("cuda", 0), torch.device("cuda", 1)], batch=4, keep_top_k=256).run(o, t)
    assert torch.equal(one["n_optical"], two["n_optical"]) and torch.equal(one["n_matches"], two["n_matches"])
    for b in range(10):
        k = int(one["n_optical"][b])
        assert torch.equal(one["kp_optical"][b, :k], two["kp_optical"][b, :k])
        assert torch.equal(one["match_idx"][b, :k], two["match_idx"][b, :k])
