"""Post-processing tail parity: detector post, NMS/top-k/keypoints (bit-exact), sampling, MNN matching (bit-exact)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _imports():
    import xpoint_b200 as X
    from oracle import oracle as O
    return X, O


def test_heads_golden():
    X, _ = _imports()
    g = golden("heads")
    prob = X.detector_post(torch.from_numpy(g["logits"]).to(DEV))
    np.testing.assert_allclose(prob.cpu().numpy(), g["prob"], rtol=1e-5, atol=1e-7)
    d = X.normalize_descriptors(torch.from_numpy(g["desc_raw"]).to(DEV))
    np.testing.assert_allclose(d.cpu().numpy(), g["desc"], rtol=1e-5, atol=1e-7)
    d, dcl = X.normalize_descriptors(torch.from_numpy(g["desc_raw"]).to(DEV), channel_last_copy=True)
    assert torch.equal(dcl, d.permute(0, 2, 3, 1).contiguous())
    # 16-bit logits are widened first, like `.to(torch.float)` in XPoint.py:349
    p16 = X.detector_post(torch.from_numpy(g["logits"]).to(DEV).half())
    ref16 = torch.nn.PixelShuffle(8)(torch.softmax(torch.from_numpy(g["logits"]).half().float(), 1)[:, :-1])
    np.testing.assert_allclose(p16.cpu().numpy(), ref16.numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("ld", [65, 72, 80])
def test_channel_last_head_kernels(dtype, ld):
    """xp_detector_post_cl / xp_l2_normalize_cl consume the heads' GEMM outputs as (cells, C) rows: same results as the
    channel-first kernels (which are pinned to the reference goldens above) on the permuted tensors."""
    import xpoint_b200.postprocess as P
    g = torch.Generator().manual_seed(ld)
    B, Hc, Wc = 3, 9, 13
    rows = torch.randn(B * Hc * Wc, ld, generator=g).to(dtype).to(DEV)
    cf = rows[:, :65].reshape(B, Hc, Wc, 65).permute(0, 3, 1, 2).contiguous()
    assert torch.equal(P.detector_post_rows(rows, B, Hc, Wc, 8), P.detector_post(cf, 8))
    assert torch.equal(P.detector_post_rows(rows[:, 1:], B, Hc, Wc, 8) if ld > 65 else P.detector_post_rows(rows, B, Hc, Wc, 8),
                       P.detector_post(rows[:, 1:66].reshape(B, Hc, Wc, 65).permute(0, 3, 1, 2).contiguous(), 8)
                       if ld > 65 else P.detector_post(cf, 8))                          # unaligned rows take the scalar loads
    d = torch.randn(B * Hc * Wc, 256, generator=g).to(dtype).to(DEV)
    dcf = d.reshape(B, Hc, Wc, 256).permute(0, 3, 1, 2).contiguous()
    ref, ref_cl = P.normalize_descriptors(dcf, channel_last_copy=True)
    out, out_cl = P.normalize_descriptor_rows(d, B, Hc, Wc)
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-7) and torch.allclose(out_cl, ref_cl, rtol=1e-6, atol=1e-7)
    assert torch.equal(out_cl, out.permute(0, 2, 3, 1).contiguous())
    none, only_cl = P.normalize_descriptor_rows(d, B, Hc, Wc, want_channel_first=False)
    assert none is None and torch.equal(only_cl, out_cl)


@pytest.mark.parametrize("CO,pdt,odt", [(48, torch.float16, torch.float16), (8, torch.float32, torch.float32),
                                        (64, torch.bfloat16, torch.bfloat16), (16, None, torch.float16), (32, torch.float16, None)])
def test_encoder_tail_vs_torch(CO, pdt, odt):
    """x + branch -> permute -> depth_to_space(4) (VMamba.py:1500-1505,1521-1523) -> clone -> ReflectionPad2d(1) ->
    compute dtype, channels-last: exact."""
    import xpoint_b200 as X
    import xpoint_b200.postprocess as P
    g = torch.Generator().manual_seed(CO)
    B, H, W = 3, 5, 7
    x = torch.randn(B, H, W, CO * 16, generator=g).to(DEV)
    pend = None if pdt is None else torch.randn(B, H, W, CO * 16, generator=g).to(pdt).to(DEV)
    assert P.encoder_tail_supported(x)
    enc, padded = P.encoder_tail(x, pend, 4, pad_dtype=odt)
    s = x if pend is None else x + pend
    ref = X.VSSM.depth_to_space(s.permute(0, 3, 1, 2), 4)
    assert torch.equal(enc, ref)
    if odt is None:
        assert padded is None
    else:
        refp = torch.nn.ReflectionPad2d(1)(ref).to(odt)
        assert padded.shape == refp.shape and padded.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(padded, refp)
    enc2, _ = P.encoder_tail(x, pend, 4, pad_dtype=odt, want_encoder_output=False) if odt is not None else (None, None)
    assert enc2 is None


def test_nms_golden_bit_exact():
    X, _ = _imports()
    g = golden("tail")
    p = torch.from_numpy(g["prob"]).to(DEV)
    pb = torch.from_numpy(g["prob_batched"]).to(DEV)
    assert np.array_equal(X.box_nms(p, 8, 0.015).cpu().numpy(), g["nms8"])
    assert np.array_equal(X.box_nms(p, 4, 0.015).cpu().numpy(), g["nms4"])
    assert np.array_equal(X.box_nms(p, 8, 0.015, keep_top_k=50).cpu().numpy(), g["nms8_top50"])
    assert np.array_equal(X.box_nms(p, 8, 0.015, iou=0.3).cpu().numpy(), g["nms8_iou03"])
    assert np.array_equal(X.box_nms(pb, 8, 0.015, keep_top_k=40, on_cpu=True).cpu().numpy(), g["nms8_batched_top40"])
    assert np.array_equal(X.box_nms(pb, 8, 0.015).cpu().numpy(), g["nms8_batched"])
    res = X.nms_keypoints(p[None], 8, 0.015)
    n = int(res.count[0])
    assert np.array_equal(res.keypoints[0, :n].cpu().numpy().astype(np.int64), g["kp8"])   # raster order


@pytest.mark.parametrize("H,W,B", [(256, 256, 2), (512, 640, 2), (1024, 1280, 1), (37, 53, 3)])
@pytest.mark.parametrize("size,topk", [(8, 0), (8, 4096), (4, 0), (4, 16384), (8, 7)])
def test_nms_vs_oracle_bit_exact(H, W, B, size, topk):
    X, O = _imports()
    g = torch.Generator().manual_seed(H + size + topk)
    prob = torch.rand(B, 1, H, W, generator=g) ** 6          # SURVEY C.8/C.9 score maps, no exact ties
    got = X.box_nms(prob.to(DEV), size, 0.015, keep_top_k=topk).cpu().numpy()
    ref = O.box_nms(prob.numpy(), size, 0.015, keep_top_k=topk)
    assert np.array_equal(got, ref)
    res = X.nms_keypoints(prob[:, 0].to(DEV), size, 0.015, keep_top_k=topk, capacity=max(topk, 1) if topk else H * W)
    for b in range(B):
        kp = O.extract_keypoints(ref[b, 0], 0.015)
        n = int(res.count[b])
        assert n == len(kp)
        assert np.array_equal(res.keypoints[b, :n].cpu().numpy().astype(np.int64), kp)


def test_nms_ties_and_degenerate():
    X, O = _imports()
    # exact ties: lower flat index wins, plateau of equal scores
    p = torch.zeros(1, 1, 24, 24)
    p[0, 0, 5, 5] = p[0, 0, 5, 8] = 0.5
    p[0, 0, 10:14, 10:14] = 0.25
    p[0, 0, 20, 3] = 0.9
    got = X.box_nms(p.to(DEV), 8, 0.015).cpu().numpy()
    assert np.array_equal(got, O.box_nms(p.numpy(), 8, 0.015))
    assert got[0, 0, 5, 5] == 0.5 and got[0, 0, 5, 8] == 0.0
    # nothing above threshold / everything above threshold / monotone ramp (worst case for the fixed point)
    z = torch.full((1, 1, 32, 40), 0.001)
    assert X.box_nms(z.to(DEV), 8, 0.015).abs().sum() == 0
    ramp = (torch.arange(48 * 64, dtype=torch.float32).reshape(1, 1, 48, 64) + 1) / (48 * 64)
    assert np.array_equal(X.box_nms(ramp.to(DEV), 8, 0.015).cpu().numpy(), O.box_nms(ramp.numpy(), 8, 0.015))
    assert np.array_equal(X.box_nms(ramp.to(DEV)[0, 0], 8, 0.015, keep_top_k=5).cpu().numpy(),
                          O.box_nms(ramp.numpy()[0, 0], 8, 0.015, keep_top_k=5))
    with pytest.raises(ValueError):
        X.box_nms(torch.zeros(3, 4, 5, device=DEV), 8, 0.015)


def test_interpolate_descriptors():
    X, O = _imports()
    g = golden("tail")
    kp = torch.from_numpy(g["kp8"]).to(DEV)
    d = X.interpolate_descriptors(kp, torch.from_numpy(g["desc_low"]).to(DEV), 96, 120)
    np.testing.assert_allclose(d.cpu().numpy(), g["desc_kp"], rtol=0, atol=2e-6)
    # batched, channel-last, ragged counts; zero rows past the count
    gen = torch.Generator().manual_seed(0)
    desc = torch.nn.functional.normalize(torch.randn(2, 256, 64, 80, generator=gen), dim=1)
    kps = torch.stack([torch.randint(0, 512, (2, 300), generator=gen), torch.randint(0, 640, (2, 300), generator=gen)], -1)
    cnt = torch.tensor([300, 123], dtype=torch.int32)
    out = X.sample_descriptors(kps.to(DEV).int(), cnt.to(DEV), desc.to(DEV), 512, 640)
    out_cl = X.sample_descriptors(kps.to(DEV).int(), cnt.to(DEV), desc.permute(0, 2, 3, 1).contiguous().to(DEV), 512, 640, True)
    for b in range(2):
        ref = O.interpolate_descriptors(kps[b, :cnt[b]].numpy(), desc[b].numpy(), 512, 640)
        np.testing.assert_allclose(out[b, :cnt[b]].cpu().numpy(), ref, rtol=0, atol=2e-6)
        np.testing.assert_allclose(out_cl[b, :cnt[b]].cpu().numpy(), ref, rtol=0, atol=2e-6)
        assert out[b, cnt[b]:].abs().sum() == 0


def _pairs(ms):
    return [(m.queryIdx, m.trainIdx) for m in ms]


@pytest.mark.parametrize("tc", [False, True])
def test_match_golden_bit_exact(tc):
    X, _ = _imports()
    g = golden("match")
    d1, d2 = torch.from_numpy(g["d1"]).to(DEV), torch.from_numpy(g["d2"]).to(DEV)
    m = X.get_matches(d1, d2, "bfmatcher", crossCheck=True, use_tensor_cores=tc)
    assert _pairs(m) == [tuple(p) for p in g["bf_pairs"].tolist()]
    np.testing.assert_allclose([x.distance for x in m], g["bf_dist"], rtol=1e-5, atol=1e-6)
    m2 = X.get_matches(g["d1"], g["d2"], "nnmatcher", threshold=10.0, use_tensor_cores=tc)   # numpy in, as the reference
    assert _pairs(m2) == [tuple(p) for p in g["nn_pairs"].tolist()]
    assert X.get_matches(d1[:0], d2, "bfmatcher", crossCheck=True) == []


def _descs(n1, n2, seed):
    g = torch.Generator().manual_seed(seed)
    d1 = torch.nn.functional.normalize(torch.randn(n1, 256, generator=g), dim=1)
    k = min(n1, n2) // 2
    d2 = torch.randn(n2, 256, generator=g)
    if k:
        d2[:k] = d1[torch.randperm(n1, generator=g)[:k]] + 0.05 * torch.randn(k, 256, generator=g)
    d2 = torch.nn.functional.normalize(d2, dim=1)[torch.randperm(n2, generator=g)]
    return d1, d2


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 50), (971, 850), (4096, 4096), (300, 4000), (130, 257)])
def test_match_vs_oracle_bit_exact(tc, n1, n2):
    """Identical fp32 unit descriptors on both sides; pairs must be bit-exact wherever the float64 oracle's
    best-vs-second gap is above fp32 noise (SURVEY C.13 min-gap guard, 1e-5 on squared distance)."""
    X, O = _imports()
    d1, d2 = _descs(n1, n2, n1 + n2)
    q, t, dist, gap_r = O.mnn_match(d1.numpy(), d2.numpy(), return_gap=True)
    _, _, _, gap_c = O.mnn_match(d2.numpy(), d1.numpy(), return_gap=True)
    res = X.mnn_match(d1[None].to(DEV), d2[None].to(DEV), use_tensor_cores=tc)
    idx = res.match_idx[0].cpu().numpy()
    got = [(i, int(j)) for i, j in enumerate(idx) if j >= 0]
    want = list(zip(q.tolist(), t.tolist()))
    # rows / columns whose best and second-best neighbours are closer than fp32 summation noise are ambiguous
    # for ANY fp32 implementation (OpenCV vs BLAS disagree there too); they are excluded, and must be rare
    amb_r = set(np.nonzero(gap_r <= 1e-5)[0].tolist()) if n2 > 1 else set()
    amb_c = set(np.nonzero(gap_c <= 1e-5)[0].tolist()) if n1 > 1 else set()
    assert len(amb_r) <= 0.01 * n1 + 1 and len(amb_c) <= 0.01 * n2 + 1
    keep = lambda pairs: [(i, j) for i, j in pairs if i not in amb_r and j not in amb_c]
    assert keep(got) == keep(want)
    if not amb_r and not amb_c:
        assert got == want and int(res.count[0]) == len(want)
    both = sorted(set(got) & set(want))
    qi = np.array([i for i, _ in both], dtype=np.int64)
    ref_d = {i: d for i, d in zip(q.tolist(), dist.tolist())}
    np.testing.assert_allclose(res.match_dist[0].cpu().numpy()[qi], [ref_d[i] for i in qi.tolist()], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("tc", [False, True])
def test_match_batched_ragged_and_ties(tc):
    X, O = _imports()
    P, s1, s2 = 3, 200, 260
    d1 = torch.zeros(P, s1, 256)
    d2 = torch.zeros(P, s2, 256)
    n1 = torch.tensor([200, 77, 0], dtype=torch.int32)
    n2 = torch.tensor([260, 131, 40], dtype=torch.int32)
    for p in range(P):
        a, b = _descs(max(int(n1[p]), 1), int(n2[p]), 100 + p)
        d1[p, :n1[p]] = a[:n1[p]]
        d2[p, :n2[p]] = b
    res = X.mnn_match(d1.to(DEV), d2.to(DEV), n1.to(DEV), n2.to(DEV), use_tensor_cores=tc)
    for p in range(P):
        q, t, _ = O.mnn_match(d1[p, :n1[p]].numpy(), d2[p, :n2[p]].numpy())
        idx = res.match_idx[p].cpu().numpy()
        assert [(i, int(j)) for i, j in enumerate(idx) if j >= 0] == list(zip(q.tolist(), t.tolist()))
        assert int(res.count[p]) == len(q)
    # duplicates resolve to the lowest index in both directions
    base = torch.nn.functional.normalize(torch.randn(4, 256, generator=torch.Generator().manual_seed(0)), dim=1)
    a = torch.stack([base[0], base[0], base[1]])
    b = torch.stack([base[0], base[1], base[1]])
    m = X.get_matches(a.to(DEV), b.to(DEV), "bfmatcher", crossCheck=True, use_tensor_cores=tc)
    assert _pairs(m) == [(0, 0), (2, 1)]


def test_match_config5_size_16384():
    """BASELINE configs[4]: 16 384 keypoints per image (1024x1280 pairs) -> a 16 384^2 x 256 similarity GEMM.  The float64
    oracle is too slow for the whole matrix, so: tcgen05 path == exact-fp32 CUDA-core path everywhere, and 512 sampled
    query rows == float64 numpy arg-min (size-independent property)."""
    X, _ = _imports()
    n = 16384
    d1, d2 = _descs(n, n, 5)
    a, b = d1[None].to(DEV), d2[None].to(DEV)
    r_tc = X.mnn_match(a, b, use_tensor_cores=True)
    r_fp = X.mnn_match(a, b, use_tensor_cores=False)
    nn_tc, nn_fp = r_tc.nn12[0].cpu().numpy(), r_fp.nn12[0].cpu().numpy()
    rows = np.random.default_rng(0).choice(n, 512, replace=False)
    dist = ((d1.numpy()[rows, None, :].astype(np.float64) - d2.numpy()[None, :, :].astype(np.float64)) ** 2).sum(-1)
    order = np.sort(dist, axis=1)
    clear = (order[:, 1] - order[:, 0]) > 1e-5                     # min-gap guard
    assert clear.mean() > 0.99
    assert np.array_equal(nn_tc[rows][clear], dist.argmin(1)[clear])
    assert np.array_equal(nn_fp[rows][clear], dist.argmin(1)[clear])
    assert (nn_tc != nn_fp).mean() < 1e-3                          # only fp32-noise ties may differ
    m_tc, m_fp = r_tc.match_idx[0].cpu().numpy(), r_fp.match_idx[0].cpu().numpy()
    assert (m_tc != m_fp).mean() < 2e-3
    assert abs(int(r_tc.count[0]) - int(r_fp.count[0])) <= 0.002 * n
    # mutual consistency of the returned pairs
    nn21 = r_tc.nn21[0].cpu().numpy()
    q = np.nonzero(m_tc >= 0)[0]
    assert np.array_equal(nn21[m_tc[q]], q)


def _hom_inputs(seed, B, k, Hh, Ww, outliers):
    rng = np.random.default_rng(seed)
    kp1 = np.stack([rng.integers(0, Hh, (B, k)), rng.integers(0, Ww, (B, k))], -1).astype(np.int32)
    kp2 = np.zeros_like(kp1)
    mi = np.full((B, k), -1, np.int32)
    n1 = np.zeros(B, np.int32)
    for b in range(B):
        ang, sc = rng.uniform(-0.3, 0.3), rng.uniform(0.8, 1.2)
        Hgt = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-40, 40)],
                        [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-40, 40)], [rng.uniform(-2e-4, 2e-4), 0.0, 1.0]])
        xy = np.concatenate([kp1[b][:, ::-1].astype(float), np.ones((k, 1))], 1) @ Hgt.T
        kp2b = np.round((xy[:, :2] / xy[:, 2:])[:, ::-1]).astype(np.int32)
        perm = rng.permutation(k)
        kp2[b] = kp2b[perm]
        inv = np.argsort(perm)
        n1[b] = rng.integers(k // 2, k + 1)
        rows = rng.choice(k, int(0.7 * k), replace=False)
        mi[b, rows] = inv[rows]
        wrong = rows[rng.random(rows.size) < outliers]
        mi[b, wrong] = rng.integers(0, k, wrong.size)
    return kp1, kp2, mi, n1


@pytest.mark.parametrize("B,k,outliers", [(3, 1024, 0.3), (2, 4096, 0.6), (4, 200, 0.1), (1, 16384, 0.4), (1, 20000, 0.2)])
def test_homography_vs_oracle(B, k, outliers):
    """xp_estimate_homography against its oracle restatement: same winning inlier set (the fp64 arithmetic of hypothesis
    scoring is identical on both sides), H to 1e-8 (the least-squares sums are added in a different order)."""
    X, O = _imports()
    Hh, Ww = 512, 640
    kp1, kp2, mi, n1 = _hom_inputs(k + B, B, k, Hh, Ww, outliers)
    r = X.estimate_homography(torch.from_numpy(kp1).to(DEV), torch.from_numpy(kp2).to(DEV), torch.from_numpy(mi).to(DEV), Hh, Ww,
                              n1=torch.from_numpy(n1).to(DEV))
    assert r.H.dtype == torch.float64 and r.H.shape == (B, 3, 3) and r.inliers.shape == (B, k)
    for b in range(B):
        H, inl, cnt, _ = O.estimate_homography(kp1[b], kp2[b], mi[b], Hh, Ww, n1=int(n1[b]), pair_index=b)
        assert int(r.n_inliers[b]) == cnt
        assert np.array_equal(r.inliers[b].cpu().numpy(), inl)
        np.testing.assert_allclose(r.H[b].cpu().numpy(), H, rtol=1e-8, atol=1e-8)
        assert cnt >= 0.5 * (1 - outliers) * min((mi[b, : n1[b]] >= 0).sum(), 11264)


def test_homography_golden_and_degenerate():
    """The OpenCV fixtures (tests/golden/homography.npz) through the GPU path, the cv2-shaped wrapper, and pairs without
    an estimate."""
    X, O = _imports()
    g = golden("homography")
    Hh, Ww = int(g["height"]), int(g["width"])
    for name in ("easy", "hard", "few"):
        kp1, kp2, mi, q = (torch.from_numpy(g[name + k]).to(DEV) for k in ("_kp1", "_kp2", "_match_idx", "_query"))
        r = X.estimate_homography(kp1[None], kp2[None], mi[None], Hh, Ww)
        mask_cv = g[name + "_mask_cv"].astype(bool)
        assert (r.inliers[0][q.long()].cpu().numpy() == mask_cv).mean() >= 0.98
        c = np.array([[0, 0, 1], [Ww, 0, 1], [0, Hh, 1], [Ww, Hh, 1]], float)
        a, bb = c @ r.H[0].cpu().numpy().T, c @ g[name + "_H_gt"].T
        assert np.linalg.norm(a[:, :2] / a[:, 2:] - bb[:, :2] / bb[:, 2:], axis=1).mean() < 2.0
        # the call shape of evaluation.py:368-375
        src = kp1[q.long()].flip(-1).float().reshape(-1, 1, 2)
        dst = kp2[mi[q.long()].long()].flip(-1).float().reshape(-1, 1, 2)
        H2, m2 = X.find_homography(src, dst, 3.0, Hh, Ww)
        assert H2 is not None and m2.shape == (len(q), 1) and (m2[:, 0].bool().cpu().numpy() == mask_cv).mean() >= 0.98
    kp1 = torch.from_numpy(g["easy_kp1"]).to(DEV)[None]
    kp2 = torch.from_numpy(g["easy_kp2"]).to(DEV)[None]
    mi = torch.full((1, kp1.shape[1]), -1, dtype=torch.int32, device=DEV)
    mi[0, :3] = torch.tensor([5, 6, 7], dtype=torch.int32)
    r = X.estimate_homography(kp1, kp2, mi, Hh, Ww)
    assert int(r.n_inliers[0]) == -1 and float(r.H.abs().sum()) == 0.0 and not bool(r.inliers.any())
    assert X.find_homography(torch.zeros(3, 2, device=DEV), torch.zeros(3, 2, device=DEV))[0] is None
    with pytest.raises(RuntimeError):
        X.estimate_homography(kp1, kp2, mi[:, :5], Hh, Ww)


def test_pipeline_with_homography():
    """PairPipeline(estimate_homography=True): thermal map = optical map shifted by (7, -5) pixels -> H is that translation."""
    X, O = _imports()
    g = torch.Generator().manual_seed(3)
    B, H, W, k = 2, 256, 320, 512
    prob_o = torch.rand(B, 1, H, W, generator=g) ** 6
    prob_t = torch.roll(prob_o, shifts=(7, -5), dims=(2, 3))
    desc_o = torch.nn.functional.normalize(torch.randn(B, 256, H // 8, W // 8, generator=g), dim=1)
    pipe = X.PairPipeline(None, nms=8, detection_threshold=0.015, keep_top_k=k, estimate_homography=True)
    # descriptors: sample the same field for both (a match is then decided by position only for identical maps)
    r = pipe.tail(prob_o.to(DEV), prob_o.to(DEV), desc_o.to(DEV), desc_o.to(DEV))
    assert r.H.shape == (B, 3, 3) and r.inliers.shape == (B, k)
    eye = torch.eye(3, dtype=torch.float64, device=DEV)
    for b in range(B):
        assert int(r.n_inliers[b]) == int(r.n_matches[b]) == k          # identical images: every keypoint matches itself
        assert torch.allclose(r.H[b], eye, atol=1e-9)
    del prob_t


@pytest.mark.parametrize("C", [48, 100, 2048])
def test_get_matches_any_descriptor_size(C):
    """descriptor sizes outside the tensor-core kernel's range (C % 32, C <= 1024) take the fp32 kernels instead
    of raising; the reference's get_matches accepts any descriptor_size."""
    import xpoint_b200 as X
    from oracle import oracle as O
    g = torch.Generator().manual_seed(C)
    d1 = torch.nn.functional.normalize(torch.randn(150, C, generator=g), dim=1)
    d2 = torch.nn.functional.normalize(torch.randn(170, C, generator=g), dim=1)
    q, t, _ = O.mnn_match(d1.numpy(), d2.numpy())
    m = X.get_matches(d1.to("cuda"), d2.to("cuda"), "bfmatcher", crossCheck=True)      # use_tensor_cores defaults to True
    assert [(a.queryIdx, a.trainIdx) for a in m] == list(zip(q.tolist(), t.tolist()))
