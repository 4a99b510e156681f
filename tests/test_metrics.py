"""Evaluation-driver metrics: oracle restatement vs the reference's own functions (golden, CPU) and
the device kernels vs both (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import golden

THR = 0.015


def _kps(prob):
    return np.argwhere(prob > THR)


def test_oracle_warp_and_metrics_vs_reference_golden():
    from oracle import oracle as O
    g = golden("metrics")
    assert np.array_equal(O.warp_keypoints(g["wk_in"], g["H_t"][2]), g["wk_int"])
    np.testing.assert_allclose(O.warp_keypoints(g["wk_in"].astype(np.float32), g["H_t"][2], as_int=False), g["wk_float"], rtol=0, atol=1e-9)
    Hh, Ww = g["prob_o"].shape[-2:]
    for b in range(g["prob_o"].shape[0]):
        kp_o, kp_t = _kps(g["prob_o"][b, 0]), _kps(g["prob_t"][b, 0])
        assert len(kp_o) == g["n_kp_o"][b] and len(kp_t) == g["n_kp_t"][b]
        rep, _, _ = O.repeatability_sample(kp_o, kp_t, g["H_o"][b], g["H_t"][b], (Hh, Ww), g["thr_rep"].tolist())
        np.testing.assert_allclose([rep[t] for t in g["thr_rep"].tolist()], g["repeatability"][b], rtol=0, atol=1e-12)
        d_o = O.interpolate_descriptors(kp_o, g["desc_o"][b], Hh, Ww)
        d_t = O.interpolate_descriptors(kp_t, g["desc_t"][b], Hh, Ww)
        q, t, _ = O.mnn_match(d_o, d_t)
        ms, _ = O.match_scores_sample(kp_o, kp_t, list(zip(q.tolist(), t.tolist())), g["H_o"][b], g["H_t"][b], (Hh, Ww), g["thr_kp"].tolist())
        for i, th in enumerate(g["thr_kp"].tolist()):
            assert ms[th]["n_correct"] == g["n_correct_optical"][b, i]
            assert abs(ms[th]["m_score"] - g["m_score_optical"][b, i]) < 1e-12


@pytest.mark.gpu
def test_device_metrics_vs_reference_golden_and_oracle():
    import xpoint_b200 as X
    from oracle import oracle as O
    from xpoint_b200 import metrics as M
    g = golden("metrics")
    dev = "cuda"
    Bn, _, Hh, Ww = g["prob_o"].shape
    w = M.warp_keypoints(torch.from_numpy(g["wk_in"]).to(dev), torch.from_numpy(g["H_t"][2]).to(dev), height=Hh, width=Ww)
    assert np.array_equal(w.points_int.cpu().numpy(), g["wk_int"])
    np.testing.assert_allclose(w.points_float.cpu().numpy(), g["wk_float"], rtol=0, atol=1e-9)
    ins = (g["wk_int"][:, 0] >= 0) & (g["wk_int"][:, 1] >= 0) & (g["wk_int"][:, 0] < Hh) & (g["wk_int"][:, 1] < Ww)
    assert np.array_equal(w.inside.cpu().numpy(), ins)
    # the golden score maps are sparse random points (no NMS structure): extract the keypoints directly, in raster order
    kp_list = [_kps(p[0]) for p in np.concatenate([g["prob_o"], g["prob_t"]])]
    k = max(len(a) for a in kp_list)
    kp = torch.zeros(2 * Bn, k, 2, dtype=torch.int32)
    cnt = torch.tensor([len(a) for a in kp_list], dtype=torch.int32)
    for i, a in enumerate(kp_list):
        kp[i, :len(a)] = torch.from_numpy(a).int()
    kp, cnt = kp.to(dev), cnt.to(dev)
    desc = torch.from_numpy(np.concatenate([g["desc_o"], g["desc_t"]])).to(dev)
    d = X.sample_descriptors(kp, cnt, desc, Hh, Ww)
    m = X.mnn_match(d[:Bn], d[Bn:], cnt[:Bn], cnt[Bn:])
    Ho, Ht = torch.from_numpy(g["H_o"]).to(dev), torch.from_numpy(g["H_t"]).to(dev)
    rep = M.repeatability(kp[:Bn], cnt[:Bn], kp[Bn:], cnt[Bn:], Ho, Ht, Hh, Ww, g["thr_rep"].tolist())
    np.testing.assert_allclose(rep.repeatability.cpu().numpy(), g["repeatability"], rtol=0, atol=1e-12)
    ms = M.matching_scores(kp[:Bn], cnt[:Bn], kp[Bn:], cnt[Bn:], m.match_idx, Ho, Ht, Hh, Ww, g["thr_kp"].tolist())
    assert np.array_equal(ms.n_correct_optical.cpu().numpy(), g["n_correct_optical"])
    np.testing.assert_allclose(ms.m_score_optical.cpu().numpy(), g["m_score_optical"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ms.m_score_thermal.cpu().numpy(), g["m_score_thermal"], rtol=0, atol=1e-12)
    # and against the oracle on a bigger random case (4096 keypoints, projective ground truth)
    rng = np.random.default_rng(0)
    n = 3000
    kpa = np.stack([rng.integers(0, 512, n), rng.integers(0, 640, n)], 1)
    Hgt = np.array([[1.01, 0.02, -6.0], [-0.015, 0.99, 4.0], [2e-5, -1e-5, 1.0]], np.float32)
    kpb = O.warp_keypoints(kpa, Hgt) + rng.integers(-2, 3, (n, 2))
    kpb = kpb[(kpb[:, 0] >= 0) & (kpb[:, 0] < 512) & (kpb[:, 1] >= 0) & (kpb[:, 1] < 640)]
    ka = torch.zeros(1, 4096, 2, dtype=torch.int32); ka[0, :n] = torch.from_numpy(kpa).int()
    kb = torch.zeros(1, 4096, 2, dtype=torch.int32); kb[0, :len(kpb)] = torch.from_numpy(kpb).int()
    na, nb = torch.tensor([n], dtype=torch.int32), torch.tensor([len(kpb)], dtype=torch.int32)
    eye = torch.eye(3)[None]
    r = M.repeatability(ka.to(dev), na.to(dev), kb.to(dev), nb.to(dev), eye.to(dev), torch.from_numpy(Hgt)[None].to(dev), 512, 640, [1, 2.5, 3])
    want, Nt, No = O.repeatability_sample(kpa, kpb, np.eye(3), Hgt, (512, 640), [1, 2.5, 3])
    assert int(r.n_warped_thermal[0]) == Nt and int(r.n_warped_optical[0]) == No
    np.testing.assert_allclose(r.repeatability[0].cpu().numpy(), [want[t] for t in (1, 2.5, 3)], rtol=0, atol=1e-12)
