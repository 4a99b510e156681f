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
