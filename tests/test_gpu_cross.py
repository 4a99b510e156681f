"""CrossScan / CrossMerge / merge+norm+gate parity (C ABI vs oracle and reference goldens).  -m gpu."""
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


def test_golden_cross():
    X, _ = _imports()
    g = golden("cross")
    x = torch.from_numpy(g["x"]).to(DEV)
    x4 = torch.from_numpy(g["x4"]).to(DEV)
    ys = torch.from_numpy(g["ys"]).to(DEV)
    for s in (0, 1, 2):
        assert np.array_equal(X.cross_scan_fn(x, True, True, False, s).cpu().numpy(), g[f"scan_s{s}"])
        assert np.array_equal(X.cross_scan_fn(x.permute(0, 2, 3, 1).contiguous(), False, False, False, s).cpu().numpy(),
                              g[f"scan_s{s}_cl"])
        assert np.array_equal(X.cross_scan_fn(x, True, False, False, s).cpu().numpy(), g[f"scan_s{s}_cf2cl"])
        assert np.array_equal(X.cross_scan_fn(x4, True, True, True, s).cpu().numpy(), g[f"scan1b1_s{s}"].reshape(2, 4, 3, 35))
        assert np.array_equal(X.cross_merge_fn(ys, True, True, True, s).cpu().numpy(), g[f"merge1b1_s{s}"].reshape(2, 4, 3, 35))
        tol = 0 if s == 0 else 1e-6
        np.testing.assert_allclose(X.cross_merge_fn(ys, True, True, False, s).cpu().numpy(), g[f"merge_s{s}"], rtol=0, atol=tol)
        np.testing.assert_allclose(
            X.cross_merge_fn(ys.permute(0, 3, 4, 1, 2).contiguous(), False, False, False, s).cpu().numpy(),
            g[f"merge_s{s}_cl"], rtol=0, atol=tol)
    assert np.array_equal(X.CrossScanF.apply(x, True, True, False, 0).cpu().numpy(), g["F_scan"])
    assert np.array_equal(X.CrossMergeF.apply(ys, True, True, False, 0).cpu().numpy(), g["F_merge"])
    assert np.array_equal(X.CrossScanTritonF.apply(x, True, True, False, 0).cpu().numpy(), g["F_scan"])


@pytest.mark.parametrize("shape", [(27, 253, 57, 58), (3, 5, 7, 9), (2, 96, 128, 160), (1, 8, 33, 31)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_cross_scan_merge_vs_oracle(shape, dtype):
    """Shapes of the reference's own self-check (csm_triton.py:604) plus the stage-0 token grid at 512x640."""
    X, O = _imports()
    B, C, H, W = shape
    if B * C * H * W > 3e7 and dtype != torch.float32:
        pytest.skip("large case covered in fp32")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    xs = X.cross_scan_fn(x.to(DEV))
    assert xs.dtype == dtype
    assert np.array_equal(xs.float().cpu().numpy(), O.cross_scan(x.float().numpy()))   # exact: pure data movement
    ys = torch.randn(B, 4, C, H, W, generator=g).to(dtype)
    y = X.cross_merge_fn(ys.to(DEV))
    ref = O.cross_merge(ys.float().numpy())
    if dtype == torch.float32:
        assert np.array_equal(y.cpu().numpy(), ref)    # same association as the torch path
    else:
        assert_close(y.float().cpu().numpy(), ref, 1e-2, "merge 16-bit")


@pytest.mark.parametrize("in_cf,out_cf", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("scans", [0, 1, 2])
def test_layout_combinations(in_cf, out_cf, scans):
    X, O = _imports()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 9, 11, generator=g)
    ref = torch.from_numpy(O.cross_scan(x.numpy(), scans))                 # (B,4,C,L)
    xin = x if in_cf else x.permute(0, 2, 3, 1).contiguous()
    got = X.cross_scan_fn(xin.to(DEV), in_cf, out_cf, False, scans).cpu()
    want = ref if out_cf else ref.permute(0, 3, 1, 2).contiguous()
    assert torch.equal(got, want)
    ys = torch.randn(2, 4, 6, 9, 11, generator=g)
    refm = torch.from_numpy(O.cross_merge(ys.numpy(), scans))              # (B,C,L)
    yin = ys if out_cf else ys.permute(0, 3, 4, 1, 2).contiguous()
    gotm = X.cross_merge_fn(yin.to(DEV), in_cf, out_cf, False, scans).cpu()
    wantm = refm if in_cf else refm.permute(0, 2, 1).contiguous()
    torch.testing.assert_close(gotm, wantm, rtol=0, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 96, 16, 20), (3, 40, 13, 7), (1, 192, 64, 80)])
@pytest.mark.parametrize("gate", [False, True])
def test_merge_norm_gate(shape, gate):
    X, O = _imports()
    B, C, H, W = shape
    g = torch.Generator().manual_seed(2)
    ys = torch.randn(B, 4, C, H * W, generator=g)
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    z = torch.randn(B, H, W, C, generator=g) if gate else None
    out = X.merge_norm_gate(ys.to(DEV), H, W, w.to(DEV), b.to(DEV), None if z is None else z.to(DEV))
    ref = O.merge_norm_gate(O.cross_merge(ys.view(B, 4, C, H, W).numpy()), w, b, None if z is None else z.reshape(B, H * W, C))
    assert_close(out.cpu().numpy().reshape(B, H * W, C), ref, 1e-5, "merge_norm_gate")


def test_errors():
    X, _ = _imports()
    with pytest.raises(RuntimeError):
        X.cross_scan_fn(torch.zeros(1, 2, 3, 4))   # CPU tensor
    with pytest.raises(RuntimeError):
        X.cross_scan_fn(torch.zeros(1, 2, 3, 4, device=DEV), scans=5)
    assert X.cross_scan_fn(torch.zeros(0, 2, 3, 4, device=DEV)).shape == (0, 4, 2, 12)


@pytest.mark.parametrize("C", [48, 96, 192, 384, 768, 1536, 37])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.float32, torch.float32), (torch.float32, torch.float16),
                                          (torch.float16, torch.float16), (torch.bfloat16, torch.float32)])
def test_layer_norm(C, dt_in, dt_out):
    """xp_layer_norm vs torch's fp32 LayerNorm on the same (pre-rounded) input."""
    from xpoint_b200.cross_scan import layer_norm
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(3, 17, 5, C, generator=g) * 2 + 0.5).to(dt_in)
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-5)
    out = layer_norm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5, dt_out)
    assert out.dtype == dt_out and out.shape == x.shape
    assert_close(out.float().cpu().numpy(), ref.numpy(), 1e-5 if dt_out == torch.float32 else 2e-3, f"layer_norm C={C}")
