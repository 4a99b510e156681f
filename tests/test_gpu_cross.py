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


# ------------------------------------------------------------------------------------------------------------------
# copy-free SS2D core (xp_ss2d_pack / xp_ss2d_dwconv_pack / xp_ss2d_merge_norm + the scan's fused addressing)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 5, 12, 20), (1, 3, 33, 70), (3, 16, 64, 80), (1, 2, 1, 7)])
def test_ss2d_pack_exact(dtype, shape):
    from xpoint_b200 import ss2d
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g).to(dtype).to(DEV)
    xx = ss2d.ss2d_pack(x)
    B, D, H, W = shape
    assert torch.equal(xx[:, 0], x.reshape(B, D, H * W))
    assert torch.equal(xx[:, 1], x.transpose(2, 3).reshape(B, D, H * W))
    # the two layouts are exactly directions 0 and 1 of the reference CrossScan (csm_triton.py:22-29)
    import xpoint_b200 as X
    xs = X.cross_scan_fn(x, True, True, False, 0)
    assert torch.equal(xx[:, 0], xs[:, 0]) and torch.equal(xx[:, 1], xs[:, 1])
    assert torch.equal(xx[:, 0].flip(-1), xs[:, 2]) and torch.equal(xx[:, 1].flip(-1), xs[:, 3])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,H,W,D,Cin,bias", [(2, 12, 20, 16, 16, False), (1, 33, 70, 24, 48, True), (2, 64, 80, 96, 96, False),
                                              (1, 16, 32, 40, 80, True), (1, 5, 3, 7, 7, True)])
def test_ss2d_dwconv_pack_vs_torch(dtype, tol, B, H, W, D, Cin, bias):
    """SS2D's depth-wise 3x3 conv + SiLU (VMamba.py:651-655) on the channel-last in_proj output, both layouts out."""
    from xpoint_b200 import ss2d
    g = torch.Generator().manual_seed(2)
    xcl = torch.randn(B, H, W, Cin, generator=g).to(dtype).to(DEV)
    w = (0.3 * torch.randn(D, 1, 3, 3, generator=g)).to(DEV)
    b = (0.1 * torch.randn(D, generator=g)).to(DEV) if bias else None
    xx = ss2d.ss2d_dwconv_pack(xcl[..., :D] if Cin != D else xcl, D, w, b, silu=True)
    ref = torch.nn.functional.silu(torch.nn.functional.conv2d(xcl[..., :D].float().permute(0, 3, 1, 2), w, b, padding=1, groups=D))
    np.testing.assert_allclose(xx[:, 0].float().cpu().numpy(), ref.reshape(B, D, H * W).cpu().numpy(), rtol=tol, atol=tol)
    assert torch.equal(xx[:, 1].view(B, D, W, H), xx[:, 0].view(B, D, H, W).transpose(2, 3))


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("B,D,H,W,gate", [(2, 16, 12, 20, False), (1, 96, 16, 24, True), (2, 192, 8, 12, False), (1, 770, 8, 8, True),
                                          (1, 1536, 4, 8, False), (1, 33, 20, 36, True)])
def test_ss2d_merge_norm_vs_torch(out_dtype, B, D, H, W, gate):
    """LayerNorm_D(ys0 + ys1 + (ys2 + ys3)^T) [* z]: CrossMerge (csm_triton.py:56-62) + out_norm (VMamba.py:644) with the
    planes already in natural memory order."""
    from xpoint_b200 import ss2d
    g = torch.Generator().manual_seed(3)
    L = H * W
    ys = torch.randn(B, 4, D, L, generator=g).to(DEV)
    gam = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
    bet = (0.1 * torch.randn(D, generator=g)).to(DEV)
    z = torch.randn(B, H, W, D, generator=g).to(out_dtype).to(DEV) if gate else None
    out = ss2d.ss2d_merge_norm(ys, H, W, gam, bet, z, 1e-5, out_dtype=out_dtype)
    merged = (ys[:, 0] + ys[:, 1]).view(B, D, H, W) + (ys[:, 2] + ys[:, 3]).view(B, D, W, H).transpose(2, 3)
    ref = torch.nn.functional.layer_norm(merged.permute(0, 2, 3, 1), (D,), gam, bet, 1e-5)
    if gate:
        ref = ref * z.float()
    tol = 1e-5 if out_dtype == torch.float32 else 2e-3
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), rtol=tol, atol=tol)
    # and against the scan-order kernel the unfused path uses (flipped planes, reference direction order)
    import xpoint_b200 as X
    ys_ref_order = torch.stack([ys[:, 0], ys[:, 2], ys[:, 1].flip(-1), ys[:, 3].flip(-1)], dim=1).contiguous()
    old = X.merge_norm_gate(ys_ref_order, H, W, gam, bet, z, 1e-5, out_dtype=out_dtype)
    np.testing.assert_allclose(out.float().cpu().numpy(), old.float().cpu().numpy(), rtol=tol, atol=tol)


def test_ss2d_fused_errors():
    from xpoint_b200 import ss2d
    ys = torch.randn(1, 4, 8, 6 * 9, device=DEV)
    w = torch.ones(8, device=DEV)
    with pytest.raises(RuntimeError):
        ss2d.ss2d_merge_norm(ys, 6, 9, w, w)                     # H, W must be multiples of 4
    with pytest.raises(RuntimeError):
        ss2d.ss2d_merge_norm(ys.half(), 6, 9, w, w)              # fp32 planes only
    with pytest.raises(RuntimeError):
        ss2d.ss2d_pack(torch.randn(1, 2, 4, 4))                  # no CPU path
    assert ss2d.ss2d_pack(torch.randn(0, 2, 4, 4, device=DEV)).shape == (0, 2, 2, 16)


# ------------------------------------------------------------------------------------------------------------------
# f2 glue kernels: add + bias + LayerNorm, patch-embed stem
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [8, 48, 96, 100, 192, 384, 768, 1536])
@pytest.mark.parametrize("mode", ["ln", "add_ln", "bias_ln", "add_only"])
@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.float16, torch.float16), (torch.float16, torch.float32)])
def test_add_layer_norm_vs_torch(C, mode, xdt, ydt):
    """x = x + branch; n = LayerNorm(x) of VSSBlock (VMamba.py:1222-1234) / conv bias + LayerNorm of patch-embed and
    downsample (VMamba.py:1405-1440) in one pass."""
    from xpoint_b200.cross_scan import add_layer_norm
    g = torch.Generator().manual_seed(C)
    x = torch.randn(3, 7, 5, C, generator=g).to(xdt).to(DEV)
    res = torch.randn(3, 7, 5, C, generator=g).to(DEV) if mode in ("add_ln", "add_only") else None
    pb = torch.randn(C, generator=g).to(DEV) if mode == "bias_ln" else None
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C, generator=g)).to(DEV)
    want_y = mode != "add_only"
    s, y = add_layer_norm(x, res, w if want_y else None, b if want_y else None, 1e-5, y_dtype=ydt, pre_bias=pb,
                          sum_dtype=torch.float32 if mode == "add_ln" else ydt, want_sum=mode in ("add_ln", "add_only"),
                          want_y=want_y)
    ref_s = x.float() + (res if res is not None else 0) + (pb if pb is not None else 0)
    tol = 2e-6 if ydt == torch.float32 else 2e-3
    if s is not None:
        np.testing.assert_allclose(s.float().cpu().numpy(), ref_s.cpu().numpy(), rtol=tol, atol=tol)
    if want_y:
        ref = torch.nn.functional.layer_norm(ref_s, (C,), w, b, 1e-5)
        np.testing.assert_allclose(y.float().cpu().numpy(), ref.cpu().numpy(), rtol=tol * 5, atol=tol * 5)
        assert y.dtype == ydt


@pytest.mark.parametrize("cin", [1, 3])
@pytest.mark.parametrize("C1,H,W,odt", [(48, 64, 96, torch.float32), (48, 64, 96, torch.float16), (8, 33, 47, torch.float32),
                                        (64, 16, 16, torch.bfloat16), (12, 20, 36, torch.float16)])
def test_patch_embed_stem_vs_torch(cin, C1, H, W, odt):
    """Conv2d(k3, s2, p1) + bias -> LayerNorm(C1) -> GELU, channel-last out (first half of patch-embed v2, VMamba.py:1405-1413)."""
    from xpoint_b200.cross_scan import patch_embed_stem
    g = torch.Generator().manual_seed(C1 + H)
    img = torch.rand(2, cin, H, W, generator=g).to(DEV)
    w = (0.3 * torch.randn(C1, cin, 3, 3, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C1, generator=g)).to(DEV)
    lw = (1 + 0.1 * torch.randn(C1, generator=g)).to(DEV)
    lb = (0.1 * torch.randn(C1, generator=g)).to(DEV)
    out = patch_embed_stem(img, w, b, lw, lb, 1e-5, out_dtype=odt, gelu=True)
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(img, w, b, stride=2, padding=1).permute(0, 2, 3, 1)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(ref, (C1,), lw, lb, 1e-5))
    assert out.shape == ref.shape and out.dtype == odt
    tol = {torch.float32: 2e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[odt]
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), rtol=tol, atol=tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,G,D,R,N,L", [(2, 4, 96, 6, 1, 1280), (1, 4, 24, 3, 2, 328), (3, 2, 40, 8, 4, 64), (1, 4, 16, 1, 1, 24)])
def test_ss2d_dt_proj_vs_torch(dtype, tol, B, G, D, R, N, L):
    """dt_proj of SS2D (VMamba.py:607-608) on a strided view of the x_proj output."""
    _dt_proj_case(dtype, tol, B, G, D, R, N, L)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,G,D,R,N,L", [(2, 4, 192, 12, 1, 1280), (1, 4, 40, 9, 2, 328), (2, 2, 24, 16, 1, 64), (1, 4, 136, 13, 1, 520)])
def test_ss2d_dt_proj_rank_9_to_16(dtype, tol, B, G, D, R, N, L):
    """Ranks 9..16 (XPoint stage 1: dt_rank 12) take the mixed-precision-FMA kernel; weights are rounded to the input dtype
    as the reference's autocast GEMM does."""
    _dt_proj_case(dtype, tol, B, G, D, R, N, L)


def _dt_proj_case(dtype, tol, B, G, D, R, N, L):
    from xpoint_b200 import ss2d
    g = torch.Generator().manual_seed(R + L)
    x_dbl = torch.randn(B, G, R + 2 * N, L, generator=g).to(dtype).to(DEV)
    w = (0.4 * torch.randn(G, D, R, generator=g)).to(DEV)
    out = ss2d.ss2d_dt_proj(x_dbl[:, :, :R], w)
    ref = torch.einsum("gdr,bgrl->bgdl", w, x_dbl[:, :, :R].float())
    assert out.dtype == dtype and out.shape == (B, G, D, L)
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), rtol=tol, atol=tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("M,K,N", [(128, 96, 384), (1000, 96, 384), (20480 * 2 + 37, 96, 384), (5120, 192, 768), (640, 384, 1536),
                                   (321, 768, 3072), (77, 16, 64), (300, 64, 320), (129, 8, 32)])
@pytest.mark.parametrize("gelu", [True, False])
def test_linear_act_tc_vs_torch(dtype, tol, M, K, N, gelu):
    """fc1 (+ exact GELU) of the VSSBlock MLP (VMamba.py:110-128) on tcgen05: fp32 accumulate, 16-bit out."""
    from xpoint_b200.cross_scan import linear_act
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g).to(dtype).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).to(DEV)
    b = (0.5 * torch.randn(N, generator=g)).to(DEV)
    out = linear_act(x, w, b, gelu=gelu)
    ref = x.float() @ w.float().t() + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert out.shape == (M, N) and out.dtype == dtype
    np.testing.assert_allclose(out.float().cpu().numpy(), ref.cpu().numpy(), rtol=tol, atol=tol)


def test_gelu_erf_epilogue_accuracy():
    """The epilogue's erf (Abramowitz-Stegun 7.1.26) against torch's exact GELU on a dense sweep: identity GEMM, K = 64."""
    from xpoint_b200.cross_scan import linear_act
    K = 64
    vals = torch.linspace(-8, 8, 64 * 4096)
    # one-hot rows select fp16-representable values: out[m, n] = gelu(v[m] * [n == m % 64] ...) -> use a diagonal weight
    x = torch.zeros(4096, K)
    vgrid = vals.view(4096, 64)
    w = torch.eye(64, K)
    x[:, :64] = vgrid
    out = linear_act(x.half().to(DEV), w.half().to(DEV), None, gelu=True).float().cpu()
    ref = torch.nn.functional.gelu(x.half().float()[:, :64])
    assert (out - ref).abs().max() <= 1e-3 * 8 / 8 + 2 ** -11 * ref.abs().max()   # fp16 output rounding only
    np.testing.assert_allclose(out.numpy(), ref.half().float().numpy(), rtol=2e-3, atol=1e-4)


# ------------------------------------------------------------------------------------------ fused SS2D core (xp_ss2d_core)
CORE_SHAPES = [(128, 160, 8), (64, 80, 12), (32, 40, 24), (16, 20, 36), (12, 20, 8), (8, 8, 4), (36, 44, 6), (4, 4, 2),
               (20, 16, 5), (128, 4, 3)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N", [1, 2])
@pytest.mark.parametrize("H,W,D", CORE_SHAPES)
def test_ss2d_core_matches_scan_plus_merge(H, W, D, N, dtype):
    """xp_ss2d_core (four directions merged in shared memory, one fp32 plane out) against the op-level path it replaces:
    xp_selective_scan_fwd with the fused addressing (four fp32 planes in natural order) summed like CrossMerge
    (csm_triton.py:56-85).  Covers whole and partial 256/512-token blocks, planes of 1..12 channels per CTA, multi-warp
    step pipelining (128x160) and both column pitches."""
    from xpoint_b200 import ss2d as S
    from xpoint_b200.selective_scan import scan_forward
    if S.core_channels(D, N, H, W, dtype) == 0:
        pytest.skip("plane does not fit in shared memory for this dtype")
    g = torch.Generator().manual_seed(H * 1000 + W * 10 + N)
    B, K, L, R = 3, 4, H * W, 3
    x = torch.randn(B, D, H, W, generator=g)
    xx = torch.stack([x.reshape(B, D, L), x.transpose(2, 3).reshape(B, D, L)], 1).to(dtype).to(DEV)      # [x ; x^T]
    delta = (0.5 * torch.rand(B, K, D, L, generator=g) - 0.3).to(dtype).to(DEV)
    x_dbl = torch.randn(B, K, R + 2 * N, L, generator=g).to(dtype).to(DEV)
    Bs, Cs = x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:]
    A = (-0.5 * torch.rand(K * D, N, generator=g) - 0.01).to(DEV)
    Ds = torch.randn(K * D, generator=g).to(DEV)
    bias = (0.5 * torch.rand(K * D, generator=g)).to(DEV)
    ys, _ = scan_forward(xx.view(B, 2 * D, L), delta.view(B, K * D, L), A, Bs, Cs, Ds, None, bias, True, True,
                         u_group_div=2, reverse_group_mask=S.REVERSE_MASK)
    ys = ys.view(B, K, D, L)
    want = (ys[:, 0] + ys[:, 1]).view(B, D, H, W) + (ys[:, 2] + ys[:, 3]).view(B, D, W, H).transpose(2, 3)
    got = S.ss2d_core(xx, delta, A, Bs, Cs, Ds, bias, H, W, True)
    torch.cuda.synchronize()
    assert got.shape == (B, D, H, W) and got.dtype == torch.float32
    assert_close(got.cpu().numpy(), want.cpu().numpy(), 2e-5, f"ss2d_core {H}x{W} D={D} N={N} {dtype}")
    # bit-reproducible: the merge order inside the kernel does not depend on which direction reaches a block first
    again = S.ss2d_core(xx, delta, A, Bs, Cs, Ds, bias, H, W, True)
    assert torch.equal(got, again)
    # out_norm on the merged plane == merge_norm on the four planes
    gam, bet = torch.randn(D, generator=g).to(DEV), torch.randn(D, generator=g).to(DEV)
    n1 = S.ss2d_plane_norm(got, gam, bet, None, 1e-5, out_dtype=dtype)
    n4 = S.ss2d_merge_norm(ys, H, W, gam, bet, None, 1e-5, out_dtype=dtype)
    assert_close(n1.float().cpu().numpy(), n4.float().cpu().numpy(), 2e-5 if dtype == torch.float32 else 1e-2, "plane_norm")


def test_ss2d_core_no_softplus_and_optional_operands():
    from xpoint_b200 import ss2d as S
    from xpoint_b200.selective_scan import scan_forward
    g = torch.Generator().manual_seed(5)
    B, D, H, W, N, K = 2, 6, 24, 28, 1, 4
    L = H * W
    x = torch.randn(B, D, H, W, generator=g)
    xx = torch.stack([x.reshape(B, D, L), x.transpose(2, 3).reshape(B, D, L)], 1).half().to(DEV)
    delta = (0.4 * torch.rand(B, K, D, L, generator=g) + 0.01).half().to(DEV)
    bc = torch.randn(B, K, 2 * N, L, generator=g).half().to(DEV)
    A = (-0.5 * torch.rand(K * D, N, generator=g)).to(DEV)
    ys, _ = scan_forward(xx.view(B, 2 * D, L), delta.view(B, K * D, L), A, bc[:, :, :N], bc[:, :, N:], None, None, None, False,
                         True, u_group_div=2, reverse_group_mask=S.REVERSE_MASK)
    ys = ys.view(B, K, D, L)
    want = (ys[:, 0] + ys[:, 1]).view(B, D, H, W) + (ys[:, 2] + ys[:, 3]).view(B, D, W, H).transpose(2, 3)
    got = S.ss2d_core(xx, delta, A, bc[:, :, :N], bc[:, :, N:], None, None, H, W, False)
    assert_close(got.cpu().numpy(), want.cpu().numpy(), 2e-5, "ss2d_core without softplus / D / bias")
    with pytest.raises(RuntimeError):
        S.ss2d_core(xx, delta, A, bc[:, :, :N], bc[:, :, N:], None, None, H, W + 4, False)


# ------------------------------------------------------------------------------------------ Linear + residual + LayerNorm (f2)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,K,N,bias", [(1000, 96, 96, False), (128 * 149 + 5, 384, 96, True), (777, 192, 192, False),
                                         (4096, 768, 192, True), (513, 384, 384, False), (2000, 1536, 384, True), (1, 96, 96, True)])
def test_linear_res_ln_matches_torch(M, K, N, bias, dtype):
    """xp_linear_res_ln (tcgen05 GEMM with the residual add and the next LayerNorm in its epilogue) against
    x + F.linear -> F.layer_norm in fp32 on the same 16-bit operands (VMamba.py:664 / :110-128 + :1222-1234)."""
    from xpoint_b200.cross_scan import linear_res_ln
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(dtype).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).to(DEV)
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    res = (3.0 * torch.randn(M, N, generator=g) + 0.7).to(DEV)
    gam, bet = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV), (0.3 * torch.randn(N, generator=g)).to(DEV)
    s, y = linear_res_ln(x, W, b, res, gam, bet, 1e-5)
    s_ref = res.double() + x.double() @ W.double().t() + (b.double() if bias else 0.0)
    y_ref = torch.nn.functional.layer_norm(s_ref, (N,), gam.double(), bet.double(), 1e-5)
    assert s.dtype == torch.float32 and y.dtype == dtype and s.shape == (M, N) and y.shape == (M, N)
    assert_close(s.cpu().numpy(), s_ref.cpu().numpy(), 2e-6, f"linear_res_ln sum M={M} K={K} N={N}")
    assert_close(y.float().cpu().numpy(), y_ref.cpu().numpy(), 1e-3 if dtype == torch.float16 else 6e-3, f"linear_res_ln y {dtype}")
    _, y2 = linear_res_ln(x, W, b, res, gam, bet, 1e-5, want_sum=False)
    assert torch.equal(y, y2)


# ------------------------------------------------------------------------------------------ whole Mlp branch in one kernel (f2)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,C,bias", [(1000, 96, True), (128 * 149 + 5, 96, True), (128 * 300, 96, False), (777, 192, True),
                                       (128 * 160 + 77, 192, False), (1, 96, True), (129, 192, True)])
def test_mlp_res_ln_matches_torch(M, C, bias, dtype):
    """xp_mlp_res_ln (fc1 + GELU + fc2 + residual + next LayerNorm, hidden activation kept in shared memory) against the same
    chain in float64 with the hidden activation rounded to the 16-bit dtype, as autocast does between Mlp.fc1 and Mlp.fc2
    (VMamba.py:110-128 + :1229-1234)."""
    from xpoint_b200.cross_scan import mlp_res_ln
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, C, generator=g).to(dtype).to(DEV)
    W1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(dtype).to(DEV)
    W2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(dtype).to(DEV)
    b1 = (0.5 * torch.randn(4 * C, generator=g)).to(DEV) if bias else None
    b2 = torch.randn(C, generator=g).to(DEV) if bias else None
    res = (3.0 * torch.randn(M, C, generator=g) + 0.7).to(DEV)
    gam, bet = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV), (0.3 * torch.randn(C, generator=g)).to(DEV)
    s, y = mlp_res_ln(x, W1, b1, W2, b2, res, gam, bet, 1e-5)
    h = torch.nn.functional.gelu(x.double() @ W1.double().t() + (b1.double() if bias else 0.0)).to(dtype).double()
    s_ref = res.double() + h @ W2.double().t() + (b2.double() if bias else 0.0)
    y_ref = torch.nn.functional.layer_norm(s_ref, (C,), gam.double(), bet.double(), 1e-5)
    assert s.dtype == torch.float32 and y.dtype == dtype and s.shape == (M, C) and y.shape == (M, C)
    # the only rounding difference: hidden values that sit on a 16-bit rounding boundary (fp32 vs fp64 fc1 accumulation)
    tol_s = 2e-4 if dtype == torch.float16 else 1.5e-3
    assert_close(s.cpu().numpy(), s_ref.cpu().numpy(), tol_s, f"mlp_res_ln sum M={M} C={C}")
    assert_close(y.float().cpu().numpy(), y_ref.cpu().numpy(), 1.5e-3 if dtype == torch.float16 else 8e-3, f"mlp_res_ln y {dtype}")
    _, y2 = mlp_res_ln(x, W1, b1, W2, b2, res, gam, bet, 1e-5, want_sum=False)
    assert torch.equal(y, y2)
    # without a LayerNorm (a stage's last block): y is the sum rounded to the 16-bit dtype
    _, y4 = mlp_res_ln(x, W1, b1, W2, b2, res, None, None, want_sum=False)
    assert torch.equal(y4, s.to(dtype))
    # against the two-kernel path of the library (linear_act + linear_res_ln): same operands, same rounding points
    from xpoint_b200.cross_scan import linear_act, linear_res_ln
    s3, y3 = linear_res_ln(linear_act(x, W1, b1, gelu=True), W2, b2, res, gam, bet, 1e-5)
    assert_close(s.cpu().numpy(), s3.cpu().numpy(), tol_s, "mlp_res_ln vs linear_act + linear_res_ln")
