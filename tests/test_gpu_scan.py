"""Parity of the CUDA selective scan (through the C ABI) with the oracle and the reference goldens.  -m gpu."""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden, rel_errs

pytestmark = pytest.mark.gpu

FP32_REL = 1e-4   # BASELINE.json north_star: fp32 rel 1e-4
B16_REL = 1e-2    # bf16 rel 1e-2 (fp16 held to the same bar)
DEV = "cuda"


def _imports():
    import xpoint_b200 as X
    from oracle import oracle as O
    return X, O


def make_inputs(seed, Bt, KD, K, N, L, dtype=torch.float32, KD1=None):
    """Distributions of the reference's kernel test (test_selective_scan.py:414-444)."""
    g = torch.Generator().manual_seed(seed)
    KD1 = KD1 or KD
    A = -0.5 * torch.rand(KD, N, generator=g)
    B = torch.randn(Bt, K, N, L, generator=g).to(dtype)
    C = torch.randn(Bt, K, N, L, generator=g).to(dtype)
    D = torch.randn(KD, generator=g)
    z = torch.randn(Bt, KD, L, generator=g).to(dtype)
    bias = 0.5 * torch.rand(KD1, generator=g)
    u = torch.randn(Bt, KD, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(Bt, KD1, L, generator=g)).to(dtype)
    return u, delta, A, B, C, D, z, bias


def cu(*ts):
    return [None if t is None else t.to(DEV) for t in ts]


@pytest.mark.parametrize("name", ["scan_n16_k4", "scan_n1_k4", "scan_n8_k2", "scan_dgroups", "scan_3d", "scan_z_last",
                                  "scan_bf16", "scan_fp16"])
def test_golden_vectors(name):
    X, _ = _imports()
    g = golden(name)
    dt = {"scan_bf16": torch.bfloat16, "scan_fp16": torch.float16}.get(name, torch.float32)
    t = lambda k: None if k not in g else torch.from_numpy(g[k]).to(DEV)
    tin = lambda k: None if k not in g else torch.from_numpy(g[k]).to(DEV).to(dt)
    sp = bool(g["delta_softplus"])
    if name in ("scan_z_last", "scan_dgroups", "scan_3d"):
        out, last = X.selective_scan_fn(tin("u"), tin("delta"), t("A"), tin("B"), tin("C"), t("D"), z=tin("z"),
                                        delta_bias=t("delta_bias"), delta_softplus=sp, return_last_state=True)
        assert_close(last.cpu().numpy(), g["last_state"], FP32_REL, name + " last_state")
    else:
        out = X.selective_scan_fn(tin("u"), tin("delta"), t("A"), tin("B"), tin("C"), t("D"), t("delta_bias"), sp, True)
        assert out.dtype == torch.float32
    assert_close(out.float().cpu().numpy(), g["out"], FP32_REL if dt == torch.float32 else B16_REL, name)
    if "out_indtype" in g:   # oflex=False returns the input dtype (csms6s.py:68)
        o2 = X.selective_scan_fn(tin("u"), tin("delta"), t("A"), tin("B"), tin("C"), t("D"), t("delta_bias"), sp, False)
        assert o2.dtype == dt
        assert_close(o2.float().cpu().numpy(), g["out_indtype"], B16_REL, name + " in-dtype out")
    # reference's own kernel-test tolerance
    rt, at = {torch.float32: (6e-4, 2e-3), torch.float16: (3e-3, 5e-3), torch.bfloat16: (3e-2, 5e-2)}[dt]
    np.testing.assert_allclose(out.float().cpu().numpy(), g["out"], rtol=rt, atol=at)


SEQLENS = [64, 65, 128, 256, 512, 1011, 1024, 2048, 4096, 320, 1280, 5120]


@pytest.mark.parametrize("N", [1, 2, 4, 8, 16, 3, 32])
@pytest.mark.parametrize("L", SEQLENS)
def test_fp32_vs_oracle_lengths(N, L):
    X, O = _imports()
    K = 4 if N != 3 else 2
    u, dl, A, B, C, D, z, bias = make_inputs(L + N, 2, 24 * K // 2, K, N, L)
    out = X.selective_scan_fn(*cu(u, dl, A, B, C, D, bias), True, True)
    ref = O.selective_scan(u, dl, A, B, C, D, None, bias, True)
    assert_close(out.cpu().numpy(), ref, FP32_REL, f"N={N} L={L}")


@pytest.mark.parametrize("N,KD,K", [(1, 384, 4), (16, 768, 4), (1, 96, 1), (16, 192, 2), (8, 40, 4), (2, 64, 2)])
@pytest.mark.parametrize("flags", ["all", "noD", "nobias", "nosoftplus", "z", "bare"])
def test_fp32_flags_and_last_state(N, KD, K, flags):
    X, O = _imports()
    L = 20480 if (N, KD) in ((1, 384), (16, 768)) and flags == "all" else 1536
    u, dl, A, B, C, D, z, bias = make_inputs(7, 2, KD, K, N, L)
    use_D = flags not in ("noD", "bare")
    use_b = flags not in ("nobias", "bare")
    sp = flags not in ("nosoftplus", "bare")
    use_z = flags == "z"
    args = cu(u, dl, A, B, C, D if use_D else None, z if use_z else None, bias if use_b else None)
    out, last = X.selective_scan_fn_mamba(*args, delta_softplus=sp, return_last_state=True)
    ref, rlast = O.selective_scan(u, dl, A, B, C, D if use_D else None, z if use_z else None, bias if use_b else None, sp,
                                  return_last_state=True)
    assert_close(out.cpu().numpy(), ref, FP32_REL, f"out {N},{KD},{K},{flags}")
    assert_close(last.cpu().numpy(), rlast, FP32_REL, f"last {N},{KD},{K},{flags}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("N,KD,K,L", [(1, 384, 4, 5120), (16, 192, 4, 2048), (2, 48, 2, 1024), (8, 64, 4, 328), (5, 20, 2, 100)])
@pytest.mark.parametrize("out_float", [True, False])
def test_16bit_inputs(dtype, N, KD, K, L, out_float):
    """Both sides get the same pre-rounded tensors."""
    X, O = _imports()
    u, dl, A, B, C, D, z, bias = make_inputs(3, 2, KD, K, N, L, dtype=dtype)
    out = X.selective_scan_fn(*cu(u, dl, A, B, C, D, bias), True, out_float)
    assert out.dtype == (torch.float32 if out_float else dtype)
    ref = O.selective_scan(u, dl, A, B, C, D, None, bias, True)
    assert_close(out.float().cpu().numpy(), ref, FP32_REL if out_float else B16_REL, f"{dtype} N={N}")


@pytest.mark.parametrize("N,KD,K,L", [(1, 96, 4, 2048), (16, 128, 4, 1024), (4, 64, 2, 512)])
def test_generic_kernel_matches_fast_kernel(N, KD, K, L):
    X, O = _imports()
    from xpoint_b200.selective_scan import scan_forward
    u, dl, A, B, C, D, z, bias = make_inputs(11, 2, KD, K, N, L)
    fast, lf = scan_forward(*cu(u, dl, A, B, C, D, z, bias), True, True, True)
    gen, lg = scan_forward(*cu(u, dl, A, B, C, D, z, bias), True, True, True, force_generic=True)
    ref, rl = O.selective_scan(u, dl, A, B, C, D, z, bias, True, return_last_state=True)
    assert_close(fast.cpu().numpy(), ref, FP32_REL, "fast")
    assert_close(gen.cpu().numpy(), ref, FP32_REL, "generic")
    assert_close(lf.cpu().numpy(), rl, FP32_REL, "fast last")
    assert_close(lg.cpu().numpy(), rl, FP32_REL, "generic last")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,L", [(1, 320), (1, 1280), (1, 2056), (1, 5120), (2, 776), (4, 264), (1, 24)])
@pytest.mark.parametrize("use_z", [False, True])
def test_fused_cross_scan_addressing(dtype, N, L, use_z):
    """xp_scan_args.u_group_div / reverse_group_mask (ABI 2): directions that share one copy of the activations and
    walk memory backwards (CrossScan / CrossMerge without copies, csm_triton.py:22-29,56-62).  The oracle sees the
    materialised equivalent: u repeated per group, reversed groups flipped along L and flipped back."""
    X, O = _imports()
    from xpoint_b200.selective_scan import scan_forward
    Bt, K, Dg, div, mask = 2, 4, 24, 2, 0b1010
    u, dl, A, B, C, D, z, bias = make_inputs(L + 3 * N, Bt, K * Dg, K, N, L, dtype=dtype)
    usrc = u[:, : (K // div) * Dg].contiguous()
    ufull = torch.cat([usrc[:, (g // div) * Dg:(g // div + 1) * Dg] for g in range(K)], dim=1)
    rev = [bool((mask >> g) & 1) for g in range(K)]

    def flip_rows(t):      # (Bt, K*Dg, L)
        t = t.clone().view(Bt, K, Dg, L)
        for g in range(K):
            if rev[g]:
                t[:, g] = t[:, g].flip(-1)
        return t.view(Bt, K * Dg, L)

    def flip_groups(t):    # (Bt, K, N, L)
        t = t.clone()
        for g in range(K):
            if rev[g]:
                t[:, g] = t[:, g].flip(-1)
        return t

    zz = z if use_z else None
    ref, rlast = O.selective_scan(flip_rows(ufull), flip_rows(dl), A, flip_groups(B), flip_groups(C), D,
                                  None if zz is None else flip_rows(zz), bias, True, return_last_state=True)
    ref = flip_rows(torch.from_numpy(np.ascontiguousarray(ref))).numpy()
    tol = FP32_REL if dtype == torch.float32 else B16_REL
    for force_generic in (False, True):
        out, last = scan_forward(*cu(usrc, dl, A, B, C, D, zz, bias), True, True, True, force_generic=force_generic,
                                 u_group_div=div, reverse_group_mask=mask)
        assert_close(out.cpu().numpy(), ref, tol, f"fused addressing {dtype} N={N} L={L} generic={force_generic}")
        assert_close(last.cpu().numpy(), rlast, tol, "fused addressing last state")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,R,L", [(1, 6, 5120), (1, 12, 1280), (1, 6, 328), (1, 7, 2056), (2, 16, 776), (1, 1, 64), (1, 3, 24),
                                   (1, 24, 320), (16, 6, 512), (4, 5, 100)])
@pytest.mark.parametrize("addressing", ["plain", "ss2d"])
def test_fused_dt_proj(dtype, N, R, L, addressing):
    """xp_scan_args.dt_weight / dt_rank (ABI 3): the scan forms delta = dt_projs_weight x dts_r itself
    (VMamba.py:605-615) from a strided view of the x_proj output.  The oracle gets the materialised fp32 delta.  Ranks
    <= 16 with N <= 2 and 16-bit inputs run on the lanes kernel (mma.sync micro-GEMM), everything else on the generic kernel; "ss2d"
    adds the shared-u / reversed-group addressing of the copy-free SS2D path."""
    X, O = _imports()
    from xpoint_b200.selective_scan import scan_forward
    Bt, K, Dg = 2, 4, 24
    u, _, A, _, _, D, _, bias = make_inputs(L + 5 * N + R, Bt, K * Dg, K, N, L, dtype=dtype)
    g = torch.Generator().manual_seed(R)
    x_dbl = torch.randn(Bt, K, R + 2 * N, L, generator=g).to(dtype)        # the fused x_proj output (VMamba.py:606)
    x_dbl[:, :, :R] *= 0.3
    wdt = (torch.randn(K * Dg, R, generator=g) * R ** -0.5).to(dtype)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    delta = torch.from_numpy(O.dt_proj(dts, wdt)).to(dtype).float()     # rounded like the reference's dt_proj output
    div, mask = (2, 0b1010) if addressing == "ss2d" else (1, 0)
    rev = [bool((mask >> k) & 1) for k in range(K)]
    usrc = u[:, : (K // div) * Dg].contiguous()
    ufull = torch.cat([usrc[:, (k // div) * Dg:(k // div + 1) * Dg] for k in range(K)], dim=1)

    def flip(t, groups_dim_rows):       # flip the reversed groups along L
        t = t.clone().float()
        v = t.view(Bt, K, -1, L)
        for k in range(K):
            if rev[k]:
                v[:, k] = v[:, k].flip(-1)
        return t

    ref, rlast = O.selective_scan(flip(ufull, True), flip(delta, True), A, flip(Bs.contiguous(), False), flip(Cs.contiguous(), False),
                                  D, None, bias, True, return_last_state=True)
    ref = flip(torch.from_numpy(np.ascontiguousarray(ref)), True).numpy()
    xd = x_dbl.to(DEV)
    dts_d, Bs_d, Cs_d = torch.split(xd, [R, N, N], dim=2)
    for force_generic in (False, True):
        out, last = scan_forward(usrc.to(DEV), dts_d, A.to(DEV), Bs_d, Cs_d, D.to(DEV), None, bias.to(DEV), True, True, True,
                                 force_generic=force_generic, u_group_div=div, reverse_group_mask=mask, dt_weight=wdt.to(DEV))
        # identical pre-rounded factors on both sides; a different fp32 summation order can move a 16-bit delta by one ulp
        tol = FP32_REL if dtype == torch.float32 else 2e-3
        assert_close(out.cpu().numpy(), ref, tol, f"fused dt_proj {dtype} N={N} R={R} L={L} generic={force_generic}")
        assert_close(last.cpu().numpy(), rlast, tol, "fused dt_proj last state")


def test_fused_dt_proj_errors():
    X, _ = _imports()
    from xpoint_b200.selective_scan import scan_forward
    u, dl, A, B, C, D, z, bias = cu(*make_inputs(0, 2, 8, 2, 1, 16))
    dts = torch.randn(2, 2, 3, 16, device=DEV)
    w = torch.randn(8, 3, device=DEV)
    assert scan_forward(u, dts, A, B, C, D, None, bias, True, dt_weight=w)[0].shape == (2, 8, 16)
    with pytest.raises(RuntimeError):
        scan_forward(u, dts, A, B, C, D, None, bias, True, dt_weight=w[:, :2])       # rank mismatch
    with pytest.raises(RuntimeError):
        scan_forward(u, dts[:, :1], A, B, C, D, None, bias, True, dt_weight=w)       # groups mismatch
    with pytest.raises(RuntimeError):
        scan_forward(u, dts, A, B, C, D, None, bias, True, dt_weight=w.half())       # dtype mismatch


def test_delta_groups_and_strided_views():
    X, O = _imports()
    u, dl, A, B, C, D, z, bias = make_inputs(5, 2, 24, 2, 4, 96, KD1=6)
    out = X.selective_scan_fn_mamba(*cu(u, dl, A, B, C, D, None, bias), delta_softplus=True)
    ref = O.selective_scan(u, dl, A, B, C, D, None, bias, True)
    assert_close(out.cpu().numpy(), ref, FP32_REL, "delta groups")
    # B/C as strided views of a fused projection, as SS2D produces them (VMamba.py:606)
    Bt, K, N, R, L, Dn = 2, 4, 16, 6, 512, 32
    g = torch.Generator().manual_seed(1)
    x_dbl = torch.randn(Bt, K, R + 2 * N, L, generator=g).to(DEV)
    _, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    u, dl, A, _, _, D, _, bias = make_inputs(6, Bt, K * Dn, K, N, L)
    out = X.selective_scan_fn(*cu(u, dl, A), Bs, Cs, *cu(D, bias), True, True)
    ref = O.selective_scan(u, dl, A, Bs.cpu(), Cs.cpu(), D, None, bias, True)
    assert_close(out.cpu().numpy(), ref, FP32_REL, "strided B/C")
    # batch-strided u (every other batch row)
    u2 = torch.randn(4, K * Dn, L, generator=g).to(DEV)
    out = X.selective_scan_fn(u2[::2], *cu(dl, A), Bs, Cs, *cu(D, bias), True, True)
    ref = O.selective_scan(u2[::2].cpu(), dl, A, Bs.cpu(), Cs.cpu(), D, None, bias, True)
    assert_close(out.cpu().numpy(), ref, FP32_REL, "strided u")


def test_edge_cases_and_errors():
    X, _ = _imports()
    u, dl, A, B, C, D, z, bias = cu(*make_inputs(0, 2, 8, 2, 4, 16))
    assert X.selective_scan_fn(u[:0], dl[:0], A, B[:0], C[:0], D, bias).shape == (0, 8, 16)
    assert X.selective_scan_fn(u[..., :0], dl[..., :0], A, B[..., :0], C[..., :0], D, bias).shape == (2, 8, 0)
    with pytest.raises(RuntimeError):
        X.selective_scan_fn(u, dl.half(), A, B, C, D, bias)            # dtype mismatch
    with pytest.raises(RuntimeError):
        X.selective_scan_fn(u, dl, A[:, :2], B, C, D, bias)            # A shape
    with pytest.raises(RuntimeError):
        X.selective_scan_fn(u, dl, A, B, C, D.half(), bias)            # D must be fp32
    with pytest.raises(RuntimeError):
        X.selective_scan_fn(u.cpu(), dl, A, B, C, D, bias)             # no CPU path
    with pytest.raises(RuntimeError):
        X.selective_scan_fn(u, dl, A, B, C, D, bias, backend="torch")  # no torch fallback
    with pytest.raises(NotImplementedError):
        X.selective_scan_cuda_oflex.bwd()
    out, x = X.selective_scan_cuda_oflex.fwd(u, dl, A, B, C, D, bias, True, 1, True)
    assert out.dtype == torch.float32 and out.shape == u.shape


def test_full_size_properties_config2():
    """BASELINE config 2 shape (B=32, K*D=768, N=16, L=20480, fp32): sampled rows vs the oracle, causality and
    linearity in u -- size-independent properties, the oracle only sees 8 rows."""
    X, O = _imports()
    Bt, KD, K, N, L = 32, 768, 4, 16, 20480
    g = torch.Generator(device=DEV).manual_seed(0)
    u = torch.randn(Bt, KD, L, generator=g, device=DEV)
    dl = 0.5 * torch.rand(Bt, KD, L, generator=g, device=DEV)
    A = -0.5 * torch.rand(KD, N, generator=g, device=DEV)
    Bm = torch.randn(Bt, K, N, L, generator=g, device=DEV)
    Cm = torch.randn(Bt, K, N, L, generator=g, device=DEV)
    D = torch.randn(KD, generator=g, device=DEV)
    bias = 0.5 * torch.rand(KD, generator=g, device=DEV)
    out = X.selective_scan_fn(u, dl, A, Bm, Cm, D, bias, True, True)
    rows = [(0, 0), (0, 191), (5, 192), (13, 400), (31, 767), (17, 575), (31, 0), (8, 383)]
    for b, d in rows:
        k = d // (KD // K)
        ref = O.selective_scan(u[b:b + 1, d:d + 1].cpu(), dl[b:b + 1, d:d + 1].cpu(), A[d:d + 1].cpu(), Bm[b:b + 1, k:k + 1].cpu(),
                               Cm[b:b + 1, k:k + 1].cpu(), D[d:d + 1].cpu(), None, bias[d:d + 1].cpu(), True)
        assert_close(out[b, d].cpu().numpy(), ref[0, 0], FP32_REL, f"row {b},{d}")
    # causality: a scan over the first 6000 tokens equals the prefix of the full scan
    Lp = 6000
    outp = X.selective_scan_fn(u[:4, :, :Lp].contiguous(), dl[:4, :, :Lp].contiguous(), A, Bm[:4, ..., :Lp].contiguous(),
                               Cm[:4, ..., :Lp].contiguous(), D, bias, True, True)
    assert_close(outp.cpu().numpy(), out[:4, :, :Lp].cpu().numpy(), 1e-5, "causality")
    # linearity in u
    out2 = X.selective_scan_fn(u[:4] * 2.0, dl[:4], A, Bm[:4], Cm[:4], D, bias, True, True)
    assert_close(out2.cpu().numpy(), 2.0 * out[:4].cpu().numpy(), 1e-5, "linearity")


def test_selective_scan_cuda_backend_output_dtype():
    """SelectiveScanCuda.apply maps `backend` like csms6s.py:76-85 -- None/'oflex' honour oflex, 'core'/'mamba'
    return the input dtype, 'torch' is refused (no fallback in this package)."""
    import xpoint_b200 as X
    g = torch.Generator().manual_seed(0)
    u = torch.randn(1, 8, 64, generator=g).half().cuda()
    dl = torch.rand(1, 8, 64, generator=g).half().cuda()
    A = -torch.rand(8, 2, generator=g).cuda()
    Bm = torch.randn(1, 1, 2, 64, generator=g).half().cuda()
    Cm = torch.randn(1, 1, 2, 64, generator=g).half().cuda()
    assert X.SelectiveScanCuda.apply(u, dl, A, Bm, Cm, None, None, True, True, None).dtype == torch.float32
    assert X.SelectiveScanCuda.apply(u, dl, A, Bm, Cm, None, None, True, True, "oflex").dtype == torch.float32
    assert X.SelectiveScanCuda.apply(u, dl, A, Bm, Cm, None, None, True, False, None).dtype == torch.float16
    assert X.SelectiveScanCuda.apply(u, dl, A, Bm, Cm, None, None, True, True, "mamba").dtype == torch.float16
    assert X.SelectiveScanCuda.apply(u, dl, A, Bm, Cm, None, None, True, True, "core").dtype == torch.float16
    assert X.selective_scan_fn_csms6s(u, dl, A, Bm, Cm, backend="core").dtype == torch.float16
    with pytest.raises(RuntimeError):
        X.SelectiveScanCuda.apply(u, dl, A, Bm, Cm, None, None, True, True, "torch")
