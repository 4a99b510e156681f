"""Selective-scan operator API, call-compatible with the reference.

Mirrors xpoint/models/vmamba_src/csms6s.py:
  * ``selective_scan_fn(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, oflex=True, backend=None)``
    (csms6s.py:112-126) -- the importable XPoint/VMamba signature;
  * the mamba_ssm-style signature used by the reference's kernel tests
    ``selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
    return_last_state=False)`` (kernels/selective_scan/test_selective_scan.py:152,309) is accepted by the same
    function (a 3-D seventh argument / ``z=`` / ``return_last_state=`` selects it) and by ``selective_scan_fn_mamba``;
  * ``selective_scan_cuda_oflex`` -- an object with the ``fwd`` / ``bwd`` entry points of the native module the
    reference imports (selective_scan_oflex.cpp:143-151,233-242,357-360).

Everything runs through the C-ABI call ``xp_selective_scan_fwd`` (include/xpoint_b200.h).  There is no torch
fallback: ``backend="torch"`` raises, because this package ships only the CUDA path.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib


def _check_shapes(u, delta, A, B, C, D, z, delta_bias, u_group_div=1, dt_weight=None):
    if u.dim() != 3:
        raise RuntimeError("u must have shape (batch, dim, seqlen)")
    batch, dim, L = u.shape
    if u_group_div > 1:       # fused CrossScan addressing: u holds groups/u_group_div distinct sources
        dim = dim * u_group_div
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    if B.dim() != 4 or C.dim() != 4:
        raise RuntimeError("B and C must have shape (batch, groups, dstate, seqlen) or (batch, dstate, seqlen)")
    groups, N = B.shape[1], B.shape[2]
    if dt_weight is not None:     # fused dt_proj: delta holds the low-rank factors dts_r (batch, groups, dt_rank, seqlen)
        if delta.dim() != 4 or delta.shape[0] != batch or delta.shape[1] != groups or delta.shape[3] != L:
            raise RuntimeError(f"with dt_weight, delta must be dts_r (batch, groups, dt_rank, seqlen), got {tuple(delta.shape)}")
        if tuple(dt_weight.shape) != (dim, delta.shape[2]) or dt_weight.dtype != u.dtype:
            raise RuntimeError(f"dt_weight must have shape (dim, dt_rank) = ({dim}, {delta.shape[2]}) and u's dtype")
        if not 1 <= delta.shape[2] <= 64:
            raise RuntimeError("fused dt_proj supports dt_rank 1..64")
        ddim = dim
    else:
        if delta.dim() != 3 or delta.shape[0] != batch or delta.shape[2] != L:
            raise RuntimeError(f"delta must have shape (batch, delta_dim, seqlen), got {tuple(delta.shape)}")
        ddim = delta.shape[1]
    if tuple(A.shape) != (dim, N):
        raise RuntimeError(f"A must have shape (dim, dstate) = ({dim}, {N}), got {tuple(A.shape)}")
    if tuple(B.shape) != (batch, groups, N, L) or tuple(C.shape) != (batch, groups, N, L):
        raise RuntimeError("B and C must both have shape (batch, groups, dstate, seqlen)")
    if dim % groups != 0:
        raise RuntimeError("dim must be divisible by the number of B/C groups")
    if dim % ddim != 0:
        raise RuntimeError("dim must be divisible by delta's channel count")
    if N > 256:
        raise RuntimeError("selective_scan only supports state dimension <= 256")
    if A.dtype != torch.float32:
        raise RuntimeError("A must be float32")
    for name, t in (("delta", delta), ("B", B), ("C", C), ("z", z)):
        if t is not None and t.dtype != u.dtype:
            raise RuntimeError(f"{name} must have the same dtype as u ({u.dtype}), got {t.dtype}")
    if D is not None and (D.dtype != torch.float32 or tuple(D.shape) != (dim,)):
        raise RuntimeError("D must be float32 with shape (dim,)")
    if delta_bias is not None and (delta_bias.dtype != torch.float32 or tuple(delta_bias.shape) != (ddim,)):
        raise RuntimeError("delta_bias must be float32 with shape (delta_dim,)")
    if z is not None and tuple(z.shape) != (batch, dim, L):
        raise RuntimeError("z must have the same shape as u")
    if u_group_div > 1 and (groups % u_group_div != 0 or ddim != dim):
        raise RuntimeError("u_group_div must divide the number of groups and needs a full-width delta")
    return B, C, batch, dim, ddim, groups, N, L


def _last_contig(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0 or t.stride(-1) == 1:
        return t
    return t.contiguous()


def scan_forward(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, out_float=True,
                 return_last_state=False, force_generic=False, u_group_div=1, reverse_group_mask=0, dt_weight=None):
    """One call of xp_selective_scan_fwd.  Returns (out, last_state or None).

    ``u_group_div`` / ``reverse_group_mask`` expose the fused-CrossScan addressing of the C ABI (xp_scan_args):
    with ``u_group_div = q > 1`` u has shape (batch, dim / q, seqlen) and groups g, g+1, .. g+q-1 (g % q == 0) all read
    source g / q; bit g of the mask makes group g run backwards through memory (its output stays unflipped).

    ``dt_weight`` (dim, dt_rank) selects the fused dt_proj of the C ABI: ``delta`` is then the low-rank
    dts_r (batch, groups, dt_rank, seqlen) of VMamba.py:605-615 and delta = dt_weight x dts_r is formed inside the scan."""
    dev = _lib.require_cuda(u, delta, A, B, C, D, z, delta_bias)
    B, C, batch, dim, ddim, groups, N, L = _check_shapes(u, delta, A, B, C, D, z, delta_bias, u_group_div, dt_weight)
    u, delta, B, C, z = map(_last_contig, (u, delta, B, C, z))
    A = A.contiguous()
    D = None if D is None else D.contiguous()
    delta_bias = None if delta_bias is None else delta_bias.contiguous()
    out_dtype = torch.float32 if out_float else u.dtype
    out = torch.empty((batch, dim, L), dtype=out_dtype, device=dev)
    last = torch.empty((batch, dim, N), dtype=torch.float32, device=dev) if return_last_state else None
    if batch == 0 or L == 0:
        if last is not None:
            last.zero_()
        return out, last
    a = _lib.ScanArgs()
    a.u, a.delta, a.A, a.B, a.C = u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
    a.D = 0 if D is None else D.data_ptr()
    a.z = 0 if z is None else z.data_ptr()
    a.delta_bias = 0 if delta_bias is None else delta_bias.data_ptr()
    a.out = out.data_ptr()
    a.last_state = 0 if last is None else last.data_ptr()
    a.batch, a.dim, a.delta_dim, a.groups, a.dstate, a.seqlen = batch, dim, ddim, groups, N, L
    a.u_batch_stride, a.u_dim_stride = u.stride(0), u.stride(1)
    if dt_weight is None:
        a.delta_batch_stride, a.delta_dim_stride = delta.stride(0), delta.stride(1)
    else:
        dt_weight = dt_weight.contiguous()
        a.delta_batch_stride, a.dt_group_stride, a.delta_dim_stride = delta.stride(0), delta.stride(1), delta.stride(2)
        a.dt_weight, a.dt_rank = dt_weight.data_ptr(), delta.shape[2]
    a.B_batch_stride, a.B_group_stride, a.B_state_stride = B.stride(0), B.stride(1), B.stride(2)
    a.C_batch_stride, a.C_group_stride, a.C_state_stride = C.stride(0), C.stride(1), C.stride(2)
    if z is not None:
        a.z_batch_stride, a.z_dim_stride = z.stride(0), z.stride(1)
    a.out_batch_stride, a.out_dim_stride = out.stride(0), out.stride(1)
    a.in_dtype, a.out_dtype = _lib.dtype_code(u), _lib.dtype_code(out)
    a.delta_softplus = int(bool(delta_softplus))
    a.force_generic = int(bool(force_generic))
    if u_group_div > 1:
        a.u_group_div = int(u_group_div)
        a.u_group_stride = (dim // groups) * u.stride(1)
    a.reverse_group_mask = int(reverse_group_mask)
    prof = _lib.scan_profile
    with torch.cuda.device(dev):
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(_lib.lib().xp_selective_scan_fwd(ctypes.byref(a), _lib.stream_ptr(dev)))
        if prof is not None:
            e1.record()
            rank = 0 if dt_weight is None else delta.shape[2]
            prof.append((e0, e1, algorithmic_bytes(batch, dim, groups, N, L, u.element_size(), out.element_size(), rank),
                         (batch, dim, groups, N, L, str(u.dtype), str(out.dtype), rank)))
    _lib.count_launches(1)
    return out, last


def algorithmic_bytes(batch, dim, groups, dstate, L, s_in, s_out, dt_rank=0):
    """Op-level bytes the selective_scan_fn contract reads and writes once:
    B*L*[(2*KD + 2*K*N)*s_in + KD*s_out] + 4*KD*(N+2).
    With the fused dt_proj the delta term KD*s_in becomes K*R*s_in (plus the (KD, R) weights once)."""
    delta_term = groups * dt_rank if dt_rank else dim
    return (batch * L * ((dim + delta_term + 2 * groups * dstate) * s_in + dim * s_out) + 4 * dim * (dstate + 2)
            + dim * dt_rank * s_in)


class _OflexModule:
    """Stand-in for the native module `selective_scan_cuda_oflex` (selective_scan_oflex.cpp:357-360)."""

    @staticmethod
    def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, out_float):
        del nrows  # accepted and ignored, as in the reference (selective_scan_oflex.cpp:150)
        out, last = scan_forward(u, delta, A, B, C, D, None, delta_bias, delta_softplus, out_float, return_last_state=True)
        # the reference returns [out, x] with x the per-chunk scan states used only by bwd; we hand back the final state
        return [out, last]

    @staticmethod
    def bwd(*_args, **_kwargs):
        _lib.check(_lib.lib().xp_selective_scan_bwd())


selective_scan_cuda_oflex = _OflexModule()


def _out_float(oflex, backend):
    """csms6s.py:76-85: backend None / 'oflex' honours `oflex` (fp32 output when set); 'core' and 'mamba' are the
    extension flavours whose output dtype is the input dtype; 'torch' is the reference's CPU/torch fallback."""
    if backend == "torch":
        raise RuntimeError("xpoint_b200 ships only the CUDA path; backend='torch' (the reference's "
                           "selective_scan_torch fallback) is not available")
    if backend not in (None, "oflex", "core", "mamba"):
        raise ValueError(f"unknown selective-scan backend {backend!r}")
    return bool(oflex) if backend in (None, "oflex") else False


class SelectiveScanCuda(torch.autograd.Function):
    """Same name / argument order as csms6s.py:71-87.  Forward only (inference tier)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, oflex=True, backend=None):
        out, _ = scan_forward(u, delta, A, B, C, D, None, delta_bias, delta_softplus, out_float=_out_float(oflex, backend))
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        raise NotImplementedError("xpoint_b200 is inference-only: selective-scan backward is not implemented")


def selective_scan_fn_mamba(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                            return_last_state=False):
    """mamba_ssm-style call (test_selective_scan.py:152-160): output in the input dtype, optional z gate and
    last state."""
    out, last = scan_forward(u, delta, A, B, C, D, z, delta_bias, delta_softplus, out_float=False,
                             return_last_state=return_last_state)
    return (out, last) if return_last_state else out


def selective_scan_fn_csms6s(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, oflex=True, backend=None):
    """Explicit entry point with the importable reference signature (csms6s.py:112-126); no argument sniffing."""
    return SelectiveScanCuda.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, oflex, backend)


def selective_scan_fn(u, delta, A, B, C, D=None, *args, **kwargs):
    """Drop-in for csms6s.selective_scan_fn; also accepts the mamba_ssm-style argument list (module docstring).
    The two styles are told apart by what follows D: a 3-D tensor (z) / a tensor-or-None in the second slot / the
    keywords z= or return_last_state= mean mamba style; everything else is the csms6s style.  Callers that want no
    guessing use selective_scan_fn_csms6s / selective_scan_fn_mamba."""
    mamba_style = "z" in kwargs or "return_last_state" in kwargs
    if args:
        first = args[0]
        if torch.is_tensor(first) and first.dim() == 3:
            mamba_style = True
        if len(args) >= 2 and (args[1] is None or torch.is_tensor(args[1])):
            mamba_style = True
    if mamba_style:
        return selective_scan_fn_mamba(u, delta, A, B, C, D, *args, **kwargs)
    names = ("delta_bias", "delta_softplus", "oflex", "backend")
    params = dict(delta_bias=None, delta_softplus=True, oflex=True, backend=None)
    if len(args) > len(names):
        raise TypeError("selective_scan_fn: too many positional arguments")
    params.update(zip(names, args))
    for k, v in kwargs.items():
        if k not in params:
            raise TypeError(f"selective_scan_fn: unexpected keyword argument {k!r}")
        params[k] = v
    return SelectiveScanCuda.apply(u, delta, A, B, C, D, params["delta_bias"], params["delta_softplus"], params["oflex"],
                                   params["backend"])
