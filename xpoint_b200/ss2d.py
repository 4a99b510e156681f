"""Copy-free SS2D core: the host side of xp_ss2d_pack / xp_ss2d_dwconv_pack / xp_ss2d_merge_norm.

The reference's SS2D core (VMamba.py:493-646 ``forward_corev2``, :305-374 ``forwardv0``) materialises the four scan
orders with CrossScan, scans them, and sums them back with CrossMerge.  Here the selective-scan kernel routes the
directions itself (``scan_forward(..., u_group_div=2, reverse_group_mask=0b1010)``): it reads ONE copy of the
activations in row-major order and one in column-major order, walks each forwards and backwards, and leaves the four
outputs in natural memory order, so CrossScan shrinks to ``[x ; x^T]`` and CrossMerge + out_norm (+ gate) to one pass.

Direction order used throughout the fused path: ``FUSED_ORDER = (0, 2, 1, 3)`` i.e. [row forward, row backward,
column forward, column backward] in the reference's numbering k (csm_triton.py:22-29).
"""
from __future__ import annotations

import ctypes
import functools

import torch

from . import _lib

FUSED_ORDER = (0, 2, 1, 3)
REVERSE_MASK = 0b1010


def ss2d_pack(x: torch.Tensor) -> torch.Tensor:
    """x (B, D, H, W) -> xx (B, 2, D, H*W) = [x ; x^T]."""
    dev = _lib.require_cuda(x)
    x = x.contiguous()
    B, D, H, W = x.shape
    xx = torch.empty((B, 2, D, H * W), dtype=x.dtype, device=dev)
    if xx.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_ss2d_pack(_lib.ptr(x), _lib.ptr(xx), B, D, H, W, _lib.dtype_code(x), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return xx


def ss2d_dwconv_pack(x_cl: torch.Tensor, D: int, weight: torch.Tensor, bias, silu: bool = True) -> torch.Tensor:
    """x_cl (B, H, W, >=D) channel-last (a view of the in_proj output is fine: only the token stride must be uniform)
    -> xx (B, 2, D, H*W) = [act(dwconv3x3(x)) ; its transpose].  weight (D, 1, 3, 3) fp32, bias (D) fp32 or None."""
    dev = _lib.require_cuda(x_cl, weight, bias)
    B, H, W, Cin = x_cl.shape
    if x_cl.stride(3) != 1 or x_cl.stride(1) != W * x_cl.stride(2) or x_cl.stride(0) != H * x_cl.stride(1):
        x_cl = x_cl.contiguous()
    if tuple(weight.shape[-2:]) != (3, 3) or weight.shape[0] != D or D > Cin:
        raise RuntimeError("ss2d_dwconv_pack expects a depth-wise 3x3 weight of shape (D, 1, 3, 3)")
    xx = torch.empty((B, 2, D, H * W), dtype=x_cl.dtype, device=dev)
    if xx.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_ss2d_dwconv_pack(_lib.ptr(x_cl), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(xx), B, D, H, W,
                                                      x_cl.stride(2), _lib.dtype_code(x_cl), int(bool(silu)),
                                                      _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return xx


def ss2d_merge_norm(ys: torch.Tensor, H: int, W: int, weight: torch.Tensor, bias: torch.Tensor, zact=None, eps=1e-5,
                    out_dtype=None) -> torch.Tensor:
    """ys (B, 4, D, L) fp32, planes in FUSED_ORDER and natural memory order -> (B, H, W, D)."""
    dev = _lib.require_cuda(ys, weight, bias, zact)
    B, K, D, L = ys.shape
    if K != 4 or L != H * W or ys.dtype != torch.float32:
        raise RuntimeError("ss2d_merge_norm expects fp32 ys of shape (B, 4, D, H*W)")
    ys = ys.contiguous()
    out_dtype = out_dtype or (zact.dtype if zact is not None else ys.dtype)
    out = torch.empty((B, H, W, D), dtype=out_dtype, device=dev)
    if zact is not None:
        zact = zact.contiguous()
        if zact.dtype != out_dtype or zact.numel() != out.numel():
            raise RuntimeError("zact must match the output dtype and shape (B, H, W, D)")
    if out.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_ss2d_merge_norm(_lib.ptr(ys), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(zact),
                                                     _lib.ptr(out), B, D, H, W, _lib.dtype_code(out), float(eps),
                                                     _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


@functools.lru_cache(maxsize=None)
def core_channels(D: int, d_state: int, H: int, W: int, dtype: torch.dtype) -> int:
    """Channels of one image a CTA of the fused core (xp_ss2d_core) holds for this shape; 0 = the token planes do not fit
    in shared memory (the caller then scans with xp_selective_scan_fwd and merges with xp_ss2d_merge_norm)."""
    if d_state > 2 or H % 4 or W % 4 or (H * W) % (4 if dtype == torch.float32 else 8):
        return 0
    return int(_lib.lib().xp_ss2d_core_channels(D, d_state, H, W, _lib._DTYPES[dtype]))


def ss2d_core(xx: torch.Tensor, delta: torch.Tensor, A: torch.Tensor, Bs: torch.Tensor, Cs: torch.Tensor, Ds, delta_bias,
              H: int, W: int, delta_softplus: bool = True) -> torch.Tensor:
    """CrossScan -> selective scan -> CrossMerge in one kernel (VMamba.py:603-632): xx (B, 2, D, L) = [x ; x^T],
    delta (B, 4, D, L), Bs / Cs (B, 4, N, L) strided views (unit stride along L), A (4*D, N), Ds / delta_bias (4*D) fp32,
    all in FUSED_ORDER -> merged y (B, D, H, W) fp32."""
    dev = _lib.require_cuda(xx, delta, A, Bs, Cs, Ds, delta_bias)
    B, two, D, L = xx.shape
    N = Bs.shape[2]
    if two != 2 or L != H * W or tuple(delta.shape) != (B, 4, D, L) or tuple(Bs.shape) != (B, 4, N, L) or Cs.shape != Bs.shape:
        raise RuntimeError("ss2d_core: expected xx (B, 2, D, L), delta (B, 4, D, L), Bs / Cs (B, 4, N, L)")
    if not (xx.dtype == delta.dtype == Bs.dtype == Cs.dtype):
        raise RuntimeError("ss2d_core: xx, delta, Bs, Cs must share one dtype")
    if Bs.stride(3) != 1 or Cs.stride(3) != 1:
        raise RuntimeError("ss2d_core: Bs / Cs need unit stride along the sequence")
    if tuple(A.shape) != (4 * D, N) or A.dtype != torch.float32:
        raise RuntimeError("ss2d_core: A must be fp32 of shape (4*D, N)")
    xx, delta, A = xx.contiguous(), delta.contiguous(), A.contiguous()
    Ds = None if Ds is None else Ds.float().contiguous()
    delta_bias = None if delta_bias is None else delta_bias.float().contiguous()
    out = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
    if out.numel():
        a = _lib.CoreArgs()
        a.xx, a.delta, a.B, a.C = _lib.ptr(xx), _lib.ptr(delta), _lib.ptr(Bs), _lib.ptr(Cs)
        a.A, a.D, a.delta_bias, a.out = _lib.ptr(A), _lib.ptr(Ds), _lib.ptr(delta_bias), _lib.ptr(out)
        a.batch, a.d_inner, a.dstate, a.H, a.W = B, D, N, H, W
        a.B_batch_stride, a.B_group_stride, a.B_state_stride = Bs.stride(0), Bs.stride(1), Bs.stride(2)
        a.C_batch_stride, a.C_group_stride, a.C_state_stride = Cs.stride(0), Cs.stride(1), Cs.stride(2)
        a.in_dtype, a.delta_softplus = _lib.dtype_code(xx), int(bool(delta_softplus))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_ss2d_core(ctypes.byref(a), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


def ss2d_plane_norm(y: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, zact=None, eps=1e-5, out_dtype=None) -> torch.Tensor:
    """y (B, D, H, W) fp32 merged plane (ss2d_core) -> (B, H, W, D) = LayerNorm_D(y) [* zact]  (VMamba.py:641-646)."""
    dev = _lib.require_cuda(y, weight, bias, zact)
    if y.dim() != 4 or y.dtype != torch.float32:
        raise RuntimeError("ss2d_plane_norm expects an fp32 (B, D, H, W) plane")
    y = y.contiguous()
    B, D, H, W = y.shape
    out_dtype = out_dtype or (zact.dtype if zact is not None else y.dtype)
    out = torch.empty((B, H, W, D), dtype=out_dtype, device=dev)
    if zact is not None:
        zact = zact.contiguous()
        if zact.dtype != out_dtype or zact.numel() != out.numel():
            raise RuntimeError("zact must match the output dtype and shape (B, H, W, D)")
    if out.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_ss2d_plane_norm(_lib.ptr(y), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(zact), _lib.ptr(out),
                                                     B, D, H, W, _lib.dtype_code(out), float(eps), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


def ss2d_dt_proj(dts_r: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """dts_r (B, G, R, L) (a strided view of the x_proj output), weight (G, D, R) fp32 -> delta (B, G, D, L)."""
    dev = _lib.require_cuda(dts_r, weight)
    B, G, R, L = dts_r.shape
    D = weight.shape[1]
    if dts_r.stride(3) != 1 or tuple(weight.shape) != (G, D, R) or weight.dtype != torch.float32:
        raise RuntimeError("ss2d_dt_proj expects dts_r (B, G, R, L) with unit token stride and fp32 weight (G, D, R)")
    out = torch.empty((B, G, D, L), dtype=dts_r.dtype, device=dev)
    if out.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_ss2d_dt_proj(_lib.ptr(dts_r), _lib.ptr(weight.contiguous()), _lib.ptr(out), B, G, D, R, L,
                                                  dts_r.stride(0), dts_r.stride(1), dts_r.stride(2), _lib.dtype_code(dts_r),
                                                  _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


def dt_proj_supported(R: int, L: int, dtype: torch.dtype, batch_groups: int, D: int = 0) -> bool:
    """Ranks 1..8 (any dtype) and 9..16 (16-bit inputs; FHFMA kernel) run on xp_ss2d_dt_proj."""
    ok_rank = R <= 8 or (R <= 16 and dtype != torch.float32 and D <= 1536)
    return ok_rank and L % (4 if dtype == torch.float32 else 8) == 0 and batch_groups <= 65535


def fused_supported(H: int, W: int, D: int, d_state: int) -> bool:
    """Shapes the copy-free path covers (the rest takes the CrossScan -> scan -> CrossMerge kernels)."""
    # d_state 1 | 2: scan_lanes_kernel, 4 | 8 | 16: scan_rows_kernel -- both route the four directions themselves
    # (u_group_div / reverse_group_mask); other state sizes would take the generic kernel, which also supports the addressing
    # but is a correctness path
    return H % 4 == 0 and W % 4 == 0 and D <= 3072 and d_state in (1, 2, 4, 8, 16) and (H * W) % 8 == 0
