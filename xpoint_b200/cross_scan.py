"""CrossScan / CrossMerge operator API, call-compatible with xpoint/models/vmamba_src/csm_triton.py.

``cross_scan_fn`` / ``cross_merge_fn`` (csm_triton.py:501-517) and the ``CrossScanF`` / ``CrossMergeF`` autograd
Functions (csm_triton.py:182-273; the Triton twins ``CrossScanTritonF`` / ``CrossMergeTritonF`` :403-498 are aliases
here) keep their names, argument order and output shapes.  The work is done by ``xp_cross_scan`` /
``xp_cross_merge`` (include/xpoint_b200.h).  Forward only.
"""
from __future__ import annotations

import torch

from . import _lib


def cross_scan_fwd(x: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
    """x: (B,C,H,W) | (B,H,W,C) | (B,4,C,H,W) | (B,H,W,4,C)  ->  (B,4,C,L) | (B,L,4,C)."""
    dev = _lib.require_cuda(x)
    x = x.contiguous()
    if one_by_one:
        if in_channel_first:
            B, K, C, H, W = x.shape
        else:
            B, H, W, K, C = x.shape
        if K != 4:
            raise RuntimeError("one_by_one cross scan expects 4 directions")
    else:
        if in_channel_first:
            B, C, H, W = x.shape
        else:
            B, H, W, C = x.shape
    shape = (B, 4, C, H * W) if out_channel_first else (B, H * W, 4, C)
    xs = torch.empty(shape, dtype=x.dtype, device=dev)
    if xs.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_cross_scan(_lib.ptr(x), _lib.ptr(xs), B, C, H, W, _lib.dtype_code(x),
                                                int(bool(in_channel_first)), int(bool(out_channel_first)),
                                                int(bool(one_by_one)), int(scans), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return xs


def cross_merge_fwd(ys: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
    """ys: (B,4,C,H,W) if out_channel_first else (B,H,W,4,C)  ->  (B,C,L) | (B,L,C)  [(B,4,C,L) | (B,L,4,C)].

    Flag names follow the reference: ``out_channel_first`` describes ys, ``in_channel_first`` the result
    (csm_triton.py:56-85)."""
    dev = _lib.require_cuda(ys)
    ys = ys.contiguous()
    if out_channel_first:
        B, K, C, H, W = ys.shape
    else:
        B, H, W, K, C = ys.shape
    if K != 4:
        raise RuntimeError("cross merge expects 4 directions")
    L = H * W
    if one_by_one:
        shape = (B, 4, C, L) if in_channel_first else (B, L, 4, C)
    else:
        shape = (B, C, L) if in_channel_first else (B, L, C)
    y = torch.empty(shape, dtype=ys.dtype, device=dev)
    if y.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_cross_merge(_lib.ptr(ys), _lib.ptr(y), B, C, H, W, _lib.dtype_code(ys),
                                                 int(bool(out_channel_first)), int(bool(in_channel_first)),
                                                 int(bool(one_by_one)), int(scans), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return y


class CrossScanF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
        return cross_scan_fwd(x, in_channel_first, out_channel_first, one_by_one, scans)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("xpoint_b200 is inference-only: CrossScan backward is not implemented")


class CrossMergeF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ys, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
        return cross_merge_fwd(ys, in_channel_first, out_channel_first, one_by_one, scans)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("xpoint_b200 is inference-only: CrossMerge backward is not implemented")


CrossScanTritonF = CrossScanF    # csm_triton.py:403 -- same contract, one CUDA implementation here
CrossMergeTritonF = CrossMergeF  # csm_triton.py:456


def cross_scan_fn(x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False):
    """csm_triton.py:501-507.  ``force_torch`` is accepted for signature parity; there is one implementation."""
    del force_torch
    return CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)


def cross_merge_fn(y, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False):
    """csm_triton.py:511-517."""
    del force_torch
    return CrossMergeF.apply(y, in_channel_first, out_channel_first, one_by_one, scans)


def merge_norm_gate(ys: torch.Tensor, H: int, W: int, weight: torch.Tensor, bias: torch.Tensor, zact=None, eps=1e-5,
                    out_dtype=None):
    """Fused tail of the SS2D core: CrossMerge -> LayerNorm(d_inner) [-> * zact]  (VMamba.py:632-646, :364-372).

    ys (B, 4, C, L) scan outputs;  zact (B, H, W, C) or None;  returns (B, H, W, C)."""
    dev = _lib.require_cuda(ys, weight, bias, zact)
    ys = ys.contiguous()
    B, K, C, L = ys.shape
    if K != 4 or L != H * W:
        raise RuntimeError("merge_norm_gate expects ys of shape (B, 4, C, H*W)")
    out_dtype = out_dtype or (zact.dtype if zact is not None else ys.dtype)
    out = torch.empty((B, H, W, C), dtype=out_dtype, device=dev)
    if zact is not None:
        zact = zact.contiguous()
        if zact.dtype != out_dtype or zact.numel() != out.numel():
            raise RuntimeError("zact must match the output dtype and shape (B, H, W, C)")
    if out.numel() == 0:
        return out
    lib = _lib.lib()
    nbytes = lib.xp_merge_norm_gate_workspace_bytes(B, C, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xp_merge_norm_gate(_lib.ptr(ys), _lib.ptr(weight.float().contiguous()),
                                          _lib.ptr(bias.float().contiguous()), _lib.ptr(zact), _lib.ptr(out), B, C, H, W,
                                          _lib.dtype_code(ys), _lib.dtype_code(out), float(eps), _lib.ptr(ws), nbytes,
                                          _lib.stream_ptr(dev)))
    _lib.count_launches(3)
    return out


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5, out_dtype=None):
    """LayerNorm over the last dimension of a channel-last tensor (the no-residual case of xp_add_layer_norm)."""
    return add_layer_norm(x, None, weight, bias, eps, y_dtype=out_dtype or x.dtype, want_sum=False)[1]


def add_layer_norm(x: torch.Tensor, res, weight, bias, eps: float = 1e-5, y_dtype=None, pre_bias=None, sum_dtype=None,
                   want_sum: bool = True, want_y: bool = True):
    """One pass of ``s = x [+ res] [+ pre_bias]; y = LayerNorm(s)`` over channel-last rows (xp_add_layer_norm).

    Returns ``(s or None, y or None)``.  x / res: (..., C) contiguous; weight / bias / pre_bias: (C)."""
    dev = _lib.require_cuda(x, res, weight, bias, pre_bias)
    x = x.contiguous()
    C = x.shape[-1]
    if res is not None:
        res = res.contiguous()
        if res.shape != x.shape:
            raise RuntimeError("add_layer_norm: x and res must have the same shape")
    s = torch.empty(x.shape, dtype=sum_dtype or (res.dtype if res is not None else x.dtype), device=dev) if want_sum else None
    y = torch.empty(x.shape, dtype=y_dtype or x.dtype, device=dev) if want_y else None
    if x.numel():
        f32 = lambda t: None if t is None else t.float().contiguous()
        w, b, pb = f32(weight), f32(bias), f32(pre_bias)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_add_layer_norm(
                _lib.ptr(x), _lib.ptr(res), _lib.ptr(pb), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), _lib.ptr(s), x.numel() // C, C,
                _lib.dtype_code(x), _lib.dtype_code(res) if res is not None else 0, _lib.dtype_code(y) if y is not None else 0,
                _lib.dtype_code(s) if s is not None else 0, float(eps), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return s, y


def patch_embed_stem(img: torch.Tensor, weight: torch.Tensor, bias, ln_weight, ln_bias, eps: float = 1e-5, out_dtype=None,
                     gelu: bool = True) -> torch.Tensor:
    """img (B, Cin, H, W) fp32, Cin in {1, 3}; weight (C1, Cin, 3, 3) -> (B, ceil(H/2), ceil(W/2), C1) channel-last:
    Conv2d(k3, s2, p1) + bias + LayerNorm(C1) [+ GELU] in one kernel (xp_patch_embed_stem)."""
    dev = _lib.require_cuda(img, weight, bias, ln_weight, ln_bias)
    img = img.float().contiguous()
    B, Cin, H, W = img.shape
    C1 = weight.shape[0]
    if tuple(weight.shape) != (C1, Cin, 3, 3):
        raise RuntimeError("patch_embed_stem: weight must have shape (C1, Cin, 3, 3)")
    out = torch.empty((B, (H + 1) // 2, (W + 1) // 2, C1), dtype=out_dtype or torch.float32, device=dev)
    if out.numel():
        f32 = lambda t: None if t is None else t.float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_patch_embed_stem(_lib.ptr(img), _lib.ptr(f32(weight)), _lib.ptr(f32(bias)),
                                                      _lib.ptr(f32(ln_weight)), _lib.ptr(f32(ln_bias)), _lib.ptr(out), B, Cin, H, W,
                                                      C1, float(eps), _lib.dtype_code(out), int(bool(gelu)), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


def linear_act(x: torch.Tensor, weight: torch.Tensor, bias, gelu: bool = False) -> torch.Tensor:
    """act(x W^T + b) on tcgen05 tensor cores (xp_linear_act).  x (..., K) fp16 | bf16 contiguous, weight (N, K) in the
    same dtype, bias (N) fp32 or None -> (..., N); gelu=True applies the exact (erf) GELU in the GEMM epilogue."""
    dev = _lib.require_cuda(x, weight, bias)
    if x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype:
        raise RuntimeError("linear_act: x and weight must both be fp16 or bf16")
    x = x.contiguous()
    weight = weight.contiguous()
    N, K = weight.shape
    if x.shape[-1] != K:
        raise RuntimeError("linear_act: shape mismatch")
    out = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=dev)
    if out.numel():
        b = None if bias is None else bias.float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_linear_act(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(b), _lib.ptr(out), x.numel() // K, N, K,
                                                _lib.dtype_code(x), int(bool(gelu)), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


def linear_res_ln(x: torch.Tensor, weight: torch.Tensor, bias, res: torch.Tensor, ln_weight, ln_bias, eps: float = 1e-5,
                  want_sum: bool = True):
    """``s = res + x W^T + b ; y = LayerNorm(s)`` in one tcgen05 GEMM (xp_linear_res_ln): the end of a VSSBlock branch
    (out_proj / fc2 -> residual add -> next norm, VMamba.py:664,110-128,1222-1234).  x (..., K) fp16 | bf16, weight (N, K) same
    dtype, res (..., N) fp32; returns (s fp32 or None, y in x.dtype).  N in {96, 192, 384}."""
    dev = _lib.require_cuda(x, weight, bias, res, ln_weight, ln_bias)
    if x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype or res.dtype != torch.float32:
        raise RuntimeError("linear_res_ln: x / weight must both be fp16 or bf16 and the residual fp32")
    x, weight, res = x.contiguous(), weight.contiguous(), res.contiguous()
    N, K = weight.shape
    if x.shape[-1] != K or res.shape[-1] != N or res.numel() // N != x.numel() // K:
        raise RuntimeError("linear_res_ln: shape mismatch")
    s = torch.empty_like(res) if want_sum else None
    y = torch.empty(res.shape, dtype=x.dtype, device=dev)
    if y.numel():
        f32 = lambda t: None if t is None else t.float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_linear_res_ln(_lib.ptr(x), _lib.ptr(weight), _lib.ptr(f32(bias)), _lib.ptr(res),
                                                   _lib.ptr(f32(ln_weight)), _lib.ptr(f32(ln_bias)), _lib.ptr(s), _lib.ptr(y),
                                                   x.numel() // K, N, K, _lib.dtype_code(x), float(eps), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return s, y


def mlp_res_ln(x: torch.Tensor, w1: torch.Tensor, b1, w2: torch.Tensor, b2, res: torch.Tensor, ln_weight, ln_bias,
               eps: float = 1e-5, want_sum: bool = True):
    """The Mlp branch of a VSSBlock in one tcgen05 kernel (xp_mlp_res_ln): ``s = res + fc2(GELU(fc1(x))) ; y = LayerNorm(s)``
    (Mlp.forward VMamba.py:110-128 + the block's residual add and the next norm, :1229-1234); the 4C-wide hidden activation
    never reaches HBM.  x (..., C) fp16 | bf16, w1 (4C, C), w2 (C, 4C) same dtype, res (..., C) fp32, C in {96, 192};
    returns (s fp32 or None, y in x.dtype); with ``ln_weight=None`` y is s rounded to x.dtype (no LayerNorm)."""
    dev = _lib.require_cuda(x, w1, b1, w2, b2, res, ln_weight, ln_bias)
    if x.dtype not in (torch.float16, torch.bfloat16) or w1.dtype != x.dtype or w2.dtype != x.dtype or res.dtype != torch.float32:
        raise RuntimeError("mlp_res_ln: x / w1 / w2 must all be fp16 or bf16 and the residual fp32")
    x, w1, w2, res = x.contiguous(), w1.contiguous(), w2.contiguous(), res.contiguous()
    C = x.shape[-1]
    if tuple(w1.shape) != (4 * C, C) or tuple(w2.shape) != (C, 4 * C) or res.shape[-1] != C or res.numel() != x.numel():
        raise RuntimeError("mlp_res_ln: shape mismatch (hidden width must be 4C)")
    s = torch.empty_like(res) if want_sum else None
    y = torch.empty(res.shape, dtype=x.dtype, device=dev)
    if y.numel():
        f32 = lambda t: None if t is None else t.float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_mlp_res_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(f32(b1)), _lib.ptr(w2), _lib.ptr(f32(b2)),
                                                _lib.ptr(res), _lib.ptr(f32(ln_weight)), _lib.ptr(f32(ln_bias)), _lib.ptr(s),
                                                _lib.ptr(y), x.numel() // C, C, _lib.dtype_code(x), float(eps),
                                                _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return s, y


def mlp_res_ln_supported(C: int, hidden: int) -> bool:
    return C in (96, 192) and hidden == 4 * C


def linear_res_ln_supported(K: int, N: int) -> bool:
    return N in (96, 192, 384) and K % 8 == 0


def linear_act_supported(K: int, N: int) -> bool:
    return K % 8 == 0 and N % 32 == 0 and N <= 8192
