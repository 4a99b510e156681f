"""VMamba encoder blocks around the B200 SS2D hot path.

Host-side mirror of the parts of xpoint/models/vmamba_src/VMamba.py that XPoint instantiates:
``SS2D`` (VMamba.py:1107-1149; forward types ``v0`` :305-374 and the ``v2`` family ``v01..v05/v2/v3`` with the
``_noz/_nozact/_oact/_no32`` postfixes :380-664), ``VSSBlock`` (:1153-1240) and ``VSSM`` (:1243-1525), with
identical constructor keywords and state_dict keys/shapes so reference checkpoints load.

What is ours: the SS2D core -- CrossScan -> selective scan -> CrossMerge -> out_norm [-> gate] -- runs on the
hand-written sm_100a kernels of libxpoint_b200.so.  What stays a library call ("OUT OF SCOPE ...
library call; stays a library call"): the dense projections / convolutions / LayerNorms around it
(torch -> cuBLAS / cuDNN).  There is no CPU path: calling a module on CPU tensors raises.
"""
from __future__ import annotations

import os

import math
from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ss2d as _ss2d
from .cross_scan import (add_layer_norm, cross_scan_fn, layer_norm, linear_act, linear_act_supported, linear_res_ln,
                         mlp_res_ln, mlp_res_ln_supported,
                         linear_res_ln_supported, merge_norm_gate, patch_embed_stem)
from .selective_scan import scan_forward, selective_scan_fn


class Permute(nn.Module):
    def __init__(self, *args):
        super().__init__()
        self.args = args

    def forward(self, x):
        return x.permute(*self.args)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm with the same parameters / state_dict keys, forward on the xp_layer_norm kernel.

    ``for_matmul``: the result only feeds a Linear, so under autocast it is written directly in the autocast
    dtype (torch computes the norm in fp32 and lets the Linear cast it: same values, one pass less)."""

    def __init__(self, normalized_shape, eps=1e-5, for_matmul=False):
        super().__init__(normalized_shape, eps=eps)
        self.for_matmul = for_matmul

    def forward(self, x):
        out_dtype = None
        if torch.is_autocast_enabled():
            out_dtype = torch.get_autocast_dtype('cuda') if self.for_matmul else torch.float32
        return layer_norm(x, self.weight, self.bias, self.eps, out_dtype)


class Mlp(nn.Module):  # VMamba.py:110-128 (channel-last only; XPoint never builds channel-first VSSMs)
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, channels_first=False):
        super().__init__()
        assert not channels_first
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.hidden(x))

    def fusable(self) -> bool:
        """fc1 -> exact GELU -> fc2 with hidden = 4 C and C in {96, 192}: the whole branch fits xp_mlp_res_ln."""
        return (isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none"
                and self.fc2.out_features == self.fc1.in_features
                and mlp_res_ln_supported(self.fc1.in_features, self.fc1.out_features) and not getattr(self, "disable_fused", False))

    def hidden(self, x):
        """act(fc1(x)): everything before fc2 (the block wrapper may run fc2 fused with the residual add and the next norm)."""
        # 16-bit activations (autocast): fc1 + bias + exact GELU in ONE tcgen05 GEMM, the hidden tensor is written once
        if (x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and isinstance(self.act, nn.GELU)
                and getattr(self.act, "approximate", "none") == "none"
                and linear_act_supported(self.fc1.in_features, self.fc1.out_features) and not getattr(self, "disable_fused", False)):
            w = self.fc1.weight
            key = (w.data_ptr(), w._version, x.dtype)
            cache = getattr(self, "_w1_cache", None)
            if cache is None or cache[0] != key:
                cache = (key, w.detach().to(x.dtype).contiguous())
                self._w1_cache = cache
            return linear_act(x, cache[1], self.fc1.bias, gelu=True)
        return self.act(self.fc1(x))


class PatchMerging2D(nn.Module):  # VMamba.py:60-98, channel-last
    def __init__(self, dim, out_dim=-1, norm_layer=nn.LayerNorm, channel_first=False):
        super().__init__()
        assert not channel_first
        self.reduction = nn.Linear(4 * dim, (2 * dim) if out_dim < 0 else out_dim, bias=False)
        self.norm = LayerNorm(4 * dim, for_matmul=True)

    def forward(self, x):
        H, W, _ = x.shape[-3:]
        if (W % 2 != 0) or (H % 2 != 0):
            x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        x = torch.cat([x[..., 0::2, 0::2, :], x[..., 1::2, 0::2, :], x[..., 0::2, 1::2, :], x[..., 1::2, 1::2, :]], -1)
        return self.reduction(self.norm(x))


def _init_dt_A_D(d_state, dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1,
                 dt_init_floor=1e-4, k_group=4):
    """Same distributions as mamba_init.init_dt_A_D (VMamba.py:165-232)."""
    ws, bs = [], []
    std = dt_rank ** -0.5 * dt_scale
    for _ in range(k_group):
        w = torch.empty(d_inner, dt_rank)
        if dt_init == "constant":
            nn.init.constant_(w, std)
        else:
            nn.init.uniform_(w, -std, std)
        dt = torch.exp(torch.rand(d_inner) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
        ws.append(w)
        bs.append(dt + torch.log(-torch.expm1(-dt)))  # softplus^-1
    A_logs = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32)).view(1, -1).repeat(k_group * d_inner, 1)
    Ds = torch.ones(k_group * d_inner)
    return nn.Parameter(A_logs), nn.Parameter(Ds), nn.Parameter(torch.stack(ws)), nn.Parameter(torch.stack(bs))


class SS2D(nn.Module):
    """Constructor keywords and parameter names of VMamba.SS2D (VMamba.py:1107-1149)."""

    def __init__(self, d_model=96, d_state=16, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU, d_conv=3, conv_bias=True,
                 dropout=0.0, bias=False, dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 initialize="v0", forward_type="v2", channel_first=False, **kwargs):
        super().__init__()
        if channel_first:
            raise NotImplementedError("channel_first SS2D is not used by XPoint and is not built")
        if dropout > 0.0:
            raise NotImplementedError("inference-only: dropout must be 0")
        self.k_group = 4
        self.d_model, self.d_state = int(d_model), int(d_state)
        self.d_inner = int(ssm_ratio * d_model)
        self.dt_rank = int(math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank)
        ft = forward_type
        if ft in ("v0", "v0seq"):
            # SS2Dv0 (VMamba.py:236-303): z gate with SiLU, conv bias on, fp32 scan inputs, mamba-style output dtype
            self.family = "v0"
            self.disable_z = self.disable_z_act = self.oact = False
            self.force_fp32, self.oflex = True, False
            d_conv, conv_bias, bias, act_layer = 3, True, False, nn.SiLU
            initialize = "v0"
        else:
            self.family = "v2"

            def cut(tag, value):
                hit = value.endswith(tag)
                return hit, (value[: -len(tag)] if hit else value)

            disable_force32, ft = cut("_no32", ft)
            self.oact, ft = cut("_oact", ft)
            self.disable_z, ft = cut("_noz", ft)
            self.disable_z_act, ft = cut("_nozact", ft)
            for tag in ("_onnone", "_ondwconv3", "_oncnorm", "_onsoftmax", "_onsigmoid"):
                if ft.endswith(tag):
                    raise NotImplementedError(f"out_norm variant {tag} is not used by XPoint and is not built")
            table = dict(v01=(not disable_force32, False), v02=(not disable_force32, False), v03=(not disable_force32, True),
                         v04=(False, True), v05=(False, True), v2=(not disable_force32, False), v3=(False, True))
            if ft not in table:
                raise NotImplementedError(f"SS2D forward_type {forward_type!r} is outside the hot path (supported: v0, "
                                          f"{', '.join(table)} with _noz/_nozact/_oact/_no32)")
            self.force_fp32, self.oflex = table[ft]
        self.forward_type = forward_type
        self.with_dconv = d_conv > 1
        d_proj = self.d_inner if self.disable_z else 2 * self.d_inner
        self.in_proj = nn.Linear(self.d_model, d_proj, bias=bias)
        self.act = act_layer()
        if self.with_dconv:
            self.conv2d = nn.Conv2d(self.d_inner, self.d_inner, groups=self.d_inner, bias=conv_bias, kernel_size=d_conv,
                                    padding=(d_conv - 1) // 2)
        self.x_proj_weight = nn.Parameter(torch.stack([
            nn.Linear(self.d_inner, self.dt_rank + 2 * self.d_state, bias=False).weight.detach() for _ in range(self.k_group)]))
        self.out_act = nn.GELU() if self.oact else nn.Identity()
        self.out_norm = nn.LayerNorm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, self.d_model, bias=bias)
        if initialize == "v0":
            self.A_logs, self.Ds, self.dt_projs_weight, self.dt_projs_bias = _init_dt_A_D(
                self.d_state, self.dt_rank, self.d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor, self.k_group)
        elif initialize == "v1":
            self.Ds = nn.Parameter(torch.ones(self.k_group * self.d_inner))
            self.A_logs = nn.Parameter(torch.randn(self.k_group * self.d_inner, self.d_state))
            self.dt_projs_weight = nn.Parameter(0.1 * torch.randn(self.k_group, self.d_inner, self.dt_rank))
            self.dt_projs_bias = nn.Parameter(0.1 * torch.randn(self.k_group, self.d_inner))
        else:
            self.Ds = nn.Parameter(torch.ones(self.k_group * self.d_inner))
            self.A_logs = nn.Parameter(torch.zeros(self.k_group * self.d_inner, self.d_state))
            self.dt_projs_weight = nn.Parameter(0.1 * torch.rand(self.k_group, self.d_inner, self.dt_rank))
            self.dt_projs_bias = nn.Parameter(0.1 * torch.rand(self.k_group, self.d_inner))

    # -------------------------------------------------------------------------------------------------------
    def forward_core(self, x: torch.Tensor, zact=None) -> torch.Tensor:
        """x (B, d_inner, H, W) -> (B, H, W, d_inner): the hot path of VMamba.py:493-646 / :314-372.

        CrossScan, selective scan and CrossMerge + out_norm (+ gate) are libxpoint_b200 kernels; the two small
        projections x_proj / dt_proj stay grouped 1x1 convolutions (cuDNN/cuBLAS), as in the reference's
        ``no_einsum`` path (VMamba.py:605-608)."""
        B, D, H, W = x.shape
        K, N, R, L = self.k_group, self.d_state, self.dt_rank, H * W
        xs = cross_scan_fn(x, True, True, False, 0)                                        # (B, 4, D, L)
        # (K, R+2N, D) @ (B, K, D, L): strided-batched GEMMs straight on the scan layout (cuBLAS); the grouped
        # conv1d of the reference's no_einsum path routes through cuDNN layout conversions that cost more than the math
        x_dbl = torch.matmul(self.x_proj_weight.unsqueeze(0), xs)                          # (B, K, R+2N, L)
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        dts = torch.matmul(self.dt_projs_weight.unsqueeze(0), dts).view(B, -1, L)          # (B, K*D, L)
        us = xs.view(B, -1, L)
        if dts.dtype != us.dtype:
            dts = dts.to(us.dtype)
        if Bs.dtype != us.dtype:
            Bs, Cs = Bs.to(us.dtype), Cs.to(us.dtype)
        if self.force_fp32:
            us, dts, Bs, Cs = us.float(), dts.float(), Bs.float(), Cs.float()
        As = -self.A_logs.float().exp()
        ys = selective_scan_fn(us, dts, As, Bs, Cs, self.Ds.float(), self.dt_projs_bias.float().view(-1), True,
                               True if self.family == "v0" else self.oflex, None)       # (B, 4*D, L)
        return merge_norm_gate(ys.view(B, K, D, L), H, W, self.out_norm.weight, self.out_norm.bias, zact,
                               self.out_norm.eps, out_dtype=x.dtype)

    # -------------------------------------------------------------------------------------------------------
    def _fused_weights(self, batch: int, dtype: torch.dtype):
        """Projection / scan parameters re-ordered to ss2d.FUSED_ORDER, cast once and cached (inference: the
        parameters are static; the cache is keyed on their versions so in-place updates invalidate it)."""
        params = (self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds, self.out_norm.weight,
                  self.out_norm.bias) + ((self.conv2d.weight,) if self.with_dconv else ())
        key = (batch, dtype, self.x_proj_weight.device) + tuple((q.data_ptr(), q._version) for q in params)
        cache = getattr(self, "_fw_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]
        K, D, N, R = self.k_group, self.d_inner, self.d_state, self.dt_rank
        order = list(_ss2d.FUSED_ORDER)
        with torch.no_grad():
            wx = self.x_proj_weight.detach()[order].reshape(1, 2, 2 * (R + 2 * N), D).to(dtype)
            wdt = self.dt_projs_weight.detach()[order].reshape(1, 2, 2, D, R).to(dtype)
            w = dict(
                wx=wx.expand(batch, -1, -1, -1).contiguous(),                     # (B, 2, 2(R+2N), D)
                wdt=wdt.expand(batch, -1, -1, -1, -1).contiguous(),               # (B, 2, 2, D, R)
                wdt32=self.dt_projs_weight.detach().float()[order].contiguous(),  # (4, D, R) for xp_ss2d_dt_proj
                wdt_scan=self.dt_projs_weight.detach()[order].reshape(K * D, R).to(dtype).contiguous(),   # fused dt_proj
                A=(-self.A_logs.detach().float().exp()).view(K, D, N)[order].reshape(K * D, N).contiguous(),
                Ds=self.Ds.detach().float().view(K, D)[order].reshape(-1).contiguous(),
                dt_bias=self.dt_projs_bias.detach().float()[order].reshape(-1).contiguous(),
                norm_w=self.out_norm.weight.detach().float().contiguous(),
                norm_b=self.out_norm.bias.detach().float().contiguous(),
            )
            if self.with_dconv:
                w["conv_w"] = self.conv2d.weight.detach().float().contiguous()
                w["conv_b"] = None if self.conv2d.bias is None else self.conv2d.bias.detach().float().contiguous()
        self._fw_cache = (key, w)
        return w

    def forward_core_fused(self, xx: torch.Tensor, H: int, W: int, zact=None, out_dtype=None) -> torch.Tensor:
        """xx (B, 2, d_inner, L) = [x ; x^T] -> (B, H, W, d_inner).  Same math as ``forward_core`` without the
        CrossScan / CrossMerge copies: both projections run directly on the two layouts, the scan kernel routes the
        four directions, and one pass merges + normalises (+ gates)."""
        B, _, D, L = xx.shape
        K, N, R = self.k_group, self.d_state, self.dt_rank
        w = self._fused_weights(B, xx.dtype)
        x_dbl = torch.matmul(w["wx"], xx)                                          # (B, 2, 2(R+2N), L)
        x_dbl = x_dbl.view(B, K, R + 2 * N, L)
        if (N <= 2 and R <= 16 and L % 8 == 0 and xx.dtype != torch.float32 and getattr(self, "fuse_dt_proj", FUSE_DT_PROJ)):
            # the scan forms delta = W_dt x dts_r itself, the (B, K*D, L) delta never reaches HBM
            ys, _ = scan_forward(xx.view(B, 2 * D, L), x_dbl[:, :, :R], w["A"], x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:],
                                 w["Ds"], None, w["dt_bias"], True, True, u_group_div=2,
                                 reverse_group_mask=_ss2d.REVERSE_MASK, dt_weight=w["wdt_scan"])
            return _ss2d.ss2d_merge_norm(ys.view(B, K, D, L), H, W, w["norm_w"], w["norm_b"], zact, self.out_norm.eps,
                                         out_dtype=out_dtype or xx.dtype)
        if _ss2d.dt_proj_supported(R, L, x_dbl.dtype, B * K, D):
            dts = _ss2d.ss2d_dt_proj(x_dbl[:, :, :R], w["wdt32"])                  # (B, 4, D, L), store-bound kernel
        else:
            dts = torch.matmul(w["wdt"], x_dbl.view(B, 2, 2, R + 2 * N, L)[:, :, :, :R])   # (B, 2, 2, D, L)
        if getattr(self, "use_core", USE_CORE) and _ss2d.core_channels(D, N, H, W, xx.dtype) > 0:
            # the four directions meet in shared memory: ONE merged fp32 plane instead of four (xp_ss2d_core)
            y = _ss2d.ss2d_core(xx, dts.view(B, K, D, L), w["A"], x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], w["Ds"],
                                w["dt_bias"], H, W, True)
            return _ss2d.ss2d_plane_norm(y, w["norm_w"], w["norm_b"], zact, self.out_norm.eps, out_dtype=out_dtype or xx.dtype)
        ys, _ = scan_forward(xx.view(B, 2 * D, L), dts.view(B, K * D, L), w["A"], x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:],
                             w["Ds"], None, w["dt_bias"], True, True, u_group_div=2,
                             reverse_group_mask=_ss2d.REVERSE_MASK)               # (B, 4*D, L) fp32, natural order
        return _ss2d.ss2d_merge_norm(ys.view(B, K, D, L), H, W, w["norm_w"], w["norm_b"], zact, self.out_norm.eps,
                                     out_dtype=out_dtype or xx.dtype)

    def _use_fused(self, H: int, W: int, dtype: torch.dtype) -> bool:
        # force_fp32 (v0, v01-v03: VMamba.py:341,625 cast xs / dts / Bs / Cs to fp32 before the scan) needs no cast here: the
        # kernels widen 16-bit operands to fp32 in registers and keep fp32 state and accumulation, which is the same arithmetic
        # on the same values -- the four (B, K*D, L) fp32 copies the reference materialises are simply not made
        return (_ss2d.fused_supported(H, W, self.d_inner, self.d_state)
                and (self.family == "v0" or self.oflex or dtype == torch.float32) and not getattr(self, "disable_fused", False))

    def forward(self, x: torch.Tensor, skip_out_proj: bool = False) -> torch.Tensor:
        """x (B, H, W, d_model) -> (B, H, W, d_model)   (forwardv0 VMamba.py:305-374, forwardv2 :648-664).
        ``skip_out_proj``: return the out_proj INPUT (B, H, W, d_inner); the block wrapper then runs out_proj fused with the
        residual add and the next LayerNorm (xp_linear_res_ln).

        in_proj / out_proj stay cuBLAS GEMMs.  The depth-wise 3x3 convolution + SiLU run on xp_ss2d_dwconv_pack, which
        reads the channel-last in_proj output directly and emits the channel-first layouts the scan needs (no permute
        copy, no cuDNN call: cuDNN's fp16 depth-wise kernel also returned wrong values for 768 channels at 16x20 on
        B200, scripts/debug_unfused.py)."""
        x = self.in_proj(x)
        Bn, H, W, _ = x.shape
        D = self.d_inner
        zact = None
        if not self.disable_z:
            zact = x[..., D:]
            zact = (zact if self.disable_z_act else self.act(zact)).contiguous()
        fused = self._use_fused(H, W, x.dtype)
        if self.with_dconv and self.conv2d.kernel_size == (3, 3) and isinstance(self.act, nn.SiLU):
            w = self._fused_weights(Bn, x.dtype)
            xx = _ss2d.ss2d_dwconv_pack(x, D, w["conv_w"], w["conv_b"], silu=True)          # (B, 2, D, L) = [x ; x^T]
            xc = None if fused else xx[:, 0].view(Bn, D, H, W)
        else:
            xc = x[..., :D].permute(0, 3, 1, 2).contiguous()
            if self.with_dconv:
                xc = self.conv2d(xc)
            xc = self.act(xc)
            xx = _ss2d.ss2d_pack(xc) if fused else None
        core = (lambda z: self.forward_core_fused(xx, H, W, zact=z)) if fused else (lambda z: self.forward_core(xc, zact=z))
        if isinstance(self.out_act, nn.Identity):
            y = core(zact)
        else:
            y = self.out_act(core(None))
            if zact is not None:
                y = y * zact
        return y if skip_out_proj else self.out_proj(y)


# SURVEY 8f row f1 (fused dt_proj, xp_scan_args.dt_weight) is built, parity-tested and measured, but it is NOT the default:
# the N=1 scan kernel is bound by instruction issue / latency as much as by HBM, so dropping the delta stream (-25 % bytes)
# buys 0.3 ms per stage-0 block against dt_proj + scan while the kernel's bytes/s falls.
# XP_FUSE_DT_PROJ=1 or `module.fuse_dt_proj = True` turns it on.
FUSE_DT_PROJ = os.environ.get("XP_FUSE_DT_PROJ", "0") not in ("", "0")
# XP_SS2D_CORE=1 (or `module.use_core = True`): the fused core xp_ss2d_core (CrossMerge inside the scan: one merged fp32 plane
# in HBM instead of four, 2.3x less DRAM traffic for scan + merge) replaces xp_selective_scan_fwd + xp_ss2d_merge_norm.  Built,
# parity-tested and measured, but NOT the default: on B200 it is bound by instruction issue (12 warps / SM, 160 registers, one
# (b, channel) plane set per SM) and runs 2.4 ms per stage-0 block against 1.5 ms for the op-level scan, which the cheaper
# out_norm pass (0.7 vs 1.0 ms) does not win back.
USE_CORE = os.environ.get("XP_SS2D_CORE", "0") not in ("", "0")
# XP_NO_LINEAR_LN=1: keep out_proj / fc2 on cuBLAS followed by xp_add_layer_norm (A/B measurements)
FUSE_LINEAR_LN_OFF = os.environ.get("XP_NO_LINEAR_LN", "0") not in ("", "0")
FUSE_MLP_OFF = os.environ.get("XP_NO_FUSED_MLP", "0") not in ("", "0")      # A/B switch for xp_mlp_res_ln (profiles/)


class VSSBlock(nn.Module):  # VMamba.py:1153-1240
    def __init__(self, hidden_dim=0, drop_path=0.0, norm_layer=nn.LayerNorm, channel_first=False, ssm_d_state=16,
                 ssm_ratio=2.0, ssm_dt_rank="auto", ssm_act_layer=nn.SiLU, ssm_conv=3, ssm_conv_bias=True, ssm_drop_rate=0.0,
                 ssm_init="v0", forward_type="v2", mlp_ratio=4.0, mlp_act_layer=nn.GELU, mlp_drop_rate=0.0, gmlp=False,
                 use_checkpoint=False, post_norm=False, **kwargs):
        super().__init__()
        if gmlp or post_norm:
            raise NotImplementedError("gmlp / post_norm are not used by XPoint and are not built")
        self.ssm_branch = ssm_ratio > 0
        self.mlp_branch = mlp_ratio > 0
        if self.ssm_branch:
            self.norm = LayerNorm(hidden_dim, for_matmul=True)
            self.op = SS2D(d_model=hidden_dim, d_state=ssm_d_state, ssm_ratio=ssm_ratio, dt_rank=ssm_dt_rank,
                           act_layer=ssm_act_layer, d_conv=ssm_conv, conv_bias=ssm_conv_bias, dropout=ssm_drop_rate,
                           initialize=ssm_init, forward_type=forward_type, channel_first=channel_first)
        if self.mlp_branch:
            self.norm2 = LayerNorm(hidden_dim, for_matmul=True)
            self.mlp = Mlp(hidden_dim, int(hidden_dim * mlp_ratio), act_layer=mlp_act_layer, drop=mlp_drop_rate)

    def forward(self, x):  # drop_path is the identity in eval mode (inference tier)
        if self.ssm_branch:
            x = x + self.op(self.norm(x))
        if self.mlp_branch:
            x = x + self.mlp(self.norm2(x))
        return x


class VSSM(nn.Module):
    """VMamba.VSSM as XPoint uses it (VMamba.py:1243-1525): no classifier, output (B, dims[-1]/16, H/8, W/8)."""

    def __init__(self, patch_size=4, in_chans=3, num_classes=1000, depths=(2, 2, 9, 2), dims=(96, 192, 384, 768),
                 ssm_d_state=16, ssm_ratio=2.0, ssm_dt_rank="auto", ssm_act_layer="silu", ssm_conv=3, ssm_conv_bias=True,
                 ssm_drop_rate=0.0, ssm_init="v0", forward_type="v2", mlp_ratio=4.0, mlp_act_layer="gelu", mlp_drop_rate=0.0,
                 gmlp=False, drop_path_rate=0.1, patch_norm=True, norm_layer="LN", downsample_version="v2",
                 patchembed_version="v1", use_checkpoint=False, posembed=False, imgsize=224, **kwargs):
        super().__init__()
        if norm_layer.lower() != "ln":
            raise NotImplementedError("only channel-last LayerNorm VSSMs are built (XPoint uses norm_layer='ln')")
        if posembed:
            raise NotImplementedError("posembed is not used by XPoint and is not built")
        self.in_chans = in_chans
        depths = list(depths)
        self.num_layers = len(depths)
        if isinstance(dims, int):
            dims = [int(dims * 2 ** i) for i in range(self.num_layers)]
        self.dims = list(dims)
        acts = dict(silu=nn.SiLU, gelu=nn.GELU, relu=nn.ReLU, sigmoid=nn.Sigmoid)
        ssm_act, mlp_act = acts[ssm_act_layer.lower()], acts[mlp_act_layer.lower()]
        LN = LayerNorm
        if patchembed_version == "v1":
            self.patch_embed = nn.Sequential(
                nn.Conv2d(in_chans, dims[0], kernel_size=patch_size, stride=patch_size, bias=True), Permute(0, 2, 3, 1),
                (LN(dims[0]) if patch_norm else nn.Identity()))
        else:
            s = patch_size // 2
            k = s + 1
            self.patch_embed = nn.Sequential(
                nn.Conv2d(in_chans, dims[0] // 2, kernel_size=k, stride=s, padding=1),
                (Permute(0, 2, 3, 1) if patch_norm else nn.Identity()),
                (LN(dims[0] // 2) if patch_norm else nn.Identity()),
                (Permute(0, 3, 1, 2) if patch_norm else nn.Identity()),
                nn.GELU(),
                nn.Conv2d(dims[0] // 2, dims[0], kernel_size=k, stride=s, padding=1), Permute(0, 2, 3, 1),
                (LN(dims[0]) if patch_norm else nn.Identity()))

        def downsample(dim, out_dim):
            if downsample_version == "v1":
                return PatchMerging2D(dim, out_dim, LN)
            if downsample_version == "v2":
                return nn.Sequential(Permute(0, 3, 1, 2), nn.Conv2d(dim, out_dim, kernel_size=2, stride=2),
                                     Permute(0, 2, 3, 1), LN(out_dim))
            if downsample_version == "v3":
                return nn.Sequential(Permute(0, 3, 1, 2), nn.Conv2d(dim, out_dim, kernel_size=3, stride=2, padding=1),
                                     Permute(0, 2, 3, 1), LN(out_dim))
            raise NotImplementedError(downsample_version)

        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            blocks = [VSSBlock(hidden_dim=self.dims[i], norm_layer=LN, ssm_d_state=ssm_d_state, ssm_ratio=ssm_ratio,
                               ssm_dt_rank=ssm_dt_rank, ssm_act_layer=ssm_act, ssm_conv=ssm_conv, ssm_conv_bias=ssm_conv_bias,
                               ssm_drop_rate=ssm_drop_rate, ssm_init=ssm_init, forward_type=forward_type, mlp_ratio=mlp_ratio,
                               mlp_act_layer=mlp_act, mlp_drop_rate=mlp_drop_rate, gmlp=gmlp) for _ in range(depths[i])]
            ds = downsample(self.dims[i], self.dims[i + 1]) if i < self.num_layers - 1 else nn.Identity()
            self.layers.append(nn.Sequential(OrderedDict(blocks=nn.Sequential(*blocks), downsample=ds)))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @staticmethod
    def depth_to_space(x, bs):  # VMamba.py:1500-1505
        N, C, H, W = x.size()
        x = x.view(N, bs, bs, C // (bs * bs), H, W).permute(0, 3, 4, 1, 5, 2).contiguous()
        return x.view(N, C // (bs * bs), H * bs, W * bs)

    # ------------------------------------------------------------------------------------------------------------
    # Forward.  Same math as VMamba.VSSM.forward (VMamba.py:1507-1525) with the glue between the library GEMMs /
    # convolutions fused: the residual stream is carried as (x, pending branch) so that every
    # `x = x + branch; n = norm(x)` pair is ONE xp_add_layer_norm pass; convolutions run channels-last so that the
    # permutes around them are views, and their bias is added inside the LayerNorm pass that follows.
    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _ln_dtype(for_matmul: bool):
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype("cuda") if for_matmul else torch.float32
        return None

    def _patch_embed(self, x):
        pe = self.patch_embed
        fast = (len(pe) == 8 and isinstance(pe[0], nn.Conv2d) and isinstance(pe[2], LayerNorm) and isinstance(pe[4], nn.GELU)
                and pe[0].kernel_size == (3, 3) and pe[0].stride == (2, 2) and pe[0].padding == (1, 1)
                and pe[0].out_channels <= 64 and pe[0].out_channels % 4 == 0 and x.dtype == torch.float32)
        if not fast:
            if self.in_chans == 3 and x.shape[1] == 1:
                x = torch.cat((x, x, x), dim=1)
            return pe(x)
        conv1, ln1, conv2, ln2 = pe[0], pe[2], pe[5], pe[7]
        w1 = conv1.weight
        if self.in_chans == 3 and x.shape[1] == 1:
            w1 = w1.sum(dim=1, keepdim=True)      # conv over three identical channels == conv with the summed kernel
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        y = patch_embed_stem(x, w1, conv1.bias, ln1.weight, ln1.bias, ln1.eps, out_dtype=cdt, gelu=True)   # (B, H/2, W/2, C1)
        y = F.conv2d(y.permute(0, 3, 1, 2), conv2.weight, None, conv2.stride, conv2.padding)              # channels-last in/out
        y = y.permute(0, 2, 3, 1)
        _, y = add_layer_norm(y, None, ln2.weight, ln2.bias, ln2.eps, y_dtype=self._ln_dtype(False) or y.dtype,
                              pre_bias=conv2.bias, want_sum=False)
        return y

    @staticmethod
    def _w16(lin: nn.Linear, dtype):
        """The Linear's weight in the autocast dtype, cast once (inference: static parameters, keyed on their version)."""
        w = lin.weight
        key = (w.data_ptr(), w._version, dtype)
        cache = getattr(lin, "_xp_w16", None)
        if cache is None or cache[0] != key:
            cache = (key, w.detach().to(dtype).contiguous())
            lin._xp_w16 = cache
        return cache[1]

    @staticmethod
    def _run_blocks(blocks, x, keep_mlp=False):
        """Returns (x, pending, deferred) with the block output = x + pending (pending may be None); ``deferred`` is None unless
        ``keep_mlp`` and the stage ends with a fusable Mlp branch: then (n, Mlp) with pending = Mlp(n) left to the caller
        (the downsample fuses it, VSSM._downsample).

        Under 16-bit autocast the projection that closes a branch (out_proj / fc2) is DEFERRED: it runs fused with the residual
        add and the LayerNorm that opens the next branch (xp_linear_res_ln, one tcgen05 GEMM) instead of cuBLAS + xp_add_layer_norm;
        a whole Mlp branch with C in {96, 192} is deferred as one unit (xp_mlp_res_ln: fc1 + GELU + fc2 + residual + next norm,
        the 4C-wide hidden activation never written).  The last branch of a stage has no next norm in this stage and is
        materialised by the plain modules."""
        pend = None
        defer = None            # (a16, Linear | Mlp): pending = module(a16), not yet computed
        cdt = VSSM._ln_dtype(True)
        fuse_ok = cdt in (torch.float16, torch.bfloat16) and x.dtype == torch.float32 and not FUSE_LINEAR_LN_OFF

        def can_defer(lin):
            return fuse_ok and linear_res_ln_supported(lin.in_features, lin.out_features)

        def open_branch(norm, x, pend, defer):
            """x [+ pending] -> (new x, LayerNorm(new x) for the branch's first GEMM)."""
            if defer is not None:
                a16, lin = defer
                if isinstance(lin, Mlp):
                    return mlp_res_ln(a16, VSSM._w16(lin.fc1, a16.dtype), lin.fc1.bias, VSSM._w16(lin.fc2, a16.dtype), lin.fc2.bias,
                                      x, norm.weight, norm.bias, norm.eps)
                return linear_res_ln(a16, VSSM._w16(lin, a16.dtype), lin.bias, x, norm.weight, norm.bias, norm.eps)
            if pend is None:
                return x, norm(x)
            return add_layer_norm(pend, x, norm.weight, norm.bias, norm.eps, y_dtype=cdt or x.dtype)

        for blk in blocks:
            if not isinstance(blk, VSSBlock):
                if defer is not None:
                    pend, defer = defer[1](defer[0]), None
                if pend is not None:
                    x, pend = x + pend, None
                x = blk(x)
                continue
            if blk.ssm_branch:
                x, n = open_branch(blk.norm, x, pend, defer)
                pend = defer = None
                if can_defer(blk.op.out_proj) and n.dtype == cdt:
                    defer = (blk.op(n, skip_out_proj=True), blk.op.out_proj)
                else:
                    pend = blk.op(n)
            if blk.mlp_branch:
                x, n = open_branch(blk.norm2, x, pend, defer)
                pend = defer = None
                if fuse_ok and not FUSE_MLP_OFF and n.dtype == cdt and blk.mlp.fusable():
                    defer = (n, blk.mlp)
                elif can_defer(blk.mlp.fc2) and n.dtype == cdt:
                    defer = (blk.mlp.hidden(n), blk.mlp.fc2)
                else:
                    pend = blk.mlp(n)
        if defer is not None:
            if keep_mlp and isinstance(defer[1], Mlp):
                return x, None, defer
            pend = defer[1](defer[0])
        return x, pend, None

    def _downsample(self, ds, x, pend, defer=None):
        fast = (isinstance(ds, nn.Sequential) and len(ds) == 4 and isinstance(ds[1], nn.Conv2d) and isinstance(ds[3], LayerNorm))
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        if defer is not None and not (fast and defer[0].dtype == cdt):
            pend, defer = defer[1](defer[0]), None
        if not fast:
            if pend is not None:
                x = x + pend
            return ds(x)
        conv, ln = ds[1], ds[3]
        if defer is not None:     # the stage's last Mlp branch + residual add, written once in the convolution's dtype
            n, mlp = defer
            _, x = mlp_res_ln(n, VSSM._w16(mlp.fc1, n.dtype), mlp.fc1.bias, VSSM._w16(mlp.fc2, n.dtype), mlp.fc2.bias, x,
                              None, None, want_sum=False)
        elif pend is not None:    # residual sum written once, already in the convolution's dtype and channels-last
            x, _ = add_layer_norm(pend, x, None, None, sum_dtype=cdt, want_y=False)
        elif x.dtype != cdt:
            x = x.to(cdt)
        y = F.conv2d(x.permute(0, 3, 1, 2), conv.weight, None, conv.stride, conv.padding).permute(0, 2, 3, 1)
        _, y = add_layer_norm(y, None, ln.weight, ln.bias, ln.eps, y_dtype=self._ln_dtype(False) or y.dtype,
                              pre_bias=conv.bias, want_sum=False)
        return y

    def forward_features(self, x):
        """Everything up to the last residual add: returns (x, pend) channel-last with the stage output = x + pend."""
        x = self._patch_embed(x)
        pend = None
        for i, layer in enumerate(self.layers):
            has_ds = not isinstance(layer.downsample, nn.Identity)
            x, pend, defer = self._run_blocks(layer.blocks, x, keep_mlp=has_ds)
            if has_ds:
                x, pend = self._downsample(layer.downsample, x, pend, defer), None
        return x, pend

    def forward(self, x):
        x, pend = self.forward_features(x)
        if pend is not None:
            x = x + pend
        return self.depth_to_space(x.permute(0, 3, 1, 2), 4)


# presets: SURVEY section 8, "V" = vanilla_vmamba_tiny (VMamba.py:1651-1662), "E" = shipped XPoint-EXP1
# (model_weights/XPoint-EXP1/params.yaml:107-129)
PRESETS = dict(
    V=dict(depths=[2, 2, 9, 2], dims=96, drop_path_rate=0.2, patch_size=4, in_chans=3, ssm_d_state=16, ssm_ratio=2.0,
           ssm_dt_rank="auto", ssm_act_layer="silu", ssm_conv=3, ssm_conv_bias=True, ssm_init="v0", forward_type="v0",
           mlp_ratio=0.0, downsample_version="v1", patchembed_version="v1", norm_layer="ln"),
    E=dict(depths=[2, 2, 2, 2], dims=96, drop_path_rate=0.2, patch_size=4, in_chans=3, ssm_d_state=1, ssm_ratio=1.0,
           ssm_dt_rank="auto", ssm_act_layer="silu", ssm_conv=3, ssm_conv_bias=False, ssm_init="v0",
           forward_type="v05_noz", mlp_ratio=4.0, downsample_version="v3", patchembed_version="v2", norm_layer="ln"),
)


def build_vssm(preset_or_kwargs) -> VSSM:
    kw = PRESETS[preset_or_kwargs] if isinstance(preset_or_kwargs, str) else dict(preset_or_kwargs)
    return VSSM(**kw)
