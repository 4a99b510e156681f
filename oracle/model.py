"""CPU restatement of the XPoint pair inference (encoder + heads + tail) -- TEST INFRASTRUCTURE ONLY.

Functional forward driven by a state_dict with the reference's key names.  Dense layers
(conv / linear / LayerNorm / BatchNorm / GELU / SiLU) call torch's CPU kernels exactly as the reference does on
CPU; the hot path (CrossScan, selective scan, CrossMerge, out_norm/gate, detector post, descriptor norm, NMS,
sampling, matching) goes through oracle/xp_oracle.c.  Follows:
    VSSM.forward            xpoint/models/vmamba_src/VMamba.py:1507-1525 (patch embed :1397-1420, downsample :60-98,:1432-1440)
    VSSBlock._forward       VMamba.py:1222-1234
    SS2D.forwardv0 / v2     VMamba.py:305-374 / :493-664
    XPoint.forward_impl     xpoint/models/XPoint.py:283-371
    evaluation tail         xpoint/utils/evaluation.py:229-301
Pinned against the real reference by tests/test_oracle_golden.py::test_xpoint_tiny_oracle (tests/golden/xpoint_tiny_*.npz).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O


def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def ss2d_forward(x, sd, p, forward_type):
    """x (B,H,W,C) -> (B,H,W,C).  `p` is the key prefix of the SS2D module ('...op')."""
    K, _, R = sd[p + ".dt_projs_weight"].shape
    N = sd[p + ".A_logs"].shape[1]
    v0 = forward_type in ("v0", "v0seq")
    noz = "_noz" in forward_type and "_nozact" not in forward_type
    x = F.linear(x, sd[p + ".in_proj.weight"], sd.get(p + ".in_proj.bias"))
    z = None
    if not noz:
        x, z = x.chunk(2, dim=-1)
        if "_nozact" not in forward_type:
            z = F.silu(z)
    x = x.permute(0, 3, 1, 2).contiguous()
    if p + ".conv2d.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv2d.weight"], sd.get(p + ".conv2d.bias"), padding=1, groups=x.shape[1])
    x = F.silu(x)
    Bt, Dn, H, W = x.shape
    L = H * W
    xs = torch.from_numpy(O.cross_scan(x.numpy()))
    if v0:
        x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, sd[p + ".x_proj_weight"])
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        dts = torch.einsum("bkrl,kdr->bkdl", dts, sd[p + ".dt_projs_weight"])
    else:  # no_einsum path, VMamba.py:605-608
        x_dbl = F.conv1d(xs.view(Bt, -1, L), sd[p + ".x_proj_weight"].reshape(-1, Dn, 1), groups=K)
        dts, Bs, Cs = torch.split(x_dbl.view(Bt, K, -1, L), [R, N, N], dim=2)
        dts = F.conv1d(dts.contiguous().view(Bt, -1, L), sd[p + ".dt_projs_weight"].reshape(K * Dn, -1, 1), groups=K)
    ys = O.selective_scan(xs.reshape(Bt, -1, L), dts.reshape(Bt, -1, L), -torch.exp(sd[p + ".A_logs"].float()),
                          Bs.contiguous(), Cs.contiguous(), sd[p + ".Ds"].float(), None,
                          sd[p + ".dt_projs_bias"].float().reshape(-1), True)
    y = O.cross_merge(ys.reshape(Bt, K, Dn, H, W))
    y = O.merge_norm_gate(y, sd[p + ".out_norm.weight"], sd[p + ".out_norm.bias"],
                          None if z is None else z.reshape(Bt, L, Dn).contiguous())
    y = torch.from_numpy(y).view(Bt, H, W, Dn)
    return F.linear(y, sd[p + ".out_proj.weight"], sd.get(p + ".out_proj.bias"))


def vssm_forward(img, sd, cfg, prefix="encoder"):
    """img (B,1|3,H,W) -> (B, dims[-1]/16, H/8, W/8)."""
    x = img
    if x.shape[1] == 1:
        x = torch.cat((x, x, x), dim=1)
    pe = prefix + ".patch_embed"
    if cfg["patchembed_version"] == "v1":
        x = F.conv2d(x, sd[pe + ".0.weight"], sd[pe + ".0.bias"], stride=cfg.get("patch_size", 4))
        x = _ln(x.permute(0, 2, 3, 1), sd, pe + ".2")
    else:
        x = F.conv2d(x, sd[pe + ".0.weight"], sd[pe + ".0.bias"], stride=2, padding=1)
        x = _ln(x.permute(0, 2, 3, 1), sd, pe + ".2").permute(0, 3, 1, 2)
        x = F.gelu(x)
        x = F.conv2d(x, sd[pe + ".5.weight"], sd[pe + ".5.bias"], stride=2, padding=1)
        x = _ln(x.permute(0, 2, 3, 1), sd, pe + ".7")
    depths = cfg["depths"]
    for i, depth in enumerate(depths):
        for j in range(depth):
            b = f"{prefix}.layers.{i}.blocks.{j}"
            x = x + ss2d_forward(_ln(x, sd, b + ".norm"), sd, b + ".op", cfg["forward_type"])
            if b + ".mlp.fc1.weight" in sd:
                h = F.linear(_ln(x, sd, b + ".norm2"), sd[b + ".mlp.fc1.weight"], sd[b + ".mlp.fc1.bias"])
                x = x + F.linear(F.gelu(h), sd[b + ".mlp.fc2.weight"], sd[b + ".mlp.fc2.bias"])
        if i < len(depths) - 1:
            d = f"{prefix}.layers.{i}.downsample"
            if cfg["downsample_version"] == "v1":
                H, W, _ = x.shape[-3:]
                if (W % 2 != 0) or (H % 2 != 0):
                    x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
                x = torch.cat([x[..., 0::2, 0::2, :], x[..., 1::2, 0::2, :], x[..., 0::2, 1::2, :], x[..., 1::2, 1::2, :]], -1)
                x = F.linear(_ln(x, sd, d + ".norm"), sd[d + ".reduction.weight"])
            else:
                k, pad = (3, 1) if cfg["downsample_version"] == "v3" else (2, 0)
                x = F.conv2d(x.permute(0, 3, 1, 2), sd[d + ".1.weight"], sd[d + ".1.bias"], stride=2, padding=pad)
                assert sd[d + ".1.weight"].shape[-1] == k
                x = _ln(x.permute(0, 2, 3, 1), sd, d + ".3")
    x = x.permute(0, 3, 1, 2)
    Nb, C, H, W = x.shape
    x = x.reshape(Nb, 4, 4, C // 16, H, W).permute(0, 3, 4, 1, 5, 2).contiguous()
    return x.view(Nb, C // 16, H * 4, W * 4)


def _head(x, sd, p):
    def bn(t, q):
        return F.batch_norm(t, sd[q + ".running_mean"], sd[q + ".running_var"], sd[q + ".weight"], sd[q + ".bias"], False, 0.0, 1e-5)
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    x = F.conv2d(x, sd[p + ".1.weight"], sd[p + ".1.bias"])
    x = bn(F.relu(x), p + ".3")
    x = F.conv2d(x, sd[p + ".4.weight"], sd[p + ".4.bias"])
    return bn(x, p + ".5")


@torch.no_grad()
def xpoint_forward(img, sd, cfg):
    """One image batch through encoder + heads.  Returns dict(prob, desc, encoder_output) of fp32 tensors."""
    x = vssm_forward(img.float(), sd, cfg)
    logits = _head(x, sd, "detector_head_convolutions")
    prob = torch.from_numpy(O.detector_post(logits.numpy(), 8))
    desc = torch.from_numpy(O.l2_normalize_channels(_head(x, sd, "descriptor_head_convolutions").numpy()))
    return {"prob": prob, "desc": desc, "encoder_output": x}


def pair_tail(prob_o, prob_t, desc_o, desc_t, nms=8, thr=0.015, iou=0.1, topk=4096):
    """evaluation.py:229-301 for one pair: returns (kp_o, kp_t, d_o, d_t, (q, t, dist))."""
    H, W = prob_o.shape[-2:]
    out = []
    for prob, desc in ((prob_o, desc_o), (prob_t, desc_t)):
        m = O.box_nms(np.asarray(prob).reshape(H, W), nms, thr, iou, topk)
        kp = O.extract_keypoints(m, thr)
        out.append((kp, O.interpolate_descriptors(kp, np.asarray(desc), H, W)))
    (kp_o, d_o), (kp_t, d_t) = out
    return kp_o, kp_t, d_o, d_t, O.mnn_match(d_o, d_t)


@torch.no_grad()
def pair_inference(optical, thermal, sd, cfg, **tail_kw):
    """Full reference-order pipeline for a batch of pairs on CPU.  Returns the list of per-pair tail results."""
    po = xpoint_forward(optical, sd, cfg)
    pt = xpoint_forward(thermal, sd, cfg)
    return [pair_tail(po["prob"][b, 0], pt["prob"][b, 0], po["desc"][b], pt["desc"][b], **tail_kw)
            for b in range(optical.shape[0])]
