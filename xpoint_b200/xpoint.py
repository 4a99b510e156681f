"""XPoint model (VMamba encoder + detector / descriptor heads) and the batched pair pipeline.

``XPoint`` mirrors xpoint/models/XPoint.py for the configuration the north star names -- VMamba encoder,
``multispectral: false``, descriptor head on, homography head off (it only works at 256x256):
same ``XPoint(config)`` constructor, ``takes_pair()``, ``forward(data)`` output structure (tuple of dicts with
``prob, logits, desc, encoder_output`` when ``takes_pair`` -- XPoint.py:181-214,283-323) and the same state_dict
keys, so reference checkpoints load with ``load_state_dict``.

``PairPipeline`` is the evaluation tail (xpoint/utils/evaluation.py:229-301) without its per-sample Python loop
and device<->host copies: NMS + top-k + raster-order keypoints, bilinear descriptor sampling and mutual-NN
matching for the whole batch on the GPU.
"""
from __future__ import annotations

import contextlib
import copy
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import postprocess as pp
from .vmamba import PRESETS, VSSM

_VSSM_KEYS = dict(PATCH_SIZE="patch_size", IN_CHANS="in_chans", DEPTHS="depths", EMBED_DIM="dims", SSM_D_STATE="ssm_d_state",
                  SSM_RATIO="ssm_ratio", SSM_DT_RANK="ssm_dt_rank", SSM_ACT_LAYER="ssm_act_layer", SSM_CONV="ssm_conv",
                  SSM_CONV_BIAS="ssm_conv_bias", SSM_DROP_RATE="ssm_drop_rate", SSM_INIT="ssm_init",
                  SSM_FORWARDTYPE="forward_type", MLP_RATIO="mlp_ratio", MLP_ACT_LAYER="mlp_act_layer",
                  MLP_DROP_RATE="mlp_drop_rate", PATCH_NORM="patch_norm", NORM_LAYER="norm_layer", DOWNSAMPLE="downsample_version",
                  PATCHEMBED="patchembed_version", POSEMBED="posembed", GMLP="gmlp")
# defaults of vmamba_src/MYCONFIG.py:72-97
_VSSM_DEFAULTS = dict(patch_size=4, in_chans=3, depths=[2, 2, 9, 2], dims=96, ssm_d_state=16, ssm_ratio=2.0, ssm_dt_rank="auto",
                      ssm_act_layer="silu", ssm_conv=3, ssm_conv_bias=True, ssm_drop_rate=0.0, ssm_init="v0", forward_type="v2",
                      mlp_ratio=4.0, mlp_act_layer="gelu", mlp_drop_rate=0.0, patch_norm=True, norm_layer="ln",
                      downsample_version="v2", patchembed_version="v2", posembed=False, gmlp=False)


def _dict_update(d, u):  # xpoint/utils/utils.py:73-89
    for k, v in u.items():
        if isinstance(v, dict):
            d[k] = _dict_update(d.get(k, {}) if isinstance(d.get(k, {}), dict) else {}, v)
        else:
            d[k] = v
    return d


def vssm_kwargs_from_config(model_parameters: dict) -> dict:
    """Translate the YAML tree the reference feeds to MYCONFIG.get_config (params.yaml:107-129) into VSSM kwargs."""
    kw = dict(_VSSM_DEFAULTS)
    model = (model_parameters or {}).get("MODEL", {})
    for k, v in model.get("VSSM", {}).items():
        if k in _VSSM_KEYS:
            kw[_VSSM_KEYS[k]] = v
    kw["drop_path_rate"] = model.get("DROP_PATH_RATE", 0.1)
    if kw["ssm_dt_rank"] != "auto":
        kw["ssm_dt_rank"] = int(kw["ssm_dt_rank"])
    return kw


class XPoint(nn.Module):
    default_config = {
        "multispectral": False, "descriptor_head": True, "intepolation_mode": "bilinear", "descriptor_size": 256,
        "normalize_descriptors": True, "final_batchnorm": True, "reflection_pad": True, "bn_first": False,
        "double_convolution": True, "channel_version": 0, "verbose": False, "mixed_precision": False,
        "force_return_logits": False, "takes_pair": False,
        "homography_regression_head": {"check": False, "type": "RegNet"},
        "use_attention": {"check": True, "type": "VMamba", "height": 256, "width": 256, "preset": "E",
                          "model_parameters": None},
    }

    def __init__(self, config=None):
        super().__init__()
        self.config = _dict_update(copy.deepcopy(self.default_config), config or {})
        c = self.config
        if c["multispectral"]:
            raise NotImplementedError("multispectral (two encoders) is not part of the accelerated path")
        if c["homography_regression_head"]["check"]:
            raise NotImplementedError("the RegNet homography head only works at 256x256 and is disabled at the benchmark "
                                      "sizes (configs/cipdp.yaml:48); it is not built")
        ua = c["use_attention"]
        if not ua["check"] or ua["type"] != "VMamba":
            raise NotImplementedError("only the VMamba encoder is built (use_attention.type == 'VMamba')")
        if ua.get("model_parameters"):
            kw = vssm_kwargs_from_config(ua["model_parameters"])
        else:
            kw = dict(PRESETS[ua.get("preset", "E")])
        self.encoder = VSSM(**kw)
        embed = kw["dims"] if isinstance(kw["dims"], int) else kw["dims"][0]
        self.n_channels = [1, 64, 64, 128, embed // 2]          # XPoint.py:96,436
        self.head_channels = 256
        self.encoder_downsample_ratio = 8
        pad = nn.ReflectionPad2d if c["reflection_pad"] else nn.ZeroPad2d
        self.detector_head_last_dim = self.encoder_downsample_ratio ** 2 + 1

        def nonlin(n):
            return (nn.BatchNorm2d(n), nn.ReLU(True)) if c["bn_first"] else (nn.ReLU(True), nn.BatchNorm2d(n))

        det = [pad(1), nn.Conv2d(self.n_channels[4], self.head_channels, 3), *nonlin(self.head_channels),
               nn.Conv2d(self.head_channels, self.detector_head_last_dim, 1)]
        if c["final_batchnorm"]:
            det.append(nn.BatchNorm2d(self.detector_head_last_dim))
        self.detector_head_convolutions = nn.Sequential(*det)
        if c["descriptor_head"]:
            desc = [pad(1), nn.Conv2d(self.n_channels[4], self.head_channels, 3), *nonlin(self.head_channels),
                    nn.Conv2d(self.head_channels, c["descriptor_size"], 1)]
            if c["final_batchnorm"]:
                desc.append(nn.BatchNorm2d(c["descriptor_size"]))
            self.descriptor_head_convolutions = nn.Sequential(*desc)

    def takes_pair(self):
        return self.config["takes_pair"]

    def get_encoder_downsample_ratio(self):
        return self.encoder_downsample_ratio

    def set_force_return_logits(self, value):
        if not isinstance(value, bool):
            raise ValueError("set_force_return_logits: The input value needs to be a bool")
        self.config["force_return_logits"] = value

    # XPoint.py:348-360
    def detector_head(self, x):
        logits = self.detector_head_convolutions(x)
        if self.training or self.config["force_return_logits"]:
            return None, logits.to(torch.float)
        return pp.detector_post(logits, self.encoder_downsample_ratio), None

    # XPoint.py:362-371
    def descriptor_head(self, x):
        x = self.descriptor_head_convolutions(x)
        if self.config["normalize_descriptors"]:
            return pp.normalize_descriptors(x)
        return x.to(torch.float)

    # -- inference-time form of the two heads (XPoint.py:112-138, 348-371) -------------------------------------------
    # pad -> conv3x3 -> ReLU -> BN -> conv1x1 -> BN, twice on the same input.  In eval mode both BatchNorms are affine
    # maps, so they fold into the 1x1 convolution:  W2' = diag(s2) W2 diag(s1),  b2' = s2 * (W2 t1 + b2) + t2  with
    # s = gamma / sqrt(var + eps), t = beta - mean * s.  The two 3x3 convolutions share their input and run as ONE
    # channels-last convolution with 512 outputs; the 1x1 convolutions are GEMMs with the bias in the epilogue.
    def _heads_foldable(self):
        c = self.config
        if self.training or c["bn_first"] or not c["final_batchnorm"] or not c["descriptor_head"] or not c["normalize_descriptors"]:
            return False
        d = self.detector_head_convolutions
        return len(d) == 6 and isinstance(d[3], nn.BatchNorm2d) and isinstance(d[5], nn.BatchNorm2d)

    def _folded_heads(self):
        mods = list(self.detector_head_convolutions) + list(self.descriptor_head_convolutions)
        params = [t for m in mods for t in list(m.parameters()) + list(m.buffers())]
        key = tuple((t.data_ptr(), t._version) for t in params)
        cache = getattr(self, "_head_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]

        def fold(seq):
            conv1, bn1, conv2, bn2 = seq[1], seq[3], seq[4], seq[5]
            s1 = bn1.weight.float() / torch.sqrt(bn1.running_var.float() + bn1.eps)
            t1 = bn1.bias.float() - bn1.running_mean.float() * s1
            s2 = bn2.weight.float() / torch.sqrt(bn2.running_var.float() + bn2.eps)
            t2 = bn2.bias.float() - bn2.running_mean.float() * s2
            W2 = conv2.weight.float().flatten(1)                                  # (out, 256)
            b2 = conv2.bias.float() if conv2.bias is not None else torch.zeros_like(s2)
            return conv1.weight.float(), conv1.bias.float(), (s2[:, None] * W2 * s1[None, :]).contiguous(), s2 * (W2 @ t1 + b2) + t2

        with torch.no_grad():
            w1d, b1d, w2d, b2d = fold(self.detector_head_convolutions)
            w1s, b1s, w2s, b2s = fold(self.descriptor_head_convolutions)
            # the 65-wide detector GEMM is padded to 72 columns (16-byte rows for 16-bit outputs, aligned cuBLAS kernels)
            npad = (-w2d.shape[0]) % 8
            w2d_p = torch.cat([w2d, w2d.new_zeros(npad, w2d.shape[1])]) if npad else w2d
            b2d_p = torch.cat([b2d, b2d.new_zeros(npad)]) if npad else b2d
            folded = dict(w1=torch.cat([w1d, w1s]).contiguous(memory_format=torch.channels_last), b1=torch.cat([b1d, b1s]),
                          w2_det=w2d_p.t().contiguous(), b2_det=b2d_p, w2_desc=w2s.t().contiguous(), b2_desc=b2s,
                          n1=w1d.shape[0], n_det=w2d.shape[0])
        self._head_cache = (key, folded)
        return folded

    def _heads_fused(self, x, padded=None, want_desc=True):
        """x (B, C, Hc, Wc) encoder output; ``padded``: its ReflectionPad2d(1) already in the compute dtype and
        channels-last (xp_encoder_tail), else it is formed here.  Returns prob, logits, desc, desc_cl."""
        f32w = self._folded_heads()
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        f = f32w.get(cdt)
        if f is None:                            # weights cast once per compute dtype
            f = {k: (v.to(cdt) if torch.is_tensor(v) else v) for k, v in f32w.items() if not isinstance(k, torch.dtype)}
            f["w1"] = f["w1"].contiguous(memory_format=torch.channels_last)
            f32w[cdt] = f
        Bn, _, Hc, Wc = x.shape
        if padded is not None and padded.dtype == cdt:
            xp = padded
        else:
            xp = self.detector_head_convolutions[0](x).to(dtype=cdt, memory_format=torch.channels_last)   # pad once for both heads
        y = None
        if not getattr(self, "_no_cudnn_relu", False):
            try:                                 # cuDNN's fused conv + bias + ReLU
                y = torch.cudnn_convolution_relu(xp, f["w1"], f["b1"], (1, 1), (0, 0), (1, 1), 1)
            except RuntimeError:
                self._no_cudnn_relu = True
        if y is None:
            y = torch.relu_(nn.functional.conv2d(xp, f["w1"], f["b1"]))
        y = y.permute(0, 2, 3, 1).reshape(-1, y.shape[1])                                            # (rows, 512) channel-last
        n1 = f["n1"]
        logits = torch.addmm(f["b2_det"], y[:, :n1], f["w2_det"])                                    # (rows, 72), 65 used
        draw = torch.addmm(f["b2_desc"], y[:, n1:], f["w2_desc"])                                    # (rows, 256)
        # the channel-last GEMM outputs are consumed where they lie (no (B, C, Hc, Wc) permute copies)
        if self.config["force_return_logits"]:
            prob = None
            lg = logits[:, :f["n_det"]].reshape(Bn, Hc, Wc, -1).permute(0, 3, 1, 2).contiguous().to(torch.float)
        else:
            prob, lg = pp.detector_post_rows(logits, Bn, Hc, Wc, self.encoder_downsample_ratio), None
        desc, desc_cl = pp.normalize_descriptor_rows(draw, Bn, Hc, Wc, want_channel_first=want_desc)
        return prob, lg, desc, desc_cl

    def forward_impl(self, data, want_desc=True):  # XPoint.py:283-323
        out = {"prob": None, "logits": None}
        if self._heads_foldable():
            # residual add -> depth_to_space -> clone -> reflection pad -> compute dtype in ONE pass (xp_encoder_tail)
            xf, pend = self.encoder.forward_features(data["image"])
            cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else torch.float32
            if pp.encoder_tail_supported(xf):
                # xp_encoder_tail's padded output is ReflectionPad2d(1); with reflection_pad=False the module's own
                # ZeroPad2d is applied inside _heads_fused instead (XPoint.py:101-104)
                encoder_output, padded = pp.encoder_tail(xf, pend, 4, pad_dtype=cdt if self.config["reflection_pad"] else None)
            else:
                x = self.encoder.depth_to_space((xf if pend is None else xf + pend).permute(0, 3, 1, 2), 4)
                encoder_output, padded = x, None
            # desc_cl: the same unit descriptors as (B, Hc, Wc, 256) -- the layout the sampler gathers 1 KiB rows from
            out["prob"], out["logits"], out["desc"], out["desc_cl"] = self._heads_fused(encoder_output, padded, want_desc)
            out["encoder_output"] = encoder_output
            return out
        x = self.encoder(data["image"])
        encoder_output = x.clone().detach()
        out["prob"], out["logits"] = self.detector_head(x)
        if self.config["descriptor_head"]:
            out["desc"] = self.descriptor_head(x)
        out["encoder_output"] = encoder_output
        return out

    def forward(self, data):  # XPoint.py:181-214
        ctx = torch.autocast("cuda", dtype=torch.float16) if self.config["mixed_precision"] else contextlib.nullcontext()
        with ctx:
            if not self.takes_pair():
                return self.forward_impl(data)
            pred_optical = self.forward_impl(data["optical"])
            pred_thermal = self.forward_impl(data["thermal"])
            return pred_optical, pred_thermal, None

    def forward_pair_batched(self, optical: torch.Tensor, thermal: torch.Tensor, want_desc=True, split=True):
        """Same weights serve both spectra (multispectral=False), so the two encoder passes of XPoint.forward
        (XPoint.py:187-188) run as one 2B batch.  Returns (pred_optical, pred_thermal) dicts, or with ``split=False`` the
        single dict of the 2B batch (optical first); ``want_desc=False`` skips the channel-first descriptor map when
        only the channel-last copy is consumed."""
        Bn = optical.shape[0]
        ctx = torch.autocast("cuda", dtype=torch.float16) if self.config["mixed_precision"] else contextlib.nullcontext()
        with ctx:
            out = self.forward_impl({"image": torch.cat([optical, thermal], 0)}, want_desc=want_desc)
        if not split:
            return out
        first = {k: (v[:Bn] if v is not None else None) for k, v in out.items()}
        second = {k: (v[Bn:] if v is not None else None) for k, v in out.items()}
        return first, second


class PairResult(NamedTuple):
    kp_optical: torch.Tensor      # (B, k, 2) int32 (y, x)
    kp_thermal: torch.Tensor
    n_optical: torch.Tensor       # (B,) int32
    n_thermal: torch.Tensor
    desc_optical: torch.Tensor    # (B, k, 256) fp32, zero rows past n
    desc_thermal: torch.Tensor
    match_idx: torch.Tensor       # (B, k) int32, thermal keypoint index matched to optical keypoint i, or -1
    match_dist: torch.Tensor      # (B, k) fp32
    n_matches: torch.Tensor       # (B,) int32
    H: Optional[torch.Tensor] = None           # (B, 3, 3) float64 optical -> thermal, with estimate_homography=True
    inliers: Optional[torch.Tensor] = None     # (B, k) bool per optical keypoint
    n_inliers: Optional[torch.Tensor] = None   # (B,) int32, -1 = no estimate


class PairPipeline:
    """net(data) -> box_nms -> nonzero -> interpolate_descriptors x2 -> get_matches, batched on the GPU
    (xpoint/utils/evaluation.py:229-301 with configs/cipdp.yaml:52-55: nms 8, detection_threshold 0.015)."""

    def __init__(self, net: Optional[XPoint], nms=8, detection_threshold=0.015, iou=0.1, keep_top_k=4096,
                 use_tensor_cores=True, estimate_homography=False, reprojection_threshold=3.0, ransac_iters=2048):
        if keep_top_k < 0:
            raise ValueError("keep_top_k must be >= 0 (0 = no cap, as utils.box_nms)")
        self.net = net
        self.nms, self.thr, self.iou, self.topk = nms, detection_threshold, iou, keep_top_k
        self.use_tensor_cores = use_tensor_cores
        # the step after matching in the evaluation (evaluation.py:359-378), off by default because
        # BASELINE's step ends at the matches
        self.estimate_homography, self.reproj_thr, self.ransac_iters = estimate_homography, reprojection_threshold, ransac_iters

    def capacity(self, H: int, W: int) -> int:
        """Keypoint rows per image.  keep_top_k = 0 is the reference's "no cap" (utils.py:179, configs/cipdp.yaml:55):
        the greedy NMS itself bounds the survivors -- the footprint of a kept pixel contains every pixel within
        Chebyshev distance 4 for box side 8, in general distance d = max |dx| with (s - d)^2 above the IoU
        bound, so kept pixels own disjoint (d+1)x(d+1) blocks."""
        if self.topk > 0:
            return self.topk
        s, t = float(self.nms), float(self.iou)
        bound = 2.0 * s * s * t / (1.0 + t)
        d = 0
        while (s - (d + 1)) > 0 and (s - (d + 1)) ** 2 > bound:
            d += 1
        blk = d + 1
        return ((H + blk - 1) // blk + 1) * ((W + blk - 1) // blk + 1)

    def tail(self, prob_o, prob_t, desc_o, desc_t, channel_last=False, valid_mask_o=None, valid_mask_t=None) -> PairResult:
        """prob (B,1,H,W) fp32; desc (B,256,Hc,Wc) fp32 [or (B,Hc,Wc,256) with channel_last] -> keypoints, descriptors
        and mutual matches.  valid_mask_* (B,1,H,W): the evaluation's `prob * data[...]['valid_mask']`
        (evaluation.py:250-251), applied before NMS."""
        if valid_mask_o is not None:
            prob_o = prob_o * valid_mask_o.reshape(prob_o.shape).to(prob_o.dtype)
        if valid_mask_t is not None:
            prob_t = prob_t * valid_mask_t.reshape(prob_t.shape).to(prob_t.dtype)
        return self.tail_batched(torch.cat([prob_o, prob_t], 0), torch.cat([desc_o, desc_t], 0), channel_last)

    def tail_batched(self, prob, desc, channel_last=False, valid_mask=None) -> PairResult:
        """The same on the 2B batch the encoder produced (optical images first, then thermal): no concatenation.
        valid_mask: optional (2B,1,H,W) / (2B,H,W) multiplier applied to prob before NMS (evaluation.py:250-251)."""
        B, H, W = prob.shape[0] // 2, prob.shape[-2], prob.shape[-1]
        prob = prob.reshape(2 * B, H, W)
        if valid_mask is not None:
            prob = prob * valid_mask.reshape(2 * B, H, W).to(prob.dtype)
        cap = self.capacity(H, W)
        kps = pp.nms_keypoints(prob, self.nms, self.thr, self.iou, self.topk, kp_threshold=self.thr, capacity=cap,
                               want_map=False)
        count = torch.clamp(kps.count, max=cap)
        d = pp.sample_descriptors(kps.keypoints, count, desc, H, W, channel_last)
        m = pp.mnn_match(d[:B], d[B:], count[:B], count[B:], use_tensor_cores=self.use_tensor_cores)
        hom = (None, None, None)
        if self.estimate_homography:
            hom = pp.estimate_homography(kps.keypoints[:B], kps.keypoints[B:], m.match_idx, H, W, count[:B],
                                         iters=self.ransac_iters, reproj_threshold=self.reproj_thr)
        return PairResult(kps.keypoints[:B], kps.keypoints[B:], count[:B], count[B:], d[:B], d[B:], m.match_idx,
                          m.match_dist, m.count, *hom)

    def capture(self, optical: torch.Tensor, thermal: torch.Tensor, warmup: int = 3) -> "GraphedPairPipeline":
        """Record one whole step (encoder, heads, NMS, sampling, matching: ~180 launches of this library, cuBLAS and cuDNN)
        into a CUDA graph for the shapes of ``optical`` / ``thermal``.  The step has no host synchronisation and no
        data-dependent launch shape, so a replay is the same work without the per-launch host cost and the gaps it leaves
        on the device."""
        return GraphedPairPipeline(self, optical, thermal, warmup)

    @torch.no_grad()
    def __call__(self, optical: torch.Tensor, thermal: torch.Tensor, valid_mask_optical: Optional[torch.Tensor] = None,
                 valid_mask_thermal: Optional[torch.Tensor] = None) -> PairResult:
        out = self.net.forward_pair_batched(optical, thermal, want_desc=False, split=False)
        mask = None
        if valid_mask_optical is not None or valid_mask_thermal is not None:
            ones = torch.ones_like(optical, dtype=torch.float32)
            mask = torch.cat([ones if valid_mask_optical is None else valid_mask_optical.reshape(optical.shape).float(),
                              ones if valid_mask_thermal is None else valid_mask_thermal.reshape(thermal.shape).float()], 0)
        if out.get("desc_cl") is not None:
            return self.tail_batched(out["prob"], out["desc_cl"], channel_last=True, valid_mask=mask)
        return self.tail_batched(out["prob"], out["desc"], valid_mask=mask)


class GraphedPairPipeline:
    """A PairPipeline step replayed from a CUDA graph.  ``__call__(optical, thermal)`` copies the images into the graph's
    static input buffers (device-to-device, or host-to-device for pinned host tensors), replays, and returns the PairResult
    whose tensors are the graph's static outputs -- valid until the next call."""

    def __init__(self, pipe: PairPipeline, optical: torch.Tensor, thermal: torch.Tensor, warmup: int = 3):
        if not optical.is_cuda:
            raise RuntimeError("GraphedPairPipeline.capture needs CUDA example inputs (their shapes are baked into the graph)")
        self.pipe = pipe
        dev = optical.device
        with torch.cuda.device(dev):
            self.static_o, self.static_t = optical.clone(), thermal.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):           # warm-up on a side stream: lazy initialisation must not be captured
                for _ in range(max(warmup, 1)):
                    pipe(self.static_o, self.static_t)
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            # an explicit capture stream ON THIS DEVICE: torch.cuda.graph's default capture stream is a class-level singleton
            # created on whichever device was current at its first use, so a capture on a second GPU would not be recorded
            with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(device=dev)):
                self.result = pipe(self.static_o, self.static_t)

    def load(self, optical: torch.Tensor, thermal: torch.Tensor, non_blocking: bool = True) -> None:
        self.static_o.copy_(optical, non_blocking=non_blocking)
        self.static_t.copy_(thermal, non_blocking=non_blocking)

    def replay(self) -> PairResult:
        self.graph.replay()
        return self.result

    def __call__(self, optical: torch.Tensor, thermal: torch.Tensor) -> PairResult:
        self.load(optical, thermal)
        return self.replay()
