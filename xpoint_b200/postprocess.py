"""Post-processing tail, call-compatible with the reference's utility functions.

  box_nms(prob, size, min_prob, iou=0.1, keep_top_k=0, on_cpu=False)      xpoint/utils/utils.py:148-192
  interpolate_descriptors(keypoints, descriptors_lowres, H, W)            xpoint/utils/utils.py:229-238
  get_matches(desc_1, desc_2, method='bfmatcher', knn_matches=False, **kw) xpoint/utils/matching.py:4-36
  NNMatcher(threshold).match(desc1, desc2)                                 xpoint/utils/matching.py:38-75
  detector_post(logits) / normalize_descriptors(x)                         XPoint.py:356-357 / :365-366

plus batched, sync-free variants used by the pair pipeline (``nms_keypoints``, ``sample_descriptors``,
``mnn_match``).  Everything runs on the GPU through libxpoint_b200.so; there is no CPU path (``on_cpu`` is accepted
and ignored, the result stays on the input's device exactly as the reference returns it there).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from . import _lib


# ------------------------------------------------------------------------------------------ heads
def detector_post(logits: torch.Tensor, r: int = 8) -> torch.Tensor:
    """Softmax2d -> drop dustbin -> PixelShuffle(r).  (B, r*r+1, Hc, Wc) -> (B, 1, r*Hc, r*Wc) fp32."""
    dev = _lib.require_cuda(logits)
    logits = logits.contiguous()
    B, Cn, Hc, Wc = logits.shape
    if Cn != r * r + 1:
        raise RuntimeError(f"detector_post: expected {r * r + 1} channels, got {Cn}")
    prob = torch.empty((B, 1, Hc * r, Wc * r), dtype=torch.float32, device=dev)
    if prob.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_detector_post(_lib.ptr(logits), _lib.ptr(prob), B, Hc, Wc, r, _lib.dtype_code(logits),
                                                   _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return prob


def normalize_descriptors(x: torch.Tensor, channel_last_copy: bool = False):
    """F.normalize(x, p=2, dim=1) for (B, C, H, W); optionally also returns a (B, H, W, C) copy."""
    dev = _lib.require_cuda(x)
    x = x.contiguous()
    B, C, H, W = x.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
    out_cl = torch.empty((B, H, W, C), dtype=torch.float32, device=dev) if channel_last_copy else None
    if out.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_l2_normalize(_lib.ptr(x), _lib.ptr(out), _lib.ptr(out_cl), B, C, H * W,
                                                  _lib.dtype_code(x), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return (out, out_cl) if channel_last_copy else out


def detector_post_rows(logits_rows: torch.Tensor, B: int, Hc: int, Wc: int, r: int = 8) -> torch.Tensor:
    """detector_post on channel-last rows: logits_rows (B*Hc*Wc, ld) with the r*r+1 logits of a cell first in its row
    (ld >= r*r+1; the model's 1x1-convolution GEMM writes 72-wide rows) -> (B, 1, r*Hc, r*Wc) fp32."""
    dev = _lib.require_cuda(logits_rows)
    if logits_rows.dim() != 2 or logits_rows.shape[0] != B * Hc * Wc or logits_rows.stride(1) != 1:
        raise RuntimeError("detector_post_rows: expected (B*Hc*Wc, ld) rows with unit column stride")
    if logits_rows.shape[1] < r * r + 1:
        raise RuntimeError(f"detector_post_rows: rows must hold at least {r * r + 1} logits")
    prob = torch.empty((B, 1, Hc * r, Wc * r), dtype=torch.float32, device=dev)
    if prob.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_detector_post_cl(_lib.ptr(logits_rows), _lib.ptr(prob), B, Hc, Wc, r, logits_rows.stride(0),
                                                      _lib.dtype_code(logits_rows), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return prob


def normalize_descriptor_rows(x_rows: torch.Tensor, B: int, Hc: int, Wc: int, want_channel_first: bool = True):
    """F.normalize over the channels of channel-last rows x_rows (B*Hc*Wc, C).  Returns (desc (B, C, Hc, Wc) fp32 or None,
    desc_cl (B, Hc, Wc, C) fp32)."""
    dev = _lib.require_cuda(x_rows)
    x_rows = x_rows.contiguous()
    C = x_rows.shape[1]
    if x_rows.dim() != 2 or x_rows.shape[0] != B * Hc * Wc:
        raise RuntimeError("normalize_descriptor_rows: expected (B*Hc*Wc, C) rows")
    out = torch.empty((B, C, Hc, Wc), dtype=torch.float32, device=dev) if want_channel_first else None
    out_cl = torch.empty((B, Hc, Wc, C), dtype=torch.float32, device=dev)
    if out_cl.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_l2_normalize_cl(_lib.ptr(x_rows), _lib.ptr(out), _lib.ptr(out_cl), B, C, Hc * Wc,
                                                     _lib.dtype_code(x_rows), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out, out_cl


def encoder_tail(x: torch.Tensor, pend, bs: int = 4, pad_dtype=None, want_encoder_output: bool = True):
    """x (B, H, W, C) fp32 channel-last residual stream [+ pend] -> depth_to_space(bs) (VMamba.py:1500-1505):
    (encoder_output (B, C/bs^2, H*bs, W*bs) fp32 or None, ReflectionPad2d(1) of it as a channels-last tensor of
    pad_dtype viewed (B, C/bs^2, H*bs+2, W*bs+2) or None)."""
    dev = _lib.require_cuda(x, pend)
    x = x.contiguous()
    pend = None if pend is None else pend.contiguous()
    B, H, W, Cn = x.shape
    CO = Cn // (bs * bs)
    enc = torch.empty((B, CO, H * bs, W * bs), dtype=torch.float32, device=dev) if want_encoder_output else None
    padded = None
    if pad_dtype is not None:
        padded = torch.empty((B, H * bs + 2, W * bs + 2, CO), dtype=pad_dtype, device=dev)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_encoder_tail(_lib.ptr(x), _lib.ptr(pend), _lib.ptr(enc), _lib.ptr(padded), B, H, W, CO, bs,
                                                  _lib.dtype_code(x), 0 if pend is None else _lib.dtype_code(pend),
                                                  0 if padded is None else _lib.dtype_code(padded), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return enc, (None if padded is None else padded.permute(0, 3, 1, 2))


def encoder_tail_supported(x: torch.Tensor, bs: int = 4) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[3] % (bs * bs) == 0 and \
        x.shape[3] // (bs * bs) in (8, 16, 32, 48, 64) and x.shape[1] * bs >= 3 and x.shape[2] * bs >= 3


# ------------------------------------------------------------------------------------------ NMS
class Keypoints(NamedTuple):
    prob_nms: Optional[torch.Tensor]   # (B, H, W) fp32 or None
    keypoints: torch.Tensor            # (B, capacity, 2) int32 (y, x), raster order, rows >= count undefined
    count: torch.Tensor                # (B,) int32


def nms_keypoints(prob: torch.Tensor, size, min_prob, iou=0.1, keep_top_k=0, kp_threshold=None, capacity=None,
                  want_map=True) -> Keypoints:
    """Batched NMS + top-k + raster-order keypoint compaction in one launch, no host sync.  prob (B,H,W)."""
    dev = _lib.require_cuda(prob)
    if prob.dtype != torch.float32:
        prob = prob.float()
    prob = prob.contiguous()
    B, H, W = prob.shape
    kp_threshold = float(min_prob if kp_threshold is None else kp_threshold)
    if capacity is None:
        capacity = keep_top_k if keep_top_k > 0 else H * W
    out = torch.empty_like(prob) if want_map else None
    kp = torch.empty((B, capacity, 2), dtype=torch.int32, device=dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    if B:
        lib = _lib.lib()
        nbytes = lib.xp_nms_workspace_bytes(B, H, W)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.xp_box_nms(_lib.ptr(prob), _lib.ptr(out), B, H, W, float(size), float(min_prob), float(iou),
                                      int(keep_top_k), kp_threshold, _lib.ptr(kp), _lib.ptr(cnt), capacity, _lib.ptr(ws),
                                      nbytes, _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return Keypoints(out, kp, cnt)


def box_nms(prob, size, min_prob, iou=0.1, keep_top_k=0, on_cpu=False):
    """Drop-in for utils.box_nms: prob (H,W) or (B,1,H,W) -> same shape, zeros + surviving scores."""
    del on_cpu
    if prob.dim() not in (2, 4):
        raise ValueError("The probability must be either 2D (H,W), or 4D (B, 1, H, W)")
    shape = prob.shape
    if prob.dim() == 4 and shape[1] != 1:
        raise ValueError("The probability must be either 2D (H,W), or 4D (B, 1, H, W)")
    res = nms_keypoints(prob.reshape(-1, shape[-2], shape[-1]), size, min_prob, iou, keep_top_k, capacity=1)
    return res.prob_nms.reshape(shape).to(prob.dtype)


# ------------------------------------------------------------------------------------------ sampling
def sample_descriptors(keypoints: torch.Tensor, count: Optional[torch.Tensor], desc: torch.Tensor, H: int, W: int,
                       channel_last: bool = False) -> torch.Tensor:
    """Batched bilinear sampling + L2 norm.  keypoints (B, n, 2) int32 (y, x); desc (B,C,Hc,Wc) or (B,Hc,Wc,C)."""
    dev = _lib.require_cuda(keypoints, count, desc)
    keypoints = keypoints.to(torch.int32).contiguous()
    desc = desc.float().contiguous()
    B, n = keypoints.shape[:2]
    if channel_last:
        _, Hc, Wc, C = desc.shape
    else:
        _, C, Hc, Wc = desc.shape
    if desc.shape[0] != B:
        raise RuntimeError("sample_descriptors: batch mismatch between keypoints and descriptors")
    out = torch.empty((B, n, C), dtype=torch.float32, device=dev)
    if out.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_sample_descriptors(_lib.ptr(keypoints), _lib.ptr(count), B, n, _lib.ptr(desc),
                                                        int(channel_last), C, Hc, Wc, H, W, _lib.ptr(out),
                                                        _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return out


def interpolate_descriptors(keypoints, descriptors_lowres, H, W):
    """Drop-in for utils.interpolate_descriptors: keypoints (N,2) (y,x); descriptors (C,Hc,Wc) -> (N,C)."""
    kp = keypoints.to(torch.int32).reshape(1, -1, 2)
    return sample_descriptors(kp, None, descriptors_lowres.unsqueeze(0), int(H), int(W))[0]


# ------------------------------------------------------------------------------------------ matching
class Matches(NamedTuple):
    match_idx: torch.Tensor    # (P, n1) int32: train index of the mutual match of query i, or -1
    match_dist: torch.Tensor   # (P, n1) fp32 L2 distance (0 where unmatched)
    count: torch.Tensor        # (P,) int32
    nn12: torch.Tensor         # (P, n1) int32
    nn21: torch.Tensor         # (P, n2) int32


def mnn_match(d1: torch.Tensor, d2: torch.Tensor, n1: Optional[torch.Tensor] = None, n2: Optional[torch.Tensor] = None,
              use_tensor_cores: bool = True) -> Matches:
    """Batched mutual-NN matching.  d1 (P, n1, C), d2 (P, n2, C) fp32; n1/n2 (P,) int32 valid counts or None."""
    dev = _lib.require_cuda(d1, d2, n1, n2)
    d1 = d1.float().contiguous()
    d2 = d2.float().contiguous()
    P, s1, C = d1.shape
    s2 = d2.shape[1]
    if d2.shape[0] != P or d2.shape[2] != C:
        raise RuntimeError("mnn_match: d1 and d2 must be (P, n, C) with equal P and C")
    nn12 = torch.full((P, s1), -1, dtype=torch.int32, device=dev)
    nn21 = torch.full((P, s2), -1, dtype=torch.int32, device=dev)
    midx = torch.full((P, s1), -1, dtype=torch.int32, device=dev)
    mdist = torch.zeros((P, s1), dtype=torch.float32, device=dev)
    cnt = torch.zeros((P,), dtype=torch.int32, device=dev)
    if P and s1 and s2:
        lib = _lib.lib()
        nbytes = lib.xp_match_workspace_bytes(P, s1, s2, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        n1 = None if n1 is None else n1.to(torch.int32).contiguous()
        n2 = None if n2 is None else n2.to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.xp_mnn_match(_lib.ptr(d1), _lib.ptr(d2), _lib.ptr(n1), _lib.ptr(n2), P, s1, s2, C,
                                        _lib.ptr(nn12), _lib.ptr(nn21), _lib.ptr(midx), _lib.ptr(mdist), _lib.ptr(cnt),
                                        int(bool(use_tensor_cores)), _lib.ptr(ws), nbytes, _lib.stream_ptr(dev)))
        _lib.count_launches(5)   # lower bound: 2 row-norm + arg-min kernels + mutual check (the tensor-core path adds the operand splits)
    return Matches(midx, mdist, cnt, nn12, nn21)


class DMatch(NamedTuple):
    """Field-compatible with cv2.DMatch as the reference consumes it (queryIdx, trainIdx, distance)."""
    queryIdx: int
    trainIdx: int
    distance: float
    imgIdx: int = 0


def _to_matches(res: Matches, max_dist: Optional[float] = None):
    idx = res.match_idx[0]
    q = torch.nonzero(idx >= 0).flatten()
    t = idx[q]
    d = res.match_dist[0][q]
    if max_dist is not None:
        keep = d < max_dist
        q, t, d = q[keep], t[keep], d[keep]
    return [DMatch(int(a), int(b), float(c)) for a, b, c in zip(q.tolist(), t.tolist(), d.tolist())]


def _as_cuda_desc(d):
    if not torch.is_tensor(d):
        d = torch.as_tensor(d)          # numpy in the reference's call sites (evaluation.py:287-291)
    if not d.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("xpoint_b200: no CUDA device; descriptor matching has no CPU path")
        d = d.cuda(non_blocking=True)
    return d.float()


class NNMatcher:
    """matching.py:38-75: mutual NN + distance threshold (0.7 default)."""

    def __init__(self, threshold=0.7, use_tensor_cores=True):
        if threshold < 0.0:
            raise ValueError("'threshold' should be non-negative")
        self.nn_thresh = threshold
        self.use_tensor_cores = use_tensor_cores

    def match(self, desc1, desc2):
        if desc1.shape[0] == 0 or desc2.shape[0] == 0:
            return []
        res = mnn_match(_as_cuda_desc(desc1)[None], _as_cuda_desc(desc2)[None], use_tensor_cores=self.use_tensor_cores)
        return _to_matches(res, self.nn_thresh)


def get_matches(desc_1, desc_2, method="bfmatcher", knn_matches=False, **kwargs):
    """Drop-in for matching.get_matches for the mutual-NN methods the XPoint evaluation uses:
    ``method='bfmatcher'`` with ``crossCheck=True`` and ``method='nnmatcher'``.  Returns DMatch-like tuples in
    increasing queryIdx.  Other methods (flann, knn ratio test, thresholdmatcher) are outside the hot path."""
    if knn_matches:
        raise NotImplementedError("knn_matches / ratio test is outside the accelerated hot path")
    if desc_1.shape[0] == 0 or desc_2.shape[0] == 0:
        return []
    use_tc = kwargs.pop("use_tensor_cores", True)
    if method == "bfmatcher":
        if not kwargs.pop("crossCheck", False):
            raise NotImplementedError("bfmatcher without crossCheck is outside the accelerated hot path")
        res = mnn_match(_as_cuda_desc(desc_1)[None], _as_cuda_desc(desc_2)[None], use_tensor_cores=use_tc)
        return _to_matches(res)
    if method == "nnmatcher":
        return NNMatcher(use_tensor_cores=use_tc, **kwargs).match(desc_1, desc_2)
    if method in ("flann", "thresholdmatcher"):
        raise NotImplementedError(f"method {method!r} is outside the accelerated hot path")
    raise ValueError("unknown matching method")


# ------------------------------------------------------------------------------------------ homography
class Homographies(NamedTuple):
    H: torch.Tensor            # (B, 3, 3) float64, maps image-1 (x, y) onto image-2; zeros where n_inliers == -1
    inliers: torch.Tensor      # (B, k) bool per keypoint row of image 1
    n_inliers: torch.Tensor    # (B,) int32, -1 = no estimate (fewer than 4 matches: the reference's H_est = None)


def estimate_homography(kp1: torch.Tensor, kp2: torch.Tensor, match_idx: torch.Tensor, height: int, width: int,
                        n1: Optional[torch.Tensor] = None, iters: int = 2048, reproj_threshold: float = 3.0,
                        lo_rounds: int = 2, seed: int = 0) -> Homographies:
    """Batched replacement of the per-pair ``cv2.findHomography(optical_pts, thermal_pts, USAC_MAGSAC, thr, ...)`` of the
    evaluation (xpoint/utils/evaluation.py:359-378) on the GPU: kp1, kp2 (B, k, 2) int32 (y, x), match_idx (B, k) int32
    (row of kp2 matched to kp1 row i, -1 = none), n1 (B,) valid rows of kp1.  Deterministic LO-RANSAC, no host sync."""
    dev = _lib.require_cuda(kp1, kp2, match_idx, n1)
    if kp1.dim() != 3 or kp1.shape[2] != 2 or kp2.dim() != 3 or kp2.shape[0] != kp1.shape[0] or kp2.shape[2] != 2:
        raise RuntimeError("estimate_homography: kp1 / kp2 must be (B, k, 2) keypoint tensors")
    B, k, _ = kp1.shape
    if tuple(match_idx.shape) != (B, k):
        raise RuntimeError("estimate_homography: match_idx must be (B, k)")
    kp1 = kp1.to(torch.int32).contiguous()
    kp2 = kp2.to(torch.int32).contiguous()
    match_idx = match_idx.to(torch.int32).contiguous()
    n1 = None if n1 is None else n1.to(torch.int32).contiguous()
    H = torch.empty((B, 3, 3), dtype=torch.float64, device=dev)
    inl = torch.empty((B, k), dtype=torch.uint8, device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    if B:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_estimate_homography(_lib.ptr(kp1), _lib.ptr(kp2), _lib.ptr(n1), _lib.ptr(match_idx), B, k,
                                                         int(height), int(width), int(iters), float(reproj_threshold),
                                                         int(lo_rounds), int(seed) & 0xffffffff, _lib.ptr(H), _lib.ptr(inl),
                                                         _lib.ptr(cnt), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    return Homographies(H, inl.bool(), cnt)


def find_homography(src_pts, dst_pts, reproj_threshold: float = 3.0, height: Optional[int] = None, width: Optional[int] = None,
                    iters: int = 2048, seed: int = 0):
    """cv2.findHomography-shaped convenience for one pair (evaluation.py:368-375): src_pts, dst_pts (N, 2) or (N, 1, 2)
    CUDA tensors of (x, y) pixel coordinates (integers, as keypoints are) -> (H (3, 3) float64 or None, mask (N, 1) uint8)."""
    s = src_pts.reshape(-1, 2)
    d = dst_pts.reshape(-1, 2)
    n = s.shape[0]
    if n < 4:
        return None, torch.zeros((n, 1), dtype=torch.uint8, device=s.device)
    if height is None:
        height = int(max(float(s[:, 1].max()), float(d[:, 1].max()))) + 1
    if width is None:
        width = int(max(float(s[:, 0].max()), float(d[:, 0].max()))) + 1
    kp1 = s.flip(-1).round().to(torch.int32)[None]
    kp2 = d.flip(-1).round().to(torch.int32)[None]
    idx = torch.arange(n, dtype=torch.int32, device=s.device)[None]
    r = estimate_homography(kp1, kp2, idx, height, width, None, iters, reproj_threshold, 2, seed)
    if int(r.n_inliers[0]) < 0:
        return None, torch.zeros((n, 1), dtype=torch.uint8, device=s.device)
    return r.H[0], r.inliers[0].to(torch.uint8)[:, None]
