"""Batched evaluation metrics on the device.

The reference evaluates one sample at a time in Python after copying keypoints to the host
(xpoint/utils/benchmark_evaluation.py:396-467 ``compute_repeatability_for_sample``, :588-750
``compute_descriptor_for_sample``; ``warp_keypoints`` / ``filter_points`` in xpoint/utils/homographies.py:479-526).
Here the same quantities are computed for a whole batch of pairs from the tensors the pipeline already holds
(``PairResult``: keypoints (B, k, 2) int32 (y, x) + counts, mutual-match indices), with no host synchronisation:

  warp_keypoints(kp, H, ...)             cv2.perspectiveTransform semantics in float64, int truncation, inside-image mask
  repeatability(...)                     per pair and threshold: (count1 + count2) / (N_thermal + N_optical)
  matching_scores(...)                   per pair and threshold: correct matches, M-score, n_gt for both directions
  evaluate_pairs(result, H_o, H_t, ...)  the two above for a PairResult
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import torch

from . import _lib


def _as_h64(H: torch.Tensor, B: int, dev) -> torch.Tensor:
    H = H.to(device=dev, dtype=torch.float64)
    if H.dim() == 2:
        H = H.unsqueeze(0).expand(B, 3, 3)
    if tuple(H.shape) != (B, 3, 3):
        raise RuntimeError("homographies must be (B, 3, 3) or (3, 3)")
    return H.contiguous()


class Warped(NamedTuple):
    points_float: Optional[torch.Tensor]   # (B, k, 2) float64 (y', x')
    points_int: Optional[torch.Tensor]     # (B, k, 2) int32, numpy astype(int) of the above
    inside: torch.Tensor                   # (B, k) bool: filter_points


def warp_keypoints(keypoints: torch.Tensor, homography: torch.Tensor, count: Optional[torch.Tensor] = None,
                   height: Optional[int] = None, width: Optional[int] = None, want_float: bool = True, want_int: bool = True,
                   filter_on_float: bool = False) -> Warped:
    """Batched ``warp_keypoints`` (homographies.py:479-495) + ``filter_points`` (:511-526): keypoints (B, k, 2) or (k, 2)
    integer (y, x), homography (B, 3, 3) / (3, 3).  Rows at or beyond ``count[b]`` are zero / outside."""
    dev = _lib.require_cuda(keypoints, count)
    single = keypoints.dim() == 2
    kp = (keypoints[None] if single else keypoints).to(torch.int32).contiguous()
    B, k, _ = kp.shape
    H = _as_h64(homography, B, dev)
    height = int(height if height is not None else 1 << 30)
    width = int(width if width is not None else 1 << 30)
    out_f = torch.empty((B, k, 2), dtype=torch.float64, device=dev) if want_float else None
    out_i = torch.empty((B, k, 2), dtype=torch.int32, device=dev) if want_int else None
    inside = torch.empty((B, k), dtype=torch.uint8, device=dev)
    cnt = None if count is None else count.to(torch.int32).contiguous()
    if B and k:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().xp_warp_keypoints(_lib.ptr(kp), _lib.ptr(cnt), _lib.ptr(H), B, k, height, width, _lib.ptr(out_f),
                                                    _lib.ptr(out_i), _lib.ptr(inside), int(filter_on_float), _lib.stream_ptr(dev)))
        _lib.count_launches(1)
    inside = inside.bool()
    if single:
        return Warped(None if out_f is None else out_f[0], None if out_i is None else out_i[0], inside[0])
    return Warped(out_f, out_i, inside)


class Repeatability(NamedTuple):
    repeatability: torch.Tensor   # (B, T) float64, NaN where N_thermal + N_optical == 0 (the reference appends nothing then)
    count1: torch.Tensor          # (B, T) int32: warped thermal points with an optical keypoint within the threshold
    count2: torch.Tensor          # (B, T) int32: warped optical points with a thermal keypoint within the threshold
    n_warped_thermal: torch.Tensor   # (B,) int32 = N_thermal (warped thermal points inside the image)
    n_warped_optical: torch.Tensor   # (B,) int32 = N_optical


def _thr(thresholds, dev):
    t = [float(x) for x in (thresholds if isinstance(thresholds, (list, tuple)) else [thresholds])]
    if not 1 <= len(t) <= 8:
        raise ValueError("1..8 distance thresholds")
    return torch.tensor(t, dtype=torch.float64, device=dev), len(t)


def repeatability(kp_optical, n_optical, kp_thermal, n_thermal, H_optical, H_thermal, height: int, width: int,
                  thresholds: Sequence[float] = (3,)) -> Repeatability:
    """compute_repeatability_for_sample (benchmark_evaluation.py:396-467) for a batch: optical keypoints are warped to the
    thermal frame by H_thermal o H_optical^-1 (two ``warp_keypoints`` calls with the integer truncation in between, exactly
    as the reference chains them) and vice versa, filtered to the image, and counted when a keypoint of the other image
    lies within each threshold."""
    dev = _lib.require_cuda(kp_optical, kp_thermal, n_optical, n_thermal)
    kp_o, kp_t = kp_optical.to(torch.int32).contiguous(), kp_thermal.to(torch.int32).contiguous()
    B, k, _ = kp_o.shape
    if kp_t.shape[1] != k:
        raise RuntimeError("repeatability: both keypoint tensors need the same capacity k")
    n_o, n_t = n_optical.to(torch.int32).contiguous(), n_thermal.to(torch.int32).contiguous()
    # the reference inverts in the dtype of the data homographies (float32 tensors) and hands the result to OpenCV
    Ho32, Ht32 = H_optical.to(dev).float(), H_thermal.to(dev).float()
    if Ho32.dim() == 2:
        Ho32, Ht32 = Ho32.expand(B, 3, 3), Ht32.expand(B, 3, 3)
    Ho_inv, Ht_inv = torch.linalg.inv(Ho32), torch.linalg.inv(Ht32)
    thr, T = _thr(thresholds, dev)

    def chain(kp, n, Hinv, Hfwd):
        first = warp_keypoints(kp, Hinv, n, want_float=False)                                   # int, no filtering
        return warp_keypoints(first.points_int, Hfwd, n, height, width, want_float=False)       # int + filter_points

    w_o = chain(kp_o, n_o, Ho_inv, Ht32)          # optical keypoints in the thermal frame
    w_t = chain(kp_t, n_t, Ht_inv, Ho32)          # thermal keypoints in the optical frame
    c1 = torch.empty((B, T), dtype=torch.int32, device=dev)
    c2 = torch.empty((B, T), dtype=torch.int32, device=dev)
    N_t = torch.empty((B,), dtype=torch.int32, device=dev)
    N_o = torch.empty((B,), dtype=torch.int32, device=dev)
    if B:
        lib = _lib.lib()
        with torch.cuda.device(dev):
            ins_t, ins_o = w_t.inside.to(torch.uint8).contiguous(), w_o.inside.to(torch.uint8).contiguous()
            _lib.check(lib.xp_repeatability_counts(_lib.ptr(w_t.points_int), _lib.ptr(ins_t), _lib.ptr(n_t), _lib.ptr(kp_o),
                                                   _lib.ptr(n_o), B, k, _lib.ptr(thr), T, _lib.ptr(c1), _lib.ptr(N_t),
                                                   _lib.stream_ptr(dev)))
            _lib.check(lib.xp_repeatability_counts(_lib.ptr(w_o.points_int), _lib.ptr(ins_o), _lib.ptr(n_o), _lib.ptr(kp_t),
                                                   _lib.ptr(n_t), B, k, _lib.ptr(thr), T, _lib.ptr(c2), _lib.ptr(N_o),
                                                   _lib.stream_ptr(dev)))
        _lib.count_launches(2)
    denom = (N_t + N_o).double().unsqueeze(1)
    rep = torch.where(denom > 0, (c1 + c2).double() / denom.clamp(min=1), torch.full_like(denom, float("nan")).expand(B, T))
    return Repeatability(rep, c1, c2, N_t, N_o)


class MatchScores(NamedTuple):
    n_correct_optical: torch.Tensor   # (B, T) int32 correct optical -> thermal matches
    n_correct_thermal: torch.Tensor   # (B, T)
    m_score_optical: torch.Tensor     # (B, T) float64 = n_correct / N_optical (0 where N_optical == 0)
    m_score_thermal: torch.Tensor
    n_gt_optical: torch.Tensor        # (B, T) optical keypoints with a thermal keypoint within the threshold of their warp
    n_gt_thermal: torch.Tensor
    n_possible_optical: torch.Tensor  # (B,) warped optical keypoints inside the image
    n_possible_thermal: torch.Tensor
    n_matches: torch.Tensor           # (B,)


def matching_scores(kp_optical, n_optical, kp_thermal, n_thermal, match_idx, H_optical, H_thermal, height: int, width: int,
                    thresholds: Sequence[float] = (2,)) -> MatchScores:
    """The correctness / M-score part of compute_descriptor_for_sample (benchmark_evaluation.py:650-690) for a batch.
    ``match_idx`` (B, k): thermal keypoint matched to optical keypoint i or -1 (mutual matches are symmetric, so the thermal
    -> optical list is its inverse).  gt_homography = H_thermal @ H_optical^-1 (float32, as the reference computes it)."""
    dev = _lib.require_cuda(kp_optical, kp_thermal, n_optical, n_thermal, match_idx)
    kp_o, kp_t = kp_optical.to(torch.int32).contiguous(), kp_thermal.to(torch.int32).contiguous()
    B, k, _ = kp_o.shape
    n_o, n_t = n_optical.to(torch.int32).contiguous(), n_thermal.to(torch.int32).contiguous()
    m_ot = match_idx.to(torch.int32).contiguous()
    # inverse match list: thermal j -> optical i
    m_to = torch.full((B, kp_t.shape[1]), -1, dtype=torch.int32, device=dev)
    rows = torch.arange(k, device=dev, dtype=torch.int32).unsqueeze(0).expand(B, k)
    valid = (m_ot >= 0) & (rows < n_o.unsqueeze(1))
    bidx = torch.arange(B, device=dev).unsqueeze(1).expand(B, k)
    m_to[bidx[valid], m_ot[valid].long()] = rows[valid]
    Ho32, Ht32 = H_optical.to(dev).float(), H_thermal.to(dev).float()
    if Ho32.dim() == 2:
        Ho32, Ht32 = Ho32.expand(B, 3, 3), Ht32.expand(B, 3, 3)
    gt = torch.bmm(Ht32, torch.linalg.inv(Ho32))
    gt_inv = torch.linalg.inv(gt)
    thr, T = _thr(thresholds, dev)
    w_o = warp_keypoints(kp_o, gt, n_o, height, width, want_int=False, filter_on_float=True)
    w_t = warp_keypoints(kp_t, gt_inv, n_t, height, width, want_int=False, filter_on_float=True)
    outs = [torch.empty((B, T), dtype=torch.int32, device=dev) for _ in range(4)]
    poss = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(2)]
    nm = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(2)]
    if B:
        lib = _lib.lib()
        with torch.cuda.device(dev):
            for (w, nq, kt, nt, mi, nc, ng, npos, nmm) in ((w_o, n_o, kp_t, n_t, m_ot, outs[0], outs[2], poss[0], nm[0]),
                                                           (w_t, n_t, kp_o, n_o, m_to, outs[1], outs[3], poss[1], nm[1])):
                ins = w.inside.to(torch.uint8).contiguous()
                _lib.check(lib.xp_match_score_counts(_lib.ptr(w.points_float), _lib.ptr(ins), _lib.ptr(nq), _lib.ptr(kt),
                                                     _lib.ptr(nt), _lib.ptr(mi), B, k, _lib.ptr(thr), T, _lib.ptr(nc), _lib.ptr(ng),
                                                     _lib.ptr(npos), _lib.ptr(nmm), _lib.stream_ptr(dev)))
        _lib.count_launches(2)

    def score(nc, npos):
        d = npos.double().unsqueeze(1)
        return torch.where(d > 0, nc.double() / d.clamp(min=1), torch.zeros_like(nc, dtype=torch.float64))
    return MatchScores(outs[0], outs[1], score(outs[0], poss[0]), score(outs[1], poss[1]), outs[2], outs[3], poss[0], poss[1], nm[0])


def evaluate_pairs(result, H_optical, H_thermal, height: int, width: int, thresh_repeatability=(3,), thresh_keypoints=(2,)):
    """Repeatability and matching scores of a ``PairResult`` (the per-batch body of compute_metrics,
    benchmark_evaluation.py:832-931), as device tensors."""
    rep = repeatability(result.kp_optical, result.n_optical, result.kp_thermal, result.n_thermal, H_optical, H_thermal, height,
                        width, thresh_repeatability)
    ms = matching_scores(result.kp_optical, result.n_optical, result.kp_thermal, result.n_thermal, result.match_idx, H_optical,
                         H_thermal, height, width, thresh_keypoints)
    return {"repeatability": rep, "matching": ms}
