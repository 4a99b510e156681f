"""xpoint_b200 -- B200-native (sm_100a) implementation of XPoint's inference hot path.

Operator API (names and signatures of the reference):
    selective_scan_fn, selective_scan_fn_mamba, SelectiveScanCuda, selective_scan_cuda_oflex
    cross_scan_fn, cross_merge_fn, CrossScanF, CrossMergeF, CrossScanTritonF, CrossMergeTritonF
    box_nms, interpolate_descriptors, get_matches, NNMatcher
    SS2D, VSSBlock, VSSM, XPoint, PairPipeline, PairStream, ShardedPairPipeline, metrics (warp_keypoints, repeatability, ...)
All of them call libxpoint_b200.so (include/xpoint_b200.h) through ctypes; there is no CPU fallback.
"""
from .cross_scan import (CrossMergeF, CrossMergeTritonF, CrossScanF, CrossScanTritonF, cross_merge_fn, cross_scan_fn,
                         merge_norm_gate)
from .postprocess import (DMatch, NNMatcher, box_nms, detector_post, estimate_homography, find_homography, get_matches,
                          interpolate_descriptors, mnn_match, nms_keypoints, normalize_descriptors, sample_descriptors)
from .selective_scan import (SelectiveScanCuda, selective_scan_cuda_oflex, selective_scan_fn, selective_scan_fn_csms6s,
                             selective_scan_fn_mamba)
from .vmamba import PRESETS, SS2D, VSSM, VSSBlock, build_vssm
from . import metrics
from .pipeline import PairStream, ShardedPairPipeline
from .xpoint import GraphedPairPipeline, PairPipeline, PairResult, XPoint

__all__ = [n for n in dir() if not n.startswith("_")]
