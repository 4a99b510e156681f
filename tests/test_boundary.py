"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, argument validation maps to the reference's error behaviour, and the product never touches the oracle."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "xpoint_b200.h")).read()
    return sorted(set(re.findall(r"^XP_API\s+[\w\s\*]+?\b(xp_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from xpoint_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/xpoint_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == syms, "ctypes signature table out of sync with the header"
    assert _lib.lib().xp_abi_version() == 5


def test_header_compiles_as_plain_c(tmp_path):
    """The boundary is a C ABI: the header must be valid C (no C++/torch types)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "xpoint_b200.h"\nint main(void){ xp_scan_args a; (void)a; return XP_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_scan_args_struct_layout_matches_header():
    from xpoint_b200 import _lib
    src = open(os.path.join(ROOT, "include", "xpoint_b200.h")).read()
    body = src[src.index("typedef struct {"):src.index("} xp_scan_args;")]
    names = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0].strip().rstrip(";")
        if not line:
            continue
        decl = line.replace("const ", "")
        _, rest = decl.split(None, 1)
        names += [n.strip().lstrip("*") for n in rest.split(",")]
    assert names == [f[0] for f in _lib.ScanArgs._fields_]
    assert ctypes.sizeof(_lib.ScanArgs) == 10 * 8 + 20 * 8 + 4 * 4 + 3 * 8 + 3 * 8


def test_validation_errors_without_gpu():
    """Argument checks run before any CUDA call, so they are testable on the build box."""
    from xpoint_b200 import _lib
    lib = _lib.lib()
    assert lib.xp_selective_scan_fwd(None, None) == _lib.XP_ERR_INVALID_ARG
    a = _lib.ScanArgs()
    assert lib.xp_selective_scan_fwd(ctypes.byref(a), None) == _lib.XP_ERR_INVALID_ARG
    assert b"non-NULL" in lib.xp_last_error()
    for k in ("u", "delta", "A", "B", "C", "out"):
        setattr(a, k, 16)
    a.batch, a.dim, a.delta_dim, a.groups, a.dstate, a.seqlen = 1, 8, 8, 3, 4, 16
    assert lib.xp_selective_scan_fwd(ctypes.byref(a), None) == _lib.XP_ERR_INVALID_ARG      # dim % groups
    a.groups, a.dstate = 2, 300
    assert lib.xp_selective_scan_fwd(ctypes.byref(a), None) == _lib.XP_ERR_INVALID_ARG      # dstate > 256
    assert b"256" in lib.xp_last_error()
    assert lib.xp_selective_scan_bwd() == _lib.XP_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _lib.check(lib.xp_selective_scan_bwd())
    assert lib.xp_box_nms(None, None, 1, 4, 4, 8.0, 0.1, 0.1, 0, 0.0, None, None, 0, None, 0, None) == _lib.XP_ERR_INVALID_ARG
    assert lib.xp_nms_workspace_bytes(2, 10, 10) == 400      # state + scan-cursor byte planes
    assert lib.xp_mnn_match(16, 16, None, None, 1, 8, 8, 256, None, None, None, None, None, 1, None, 0, None) == _lib.XP_ERR_WORKSPACE


def test_ops_refuse_cpu_tensors():
    import xpoint_b200 as X
    u = torch.zeros(1, 4, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        X.selective_scan_fn(u, u, torch.zeros(4, 1), torch.zeros(1, 1, 1, 8), torch.zeros(1, 1, 1, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        X.cross_scan_fn(torch.zeros(1, 2, 3, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        X.box_nms(torch.zeros(8, 8), 8, 0.015)
    with pytest.raises(RuntimeError, match="CUDA"):
        X.interpolate_descriptors(torch.zeros(3, 2, dtype=torch.long), torch.zeros(16, 4, 4), 32, 32)
    with pytest.raises(RuntimeError):
        X.SS2D(d_model=16, d_state=1, ssm_ratio=1.0, forward_type="v05_noz")(torch.zeros(1, 4, 4, 16))


def test_api_signatures_match_reference():
    import inspect

    import xpoint_b200 as X
    assert list(inspect.signature(X.cross_scan_fn).parameters) == ["x", "in_channel_first", "out_channel_first", "one_by_one",
                                                                   "scans", "force_torch"]
    assert list(inspect.signature(X.cross_merge_fn).parameters) == ["y", "in_channel_first", "out_channel_first", "one_by_one",
                                                                    "scans", "force_torch"]
    assert list(inspect.signature(X.box_nms).parameters) == ["prob", "size", "min_prob", "iou", "keep_top_k", "on_cpu"]
    assert list(inspect.signature(X.interpolate_descriptors).parameters) == ["keypoints", "descriptors_lowres", "H", "W"]
    assert list(inspect.signature(X.get_matches).parameters)[:4] == ["desc_1", "desc_2", "method", "knn_matches"]
    assert list(inspect.signature(X.selective_scan_fn_mamba).parameters) == [
        "u", "delta", "A", "B", "C", "D", "z", "delta_bias", "delta_softplus", "return_last_state"]


def test_model_state_dict_matches_reference_keys():
    import numpy as np

    import xpoint_b200 as X
    from test_gpu_model import TINY
    for tag in "EV":
        g = np.load(os.path.join(ROOT, "tests", "golden", f"xpoint_tiny_{tag}.npz"))
        ref = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
        net = X.XPoint({"takes_pair": True,
                        "use_attention": {"model_parameters": {"MODEL": {"DROP_PATH_RATE": 0.2, "VSSM": TINY[tag]}}}})
        mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert mine == {k: tuple(v) for k, v in ref.items()}
    full = X.XPoint({"takes_pair": True})     # preset E = shipped XPoint-EXP1 encoder
    n = sum(p.numel() for p in full.parameters())
    assert 20_000_000 < n < 21_000_000
    assert full.takes_pair() and full.get_encoder_downsample_ratio() == 8


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/ (the oracle is a checker, not a fallback)."""
    pkg = os.path.join(ROOT, "xpoint_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower().replace("# no oracle", ""), f"{f} mentions the oracle"
    assert "/root/reference" not in open(os.path.join(ROOT, "bench.py")).read()


def test_algorithmic_bytes_formula():
    """B*L*[(2*KD + 2*K*N)*s_in + KD*s_out] + 4*KD*(N+2); config 2 fp32 = 6.375 GB, bf16->fp32 = 4.194 GB; with
    the fused dt_proj the delta term KD*s_in becomes K*R*s_in (+ the weights)."""
    from xpoint_b200.selective_scan import algorithmic_bytes
    assert abs(algorithmic_bytes(32, 768, 4, 16, 20480, 4, 4) / 1e9 - 6.375) < 1e-3
    assert abs(algorithmic_bytes(32, 768, 4, 16, 20480, 2, 4) / 1e9 - 4.194) < 1e-3
    assert abs(algorithmic_bytes(32, 768, 4, 16, 20480, 2, 2) / 1e9 - 3.188) < 1e-3
    plain = algorithmic_bytes(128, 384, 4, 1, 20480, 2, 4)
    fused = algorithmic_bytes(128, 384, 4, 1, 20480, 2, 4, dt_rank=6)
    assert plain - fused == 128 * 20480 * (384 - 4 * 6) * 2 - 384 * 6 * 2
