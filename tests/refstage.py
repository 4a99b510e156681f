"""Loader for the staged, UNMODIFIED reference (test / bench infrastructure, never imported by the product).

`oracle/build_ref.py` (run in the build container) leaves
  * oracle/_ref/selective_scan_cuda_oflex.*.so -- the reference's CUDA extension compiled for sm_100a from
    /root/reference/xpoint/models/vmamba_src/kernels/selective_scan/csrc (its own sources, its own flags);
  * baseline/_ref/{xpoint, timm, fvcore, yacs}  -- the reference's Python package + the three import shims.
Both are git-ignored and travel to the GPU box with the snapshot.  Nothing here reads /root/reference.

With the extension importable the reference takes its own CUDA path (csms6s.py:9-22,112-126: WITH_SELECTIVESCAN_OFLEX)
and its own Triton CrossScan/CrossMerge (csm_triton.py:403-517) on a GPU -- i.e. exactly what a user of the reference
runs on this box.
"""
import contextlib
import io
import os
import sys
import tempfile
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY_DIR = os.path.join(ROOT, "baseline", "_ref")
EXT_DIR = os.path.join(ROOT, "oracle", "_ref")

# VSSM trees of the two presets (E = model_weights/XPoint-EXP1/params.yaml:107-129, V = vanilla_vmamba_tiny,
# VMamba.py:1651-1662)
VSSM_PRESETS = dict(
    E=dict(DEPTHS=[2, 2, 2, 2], DOWNSAMPLE="v3", EMBED_DIM=96, MLP_RATIO=4.0, PATCHEMBED="v2", SSM_CONV=3,
           SSM_CONV_BIAS=False, SSM_DT_RANK="auto", SSM_D_STATE=1, SSM_FORWARDTYPE="v05_noz", SSM_RATIO=1.0),
    V=dict(DEPTHS=[2, 2, 9, 2], DOWNSAMPLE="v1", EMBED_DIM=96, MLP_RATIO=0.0, PATCHEMBED="v1", SSM_CONV=3,
           SSM_CONV_BIAS=True, SSM_DT_RANK="auto", SSM_D_STATE=16, SSM_FORWARDTYPE="v0", SSM_RATIO=2.0,
           SSM_INIT="v0", NORM_LAYER="ln"),
)

_state = {}


def ext_path():
    if not os.path.isdir(EXT_DIR):
        return None
    for f in os.listdir(EXT_DIR):
        if f.startswith("selective_scan_cuda_oflex") and f.endswith(".so"):
            return os.path.join(EXT_DIR, f)
    return None


def available(need_ext=True, need_py=True):
    ok = True
    if need_ext:
        ok = ok and ext_path() is not None
    if need_py:
        ok = ok and os.path.isfile(os.path.join(PY_DIR, "xpoint", "models", "XPoint.py"))
    return ok


def why_missing():
    return ("staged reference not present (run `python oracle/build_ref.py` in the build container: "
            f"ext={ext_path() is not None}, python={os.path.isdir(os.path.join(PY_DIR, 'xpoint'))})")


def load_ext():
    """import selective_scan_cuda_oflex (the reference's pybind module, selective_scan_oflex.cpp:357-360)."""
    if "ext" not in _state:
        import torch  # noqa: F401  (the extension links against libtorch)
        if EXT_DIR not in sys.path:
            sys.path.insert(0, EXT_DIR)
        import selective_scan_cuda_oflex
        _state["ext"] = selective_scan_cuda_oflex
    return _state["ext"]


def load(with_ext=True, patch_cross_torch=False):
    """Import the staged reference package.  Returns a namespace with models, utils, VMamba (RV), csm_triton (RC),
    csms6s (RS) and the extension module (or None)."""
    if "ns" in _state:
        return _state["ns"]
    ext = load_ext() if with_ext and ext_path() else None
    if PY_DIR not in sys.path:
        sys.path.insert(0, PY_DIR)
    warnings.filterwarnings("ignore")
    hidden, had_path = None, EXT_DIR in sys.path
    if not with_ext:
        # the reference decides at import time whether it has a CUDA extension (csms6s.py:9-22); for its CPU / torch path the
        # extension must not be importable while the package loads, even if this process has already loaded it for timing
        hidden = sys.modules.pop("selective_scan_cuda_oflex", None)
        if had_path:
            sys.path.remove(EXT_DIR)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import xpoint.models as xmodels
            import xpoint.utils as xutils
            from xpoint.models.vmamba_src import VMamba as RV
            from xpoint.models.vmamba_src import csm_triton as RC
            from xpoint.models.vmamba_src import csms6s as RS
    finally:
        if hidden is not None:
            sys.modules["selective_scan_cuda_oflex"] = hidden
        if not with_ext and had_path:
            sys.path.insert(0, EXT_DIR)
    assert os.path.realpath(xmodels.__file__).startswith(os.path.realpath(PY_DIR)), xmodels.__file__
    ns = types.SimpleNamespace(models=xmodels, utils=xutils, RV=RV, RC=RC, RS=RS, ext=ext)
    if patch_cross_torch:
        patch_cross_torch_path(ns)
    _state["ns"] = ns
    return ns


def patch_cross_torch_path(ns):
    """cross_scan_fn wraps its call in torch.cuda.device(x.device), which raises on CPU tensors; rebinding
    VMamba's two names to the reference's own torch autograd Functions is the one behavioural patch (also the fallback
    when Triton cannot JIT on the box)."""
    RV, RC = ns.RV, ns.RC
    RV.cross_scan_fn = lambda x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False: \
        RC.CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)
    RV.cross_merge_fn = lambda y, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0, force_torch=False: \
        RC.CrossMergeF.apply(y, in_channel_first, out_channel_first, one_by_one, scans)
    ns.cross_patched = True


def build_xpoint(ns, preset="E", mixed_precision=False, height=512, width=640, seed=0, vssm=None):
    """The reference's XPoint(config) (XPoint.py:27-178) with a VMamba encoder of the given preset, random init (seed),
    homography head off, eval mode."""
    import torch
    import yaml
    vssm = dict(vssm or VSSM_PRESETS[preset])
    with tempfile.TemporaryDirectory() as td:
        ypath = os.path.join(td, "vssm.yaml")
        with open(ypath, "w") as f:
            yaml.safe_dump({"MODEL": {"TYPE": "vssm", "NAME": "staged", "DROP_PATH_RATE": 0.2, "VSSM": vssm}}, f)
        cfg = dict(
            multispectral=False, descriptor_head=True, descriptor_size=256, normalize_descriptors=True,
            final_batchnorm=True, reflection_pad=True, bn_first=False, mixed_precision=mixed_precision, takes_pair=True,
            homography_regression_head=dict(check=False, type="RegNet"),
            use_attention=dict(check=True, type="VMamba", height=height, width=width,
                               pretrained=dict(check=True, yaml_file=ypath), model_parameters={}),
        )
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            net = ns.models.XPoint(cfg).eval()
    return net


def randomise_stats(net, seed=1):
    """Move BatchNorm running stats and the SSM parameters away from their init values (A = -(n+1), Ds = 1, zero-mean
    unit-variance BN) so that parity does not pass on initialisation structure."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=g))
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
        for n_, p in net.named_parameters():
            if n_.endswith("A_logs"):
                p.copy_(torch.log(0.5 * torch.rand(p.shape, generator=g) + 0.05))
            if n_.endswith("Ds"):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
    return net
