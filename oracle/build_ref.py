#!/usr/bin/env python
"""Stage the UNMODIFIED reference next to the oracle (test infrastructure; nothing here is product code).

Run in the build container (the GPU box has no /root/reference; it only uses what this script left behind):

    python oracle/build_ref.py            # both steps
    python oracle/build_ref.py --ext      # only the CUDA extension
    python oracle/build_ref.py --py       # only the Python package

1. oracle/_ref/selective_scan_cuda_oflex.<abi>.so -- the reference's own selective-scan CUDA extension
   (xpoint/models/vmamba_src/kernels/selective_scan/csrc/selective_scan/cusoflex/{selective_scan_oflex.cpp,
   selective_scan_core_fwd.cu, selective_scan_core_bwd.cu}), compiled from the sources WHERE THEY LIE under
   /root/reference with the flags of its setup.py:119-136, for sm_100a (setup.py:68 would emit -arch=sm_<cc of the
   build machine>; there is no GPU here, nvcc cross-compiles).  Not the reference's build system: three nvcc/g++
   commands and a link.  This is "the kernel to beat on the same box"  and the GPU-side
   parity anchor of tests/test_gpu_reference.py.
2. baseline/_ref/xpoint/ -- the reference's Python package (the equivalent of `pip install --target baseline/_ref`;
   the reference's setup.py does not package `xpoint.models.vmamba_src` data files, so it is a plain copy of the
   package directory) plus the three import shims the image lacks (timm / fvcore / yacs, tests/golden/_shims).
   Git-ignored, travels to the GPU box with the snapshot.  The GPU parity tests import XPoint / VSSM / box_nms /
   get_matches from there and run them on the B200 beside this repo's kernels; bench.py --impl reference times it.

Both output directories are git-ignored (.gitignore: oracle/_ref/, baseline/_ref/): no reference source enters history.
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("XPOINT_REFERENCE", "/root/reference")
KSRC = os.path.join(REF, "xpoint/models/vmamba_src/kernels/selective_scan/csrc/selective_scan")
OUT_EXT = os.path.join(HERE, "_ref")
OUT_PY = os.path.join(ROOT, "baseline", "_ref")
NAME = "selective_scan_cuda_oflex"


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_ext():
    import torch
    from torch.utils import cpp_extension as C
    os.makedirs(os.path.join(OUT_EXT, "obj"), exist_ok=True)
    inc = [f"-I{p}" for p in C.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{KSRC}"]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc_flags = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                  "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                  "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                  "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo",
                  "-gencode", "arch=compute_100a,code=sm_100a", "--threads", "4",
                  "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
    objs = []
    procs = []
    for src in ("cusoflex/selective_scan_core_fwd.cu", "cusoflex/selective_scan_core_bwd.cu"):
        obj = os.path.join(OUT_EXT, "obj", os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = ["nvcc", "-c", os.path.join(KSRC, src), "-o", obj] + nvcc_flags + inc + defs
        print("+", " ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    obj = os.path.join(OUT_EXT, "obj", "selective_scan_oflex.cpp.o")
    objs.append(obj)
    run(["/usr/bin/g++", "-c", os.path.join(KSRC, "cusoflex/selective_scan_oflex.cpp"), "-o", obj, "-O3", "-std=c++17",
         "-fPIC"] + inc + defs)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    libdirs = C.library_paths("cuda")
    so = os.path.join(OUT_EXT, NAME + sysconfig.get_config_var("EXT_SUFFIX"))
    run(["/usr/bin/g++", "-shared", "-o", so] + objs + [f"-L{p}" for p in libdirs]
        + [f"-Wl,-rpath,{libdirs[0]}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-lcudart"])
    shutil.rmtree(os.path.join(OUT_EXT, "obj"))
    print("built", so)


def stage_py():
    dst = os.path.join(OUT_PY, "xpoint")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(OUT_PY, exist_ok=True)
    shutil.copytree(os.path.join(REF, "xpoint"), dst,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "datasets", "csrc"))
    # configs the reference reads at model construction (XPoint.py:435 -> MYCONFIG.get_config needs a yaml file)
    for sub in ("configs", "model_weights"):
        d = os.path.join(OUT_PY, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(REF, sub), d, ignore=shutil.ignore_patterns("*.model", "*.pth", "*.png"))
    shims = os.path.join(ROOT, "tests", "golden", "_shims")
    for name in ("timm", "fvcore", "yacs"):
        d = os.path.join(OUT_PY, name)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(shims, name), d, ignore=shutil.ignore_patterns("__pycache__"))
    for root, dirs, files in os.walk(OUT_PY):          # /root/reference is read-only; the copy must stay replaceable
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    print("staged", OUT_PY)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ext", action="store_true")
    ap.add_argument("--py", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print(f"{REF} not present: nothing to stage (the GPU box uses the prebuilt files)")
        sys.exit(0)
    both = not (a.ext or a.py)
    if a.py or both:
        stage_py()
    if a.ext or both:
        build_ext()
