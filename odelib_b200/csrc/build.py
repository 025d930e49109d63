#!/usr/bin/env python
"""Build libodelib_b200.so in-tree (odelib_b200/libodelib_b200.so) with nvcc for sm_100a.

The device kernels of the hot path are compiled at run time by NVRTC (the model's right-hand side is
only known then); their source (odl_kernels.cuh + odl_abi.h) is embedded into the library as string
literals here.  ``check_kernels()`` additionally compiles the kernels for the three demo models with
nvcc -Xptxas -v so that register / spill regressions show up at build time, without a GPU.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libodelib_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _embed(src, dst):
    text = open(os.path.join(HERE, src)).read()
    assert ')ODLSRC"' not in text
    # one raw literal per ~8 kB keeps every compiler's string-literal limit out of the picture
    chunks, cur = [], []
    size = 0
    for line in text.splitlines(keepends=True):
        cur.append(line); size += len(line)
        if size > 8000:
            chunks.append("".join(cur)); cur, size = [], 0
    chunks.append("".join(cur))
    body = "\n".join('R"ODLSRC(' + c + ')ODLSRC"' for c in chunks) + "\n"
    path = os.path.join(HERE, dst)
    if not os.path.exists(path) or open(path).read() != body:
        open(path, "w").write(body)


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    _embed("odl_abi.h", "odl_abi_embedded.inc")
    _embed("odl_kernels.cuh", "odl_kernels_embedded.inc")
    srcs = [os.path.join(HERE, f) for f in ("odl_capi.cu", "odl_abi.h", "odl_kernels.cuh", "build.py")]
    srcs.append(os.path.join(ROOT, "include", "odelib_b200.h"))
    if not force and not _stale(LIB, srcs):
        return LIB
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
    cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", LIB,
           os.path.join(HERE, "odl_capi.cu"), "-L" + cuda_lib, "-lnvrtc", "-Xlinker", "-rpath=" + cuda_lib]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


def check_kernels(models=None, outdir=None, block=128, minblocks=4, dense=1):
    """nvcc -Xptxas -v of the integrator kernels for traced demo models -> {model: {kernel: (regs, spill_bytes)}}."""
    sys.path.insert(0, ROOT)
    from odelib_b200 import demo_models
    from odelib_b200.tracer import trace
    outdir = outdir or os.path.join(HERE, "_check")
    os.makedirs(outdir, exist_ok=True)
    report = {}
    for name in models or ("zero_i", "one_i", "two_i"):
        f, n, P, groups = demo_models.MODELS[name]
        src = trace(f, n, P).cuda_source(fmad=True, observe_groups=groups)
        cu = os.path.join(outdir, f"{name}.cu")
        open(cu, "w").write(src + '#include "odl_kernels.cuh"\n')
        cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-I" + HERE, f"-DODL_BLOCK={block}",
               f"-DODL_MINBLOCKS={minblocks}", f"-DODL_DENSE={dense}", "-DODL_Y0P=0", "-Xptxas", "-v", "-cubin", "-o",
               os.path.join(outdir, f"{name}.cubin"), cu]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{res.stderr}")
        info, cur = {}, None
        for line in res.stderr.splitlines():
            mm = re.search(r"Compiling entry function '(\w+)'", line)
            if mm:
                cur = mm.group(1)
            mm = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if mm and cur:
                info.setdefault(cur, {})["spill"] = int(mm.group(1)) + int(mm.group(2))
            mm = re.search(r"Used (\d+) registers", line)
            if mm and cur:
                info.setdefault(cur, {})["regs"] = int(mm.group(1))
        report[name] = info
    return report


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--check" in sys.argv:
        import json
        print(json.dumps(check_kernels(), indent=1))
