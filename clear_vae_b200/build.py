"""In-tree build of the native code (sm_100a only).

  clear_vae_b200/_lib/libclearvae_b200.so  — CUDA kernels + C ABI (include/clearvae_b200.h), no torch
  clear_vae_b200/_C.so                     — torch custom-op shim (TORCH_LIBRARY clearvae) over the C ABI

Both are built with explicit nvcc / g++ commands (no JIT cache) so the `.so`
files travel with the repo snapshot to the GPU box.  Rebuilds are skipped when
the outputs are newer than every source.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libclearvae_b200.so")
EXT = os.path.join(PKG, "_C.so")
OBJ_DIR = os.path.join(PKG, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", f"-I{INC}", f"-I{CSRC}"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(out, srcs):
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(s) <= t for s in srcs)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    return r.stdout


def build_lib(verbose=False, force=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    cus = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INC, "*.h"))
    objs = []
    procs = []
    for cu in cus:
        obj = os.path.join(OBJ_DIR, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        if not force and _newer(obj, [cu] + hdrs):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", cu, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    if force or procs or not _newer(LIB, objs):
        _run([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"])
    return LIB


def build_ext(verbose=False, force=False):
    src = os.path.join(CSRC, "torch_binding.cpp")
    hdrs = glob.glob(os.path.join(INC, "*.h"))
    if not force and _newer(EXT, [src] + hdrs):
        return EXT
    import torch
    from torch.utils import cpp_extension as ce

    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_C",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", f"-I{INC}"]
    cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_home}/include"]
    cmd += [src, "-o", EXT, f"-L{LIB_DIR}", "-lclearvae_b200", f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
            "-ltorch_cuda", "-ltorch", f"-L{cuda_home}/lib64", "-lcudart",
            "-Wl,-rpath,$ORIGIN/_lib", f"-Wl,-rpath,{tlib}"]
    if verbose:
        print(" ".join(cmd), flush=True)
    _run(cmd)
    return EXT


def build_tools(verbose=False, force=False):
    """tools/mufu_probe: the ex2 throughput probe bench.py runs for the latent-loss roofline denominator."""
    src = os.path.join(ROOT, "tools", "mufu_probe.cu")
    out = os.path.join(ROOT, "tools", "mufu_probe")
    if not os.path.exists(src) or (not force and _newer(out, [src])):
        return out
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", src, "-o", out]
    if verbose:
        print(" ".join(cmd), flush=True)
    _run(cmd)
    return out


def build_all(verbose=False, force=False):
    build_lib(verbose, force)
    build_ext(verbose, force)
    build_tools(verbose, force)
    return LIB, EXT


if __name__ == "__main__":
    build_all(verbose=True, force="--force" in sys.argv)
    print("built", LIB, EXT)
