"""Build recipes (explicit nvcc / g++ invocations, outputs in-tree so they travel with gpurun).

  build_host()   -> gpu-raytracer_b200/host/libptb_host.so   C++ CPU BVH builder (SAH + CWBVH)
  build_cuda()   -> gpu-raytracer_b200/csrc/libptb.so        sm_100a kernels + the C ABI of include/ptb.h
  build_facade() -> gpu-raytracer_b200/host/libptb_pathtracer.so  compiled C++ facade (Integrator / Pathtracer entry points) + tests/cpp/facade_render
  build_oracle() -> oracle/libpt_oracle.so                    CPU restatement (test infrastructure only)
  build_ref()    -> oracle/_ref/*                             reference kernels compiled from /root/reference (if present)
"""
from __future__ import annotations

import glob
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout)
    return r.stdout


def build_host(force=False):
    src = os.path.join(PKG_DIR, "host", "bvh_build.cpp")
    out = os.path.join(PKG_DIR, "host", "libptb_host.so")
    if force or _stale(out, [src, os.path.join(PKG_DIR, "host", "static_merge.h")]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-o", out, src])
    return out


def cuda_sources():
    d = os.path.join(PKG_DIR, "csrc")
    return sorted(glob.glob(os.path.join(d, "*.cu"))), sorted(glob.glob(os.path.join(d, "*.cuh")) + glob.glob(os.path.join(REPO_ROOT, "include", "*.h")))


def build_cuda(force=False, verbose=False, defines=(), suffix=""):
    """defines/suffix build tuning variants (libptb<suffix>.so) for A/B measurements; the product is the plain libptb.so."""
    srcs, hdrs = cuda_sources()
    out = os.path.join(PKG_DIR, "csrc", f"libptb{suffix}.so")
    if force or _stale(out, srcs + hdrs + [os.path.join(PKG_DIR, "host", "bvh_build.cpp"), os.path.join(PKG_DIR, "host", "static_merge.h")]):
        cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC", "-shared",
               *[f"-D{d}" for d in defines],
               "-Xcompiler", "-ffp-contract=off",
               "-I", os.path.join(REPO_ROOT, "include"), "-I", os.path.join(PKG_DIR, "csrc"), "-o", out, *srcs,
               os.path.join(PKG_DIR, "host", "bvh_build.cpp")]          # the CPU SAH/CWBVH builder, also used inside libptb (static merge)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        log = _run(cmd)
        if verbose:
            print(log)
    return out


def build_facade(force=False):
    """host/libptb_pathtracer.so: the compiled C++ facade (Integrator / Pathtracer entry points over the C ABI), and the small
    test driver tests/cpp/facade_render that renders through it."""
    src = os.path.join(PKG_DIR, "host", "ptb_pathtracer.cpp")
    hdr = os.path.join(PKG_DIR, "host", "ptb_pathtracer.h")
    out = os.path.join(PKG_DIR, "host", "libptb_pathtracer.so")
    libdir = os.path.join(PKG_DIR, "csrc")
    inc = ["-I", os.path.join(REPO_ROOT, "include"), "-I", os.path.join(PKG_DIR, "host")]
    if force or _stale(out, [src, hdr, os.path.join(REPO_ROOT, "include", "ptb.h")]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", *inc, "-o", out, src, "-L", libdir, "-lptb", "-Wl,-rpath,$ORIGIN/../csrc"])
    drv_src = os.path.join(REPO_ROOT, "tests", "cpp", "facade_render.cpp")
    drv = os.path.join(REPO_ROOT, "tests", "cpp", "facade_render")
    if force or _stale(drv, [drv_src, hdr, out]):
        _run(["g++", "-O2", "-std=c++17", *inc, "-o", drv, drv_src, "-L", os.path.join(PKG_DIR, "host"), "-lptb_pathtracer", "-L", libdir, "-lptb",
              "-Wl,-rpath,$ORIGIN/../../gpu-raytracer_b200/host", "-Wl,-rpath,$ORIGIN/../../gpu-raytracer_b200/csrc", "-Wl,-rpath-link," + libdir])
    return out


def build_oracle(force=False):
    src = os.path.join(REPO_ROOT, "oracle", "pt_oracle.c")
    out = os.path.join(REPO_ROOT, "oracle", "libpt_oracle.so")
    if force or _stale(out, [src]):
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", out, src, "-lm"])
    return out


def build_ref(force=False):
    """Compile the reference's own kernels where they lie (needs /root/reference); no-op when absent."""
    mk = os.path.join(REPO_ROOT, "oracle", "Makefile")
    if not os.path.isdir("/root/reference/Src/CUDA") or not os.path.exists(mk):
        return None
    _run(["make", "-C", os.path.join(REPO_ROOT, "oracle"), "ref"] + (["-B"] if force else []))
    return os.path.join(REPO_ROOT, "oracle", "_ref")
