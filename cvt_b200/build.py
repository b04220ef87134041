"""In-tree build of the CUDA library and the C++ tools: plain nvcc for sm_100a, no JIT cache.

    python -m cvt_b200.build            # incremental
    python -m cvt_b200.build --force

Outputs (git-ignored, but they travel with the gpurun snapshot):
    cvt_b200/lib/libb200nn.so      -- the C-ABI shared library declared in include/b200nn.h
    tools/bin/*                    -- C++ programs written against the drop-in headers
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libb200nn.so")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"]


def _newer(src_paths, out) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_paths)


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer([s] + headers, o):
            jobs.append(([NVCC] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", s, "-o", o], o + ".log"))
    if jobs:
        if verbose:
            print(f"[build] compiling {len(jobs)} CUDA translation unit(s) for sm_100a")
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    if force or jobs or not os.path.exists(LIB):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"], os.path.join(OBJDIR, "link.log"))
    return LIB


def build_tools(force: bool = False, verbose: bool = False):
    """C++ programs written against include/b200nn/*.hpp (the reference-facing classes)."""
    tdir = os.path.join(ROOT, "tools")
    bdir = os.path.join(tdir, "bin")
    os.makedirs(bdir, exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    outs = []
    hdrs = glob.glob(os.path.join(ROOT, "include", "**", "*"), recursive=True)
    hdrs = [h for h in hdrs if os.path.isfile(h)]
    for s in sorted(glob.glob(os.path.join(tdir, "*.cpp"))):
        o = os.path.join(bdir, os.path.basename(s)[:-4])
        outs.append(o)
        if force or _newer([s, LIB] + hdrs, o):
            _run([gxx, "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), s, "-o", o, "-L", LIBDIR, "-lb200nn",
                  "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../../cvt_b200/lib"], o + ".log")
    # standalone CUDA measurement programs (tools/*.cu), e.g. the bare int8 MMA peak
    for s_ in sorted(glob.glob(os.path.join(tdir, "*.cu"))):
        o = os.path.join(bdir, os.path.basename(s_)[:-3])
        outs.append(o)
        if force or _newer([s_], o):
            _run([NVCC] + ARCH + ["-O3", "-lineinfo", "-o", o, s_], o + ".log")
    # the reference's OWN mains, unmodified, compiled in place against the forwarding headers and linked with
    # libb200nn: the drop-in proof (only where /root/reference exists; the binaries then travel with the snapshot).
    # A quoted #include resolves next to the including file FIRST, so compiling the reference file by path would
    # silently pick up the reference's own CPU headers sitting beside it; the source is therefore piped to the
    # compiler on stdin (the "current directory" of the translation unit is then a neutral one) and the result is
    # checked to import the C ABI.
    compat = os.path.join(ROOT, "include", "b200nn", "compat")
    for ref_src, name, std, need in (
            ("/root/reference/brute_force_search/src/brute_force.cpp", "ref_brute_force_on_b200nn", "c++11", "b200nn_flat_search"),
            ("/root/reference/opq/train_codebook/train_PQ.cpp", "ref_train_PQ_on_b200nn", "c++11", "b200nn_kmeans"),
            # c++17: the test copy-initialises an Int8Quan from a temporary and the drop-in class is non-copyable
            ("/root/reference/scalar_quantization/scalar_quantization/int8_quan_test.cpp", "ref_int8_quan_test_on_b200nn", "c++17",
             "b200nn_sq_encode")):
        o = os.path.join(bdir, name)
        if os.path.exists(ref_src) and (force or _newer([ref_src, LIB] + hdrs, o)):
            with open(ref_src, "rb") as src:
                r = subprocess.run([gxx, "-O2", "-std=" + std, "-fno-operator-names", "-I", compat, "-x", "c++", "-", "-o", o, "-L", LIBDIR,
                                    "-lb200nn", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../../cvt_b200/lib"],
                                   stdin=src, cwd=OBJDIR, capture_output=True, text=True)
            open(o + ".log", "w").write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"build failed: {name} from {ref_src}")
            if need.encode() not in open(o, "rb").read():
                os.remove(o)
                raise RuntimeError(f"{name} does not import {need}: it was not compiled against the drop-in headers")
        if os.path.exists(o):
            outs.append(o)
    # hnsw_sifts_retrieval/makeSearch.cpp + siftsIndex.cpp (the other caller north_star names) on the drop-in headers:
    # -DB200NN_HNSW_DROP_IN makes hnswlib::HierarchicalNSW<float> (siftsIndex.hpp:49, siftsIndex.cpp:6) the exact GPU index
    # that reads the HNSW index file.  OpenCV is absent from this image: tests/stubs/opencv2 stands in for it (float matrices;
    # the "SIFT detector" reads descriptors the real cv2 SIFT produced).  The only edit of the sources: the hard-coded
    # /Users/willard/... path prefixes become the relative data/ (sed) -- the same edit the CPU golden build of the tests makes.
    ref_dir = "/root/reference/hnsw_sifts_retrieval"
    o = os.path.join(bdir, "ref_makeSearch_on_b200nn")
    srcs = [os.path.join(ref_dir, "makeSearch.cpp"), os.path.join(ref_dir, "siftsIndex.cpp"), os.path.join(ref_dir, "siftsIndex.hpp")]
    stubs = os.path.join(ROOT, "tests", "stubs")
    if all(os.path.exists(s) for s in srcs) and (force or _newer(srcs + [LIB] + hdrs + glob.glob(os.path.join(stubs, "opencv2", "*.hpp")), o)):
        txt = open(srcs[0], "rb").read()
        for pre in (b"/Users/willard/codes/cpp/hnsw_sift_retrieval/hnsw_sifts_retrieval/data/", b"/Users/willard/projects/bovw/data/"):
            txt = txt.replace(pre, b"data/")
        common = [gxx, "-O2", "-std=c++11", "-DB200NN_HNSW_DROP_IN", "-I", compat, "-I", ref_dir, "-I", stubs]
        o1, o2 = os.path.join(OBJDIR, "makeSearch_dropin.o"), os.path.join(OBJDIR, "siftsIndex_dropin.o")
        r = subprocess.run(common + ["-x", "c++", "-", "-c", "-o", o1], input=txt, cwd=OBJDIR, capture_output=True)
        log = r.stdout + r.stderr
        if r.returncode == 0:
            r = subprocess.run(common + ["-c", srcs[1], "-o", o2], cwd=OBJDIR, capture_output=True)
            log += r.stdout + r.stderr
        if r.returncode == 0:
            r = subprocess.run([gxx, o1, o2, "-o", o, "-L", LIBDIR, "-lb200nn", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../../cvt_b200/lib"],
                               capture_output=True)
            log += r.stdout + r.stderr
        open(o + ".log", "wb").write(log)
        if r.returncode != 0:
            sys.stderr.write(log.decode(errors="replace"))
            raise RuntimeError("build failed: ref_makeSearch_on_b200nn")
        if b"b200nn_flat_load_hnsw" not in open(o, "rb").read():
            os.remove(o)
            raise RuntimeError("ref_makeSearch_on_b200nn was not compiled against the drop-in headers")
    if os.path.exists(o):
        outs.append(o)
    return outs


def build_all(force: bool = False, verbose: bool = False):
    lib = build_lib(force, verbose)
    tools = build_tools(force, verbose)
    return lib, tools


if __name__ == "__main__":
    lib, tools = build_all(force="--force" in sys.argv, verbose=True)
    print(lib)
    for t in tools:
        print(t)
