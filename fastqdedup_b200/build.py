#!/usr/bin/env python3
"""In-tree build of the native parts of fastqdedup_b200.

* ``libfqd_b200.so``  -- the C-ABI library (include/fqd_b200.h): CUDA kernels for sm_100a
  plus the host launch plan, compiled with nvcc (cross-compiles without a GPU).
* ``_trie / _distance / _fastq`` CPython extension modules -- thin shims over the library
  with the exact Python surface of the reference's extensions.

Everything lands next to this file so the built objects travel with the source tree
(they are git-ignored).  ``python -m fastqdedup_b200.build`` or ``build_all()``.
"""
import concurrent.futures
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libfqd_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
N_GROUPS = 7   # instance groups of csrc/instances.h: pipeline_inst.cu is compiled once per group, in parallel
# (source, object stem, extra flags); the heavy groups first so the pool finishes early
CU_SOURCES = [("pipeline_inst.cu", f"pipeline_g{g}", [f"-DFQD_GROUP={g}"]) for g in (2, 6, 4, 0, 1, 3, 5)] + \
             [(s, os.path.splitext(s)[0], []) for s in ("api.cu", "pipeline.cu", "trie_shim.cu", "nccl_exchange.cu",
                                                        "microbench.cu")]
HEADERS = ["key.cuh", "pipeline.cuh", "partitioned.cuh", "pipeline_impl.cuh", "instances.h", "common.h", "exchange.h",
           "phred_lut.h", os.path.join("..", "..", "include", "fqd_b200.h")]
PY_MODULES = {"_trie": "py_trie.c", "_distance": "py_distance.c", "_fastq": "py_fastq.c"}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _nvcc(item):
    src, stem, extra = item
    obj = os.path.join(OBJ, stem + ".o")
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    if not _stale(obj, deps):
        return obj, ""
    # development shortcut, never used by build_all() callers that ship: FQD_BUILD_GROUPS=0,3 recompiles only
    # those instance groups and links the existing objects of the others (valid while the cross-TU structs
    # of common.h are unchanged)
    only = os.environ.get("FQD_BUILD_GROUPS")
    if only and stem.startswith("pipeline_g") and os.path.exists(obj) and stem[len("pipeline_g"):] not in only.split(","):
        return obj, ""
    cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".ptxas.log", "wt") as fh:
        fh.write(r.stderr)
    return obj, r.stderr


def _gxx(src):
    """Host-only C++ (the native FASTQ passes): plain g++."""
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    srcp = os.path.join(CSRC, src)
    if _stale(obj, [srcp, os.path.join(CSRC, "..", "..", "include", "fqd_b200.h")]):
        subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-Wall", "-Wno-stringop-overflow", "-c", srcp, "-o", obj], check=True)
    return obj


def build_library(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    sources = [s for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s[0]))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 4)) as ex:
        objs = [o for o, _ in ex.map(_nvcc, sources)]
    objs.append(_gxx("fastq_native.cpp"))
    objs.append(_gxx("host_pack.cpp"))
    if _stale(LIB, objs):
        # static cudart (nvcc's default), deliberately: the library must not depend on WHICH libcudart.so.12
        # the host process has mapped (a torch process brings the 12.8 runtime, this toolkit is 12.9)
        cmd = ["nvcc", "-shared", "-o", LIB + ".tmp"] + objs + ["-lz", "-lpthread"]
        subprocess.run(cmd, check=True)
        os.replace(LIB + ".tmp", LIB)   # never a half-written library in the tree
    if verbose:
        print("built", LIB)
    return LIB


def build_python_modules(verbose=False):
    include = sysconfig.get_paths()["include"]
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    outs = []
    for mod, src in PY_MODULES.items():
        srcp = os.path.join(CSRC, src)
        if not os.path.exists(srcp):
            continue
        out = os.path.join(HERE, mod + suffix)
        if _stale(out, [srcp, os.path.join(CSRC, "py_common.h"), LIB]):
            cmd = ["gcc", "-O2", "-fPIC", "-shared", "-Wall", f"-I{include}",
                   f"-I{os.path.join(HERE, '..', 'include')}", srcp, "-o", out + ".tmp",
                   f"-L{HERE}", "-lfqd_b200", "-Wl,-rpath,$ORIGIN"]
            subprocess.run(cmd, check=True)
            os.replace(out + ".tmp", out)
        outs.append(out)
        if verbose:
            print("built", out)
    return outs


def build_all(verbose=False):
    build_library(verbose)
    build_python_modules(verbose)


if __name__ == "__main__":
    build_all(verbose=True)
    sys.exit(0)
