"""Build script: libsoketb200.so (nvcc, sm_100a) + the Cython host extension.

    python soket_b200/build.py [--force] [--verbose]

Everything is built IN-TREE (``soket_b200/lib/libsoketb200.so`` and
``soket_b200/_core.*.so``) so the artefacts travel with the repository
snapshot to the GPU box; nothing is cached outside the tree.  nvcc
cross-compiles for sm_100a without a GPU present.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build", "obj")
LIB = os.path.join(LIBDIR, "libsoketb200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # the CPU oracle runs with FTZ|DAZ set (soket/utils/ftz.pyx:18-24); IEEE div/sqrt
    "-ftz=true", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]

CU_SOURCES = [
    "runtime.cu", "ewise.cu", "reduce.cu", "index.cu", "rng.cu",
    "matmul.cu", "matmul_simt.cu", "matmul_tc.cu", "matmul_tc2.cu", "matmul_split.cu", "nn_fused.cu", "optim.cu", "dp.cu", "dp_p2p.cu",
]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "soket_b200.h"))
    return hs


def _compile_one(name, force, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJDIR, name.replace(".cu", ".o"))
    if not force and not _newer([src] + _headers(), obj):
        return obj, ""
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda n: _compile_one(n, force, verbose), CU_SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for (_, log), n in zip(results, CU_SOURCES):
            if log:
                print(f"--- {n}\n{log}")
    if force or _newer(objs, LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["--cudart", "static", "-ldl", "-lpthread", "-lrt",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_cython(force=False, verbose=False):
    """Cythonize + compile soket_b200/*.pyx against libsoketb200.so."""
    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution

    pyx = sorted(f for f in os.listdir(HERE) if f.endswith(".pyx"))
    if not pyx:
        return []
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pxds = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".pxd", ".pxi"))]
    todo = []
    for f in pyx:
        target = os.path.join(HERE, f[:-4] + suffix)
        if force or _newer([os.path.join(HERE, f), os.path.join(INCLUDE, "soket_b200.h")] + pxds, target):
            todo.append(f)
    if not todo:
        return []
    exts = [
        Extension(
            "soket_b200." + f[:-4], [os.path.join(HERE, f)],
            include_dirs=[INCLUDE, numpy.get_include()],
            library_dirs=[LIBDIR], libraries=["soketb200"],
            runtime_library_dirs=["$ORIGIN/lib"],
            extra_compile_args=["-O2", "-w"],
            define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
        )
        for f in todo
    ]
    cwd = os.getcwd()
    os.chdir(os.path.dirname(HERE))
    try:
        ext_modules = cythonize(
            exts, quiet=not verbose, include_path=[HERE],
            # boundscheck / wraparound are set per file in the `# cython:` headers
            compiler_directives=dict(language_level="3", initializedcheck=False, cdivision=True),
            build_dir=os.path.join(HERE, "build", "cython"),
        )
        dist = Distribution({"name": "soket_b200", "ext_modules": ext_modules})
        cmd = dist.get_command_obj("build_ext")
        cmd.inplace = True
        cmd.build_temp = os.path.join(HERE, "build", "temp")
        cmd.parallel = min(8, os.cpu_count() or 4)
        if not verbose:
            cmd.verbose = 0
        cmd.ensure_finalized()
        cmd.run()
    finally:
        os.chdir(cwd)
    return todo


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    built = build_cython(force, verbose)
    return lib, built


if __name__ == "__main__":
    force = "--force" in sys.argv
    verbose = "--verbose" in sys.argv
    lib, built = build_all(force, verbose)
    print("built", lib, built)
