#!/usr/bin/env python
"""Build the UNMODIFIED reference (singul4ri7y/soket) into ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY.  ``oracle/_ref`` is the reference's own CPU/NumPy
implementation, compiled from the sources where they lie under
``/root/reference``; it is used (a) to pin the NumPy restatement in
``oracle/soket_np.py``, (b) to generate ``tests/golden/*.npz`` and (c) as the
``--impl reference`` arm / ``cpu_baseline`` of ``bench.py``.  Nothing under
``soket_b200/`` imports it.

Recipe:
  * copy the tree to a scratch dir (``/root/reference`` is read-only),
  * add an empty ``cupy.pxd`` on the Cython include path -- the reference
    ``cimport``s cupy at ``soket/backend/device.pyx:7``, ``soket/tensor/
    tensor.pyx:19`` and ``soket/tensor/ops/intern.pyx:3`` but never uses a
    C-level CuPy symbol,
  * apply the Cython-3.3 ``__dict__`` compatibility edit to
    ``soket/nn/module.pyx`` (lines 28, 60, 87, 106) -- build-only, no
    numerical code is touched,
  * cythonize with the directives of the reference's ``setup.py:65-70`` and
    build in place,
  * install ONLY the built extension modules and the package's ``.py`` files
    into ``oracle/_ref/soket`` (git-ignored; travels to the GPU box).

No reference source is copied into the tracked repository.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SOKET_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

_SETUP = r'''
import os, numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
exts = []
for root, _, files in os.walk("soket"):
    for f in files:
        if f.endswith(".pyx"):
            p = os.path.join(root, f)
            exts.append(Extension(os.path.splitext(p.replace(os.sep, "."))[0], [p],
                                  include_dirs=[numpy.get_include()],
                                  extra_compile_args=["-O3", "-funroll-loops", "-fno-strict-aliasing", "-w"],
                                  define_macros=[("CYTHON_LIMITED_API", "0")]))
setup(name="soket", ext_modules=cythonize(exts, include_path=["_stubs"], nthreads=8,
      compiler_directives=dict(language_level="3", boundscheck=False, wraparound=False,
                               initializedcheck=False, cdivision=True)), script_args=["build_ext", "--inplace", "-j8"])
'''


def _patch_module_pyx(path: str) -> None:
    src = open(path).read()
    # Cython 3.3 rejects ``x.__dict__`` on a cdef-class typed variable.
    src = src.replace("module.__dict__", "(<object> module).__dict__")
    src = re.sub(r"^\s*self\.__dict__ = PyDict_New\(\)\s*$", "", src, flags=re.M)
    open(path, "w").write(src)


def build(force: bool = False) -> str | None:
    """Returns the path of the built package dir, or None if unavailable."""
    marker = os.path.join(OUT, "soket", "__init__.py")
    if os.path.exists(marker) and not force:
        return OUT
    if not os.path.isdir(os.path.join(REF_SRC, "soket")):
        return None
    tmp = tempfile.mkdtemp(prefix="soket_ref_build_")
    try:
        work = os.path.join(tmp, "src")
        shutil.copytree(REF_SRC, work)
        os.makedirs(os.path.join(work, "_stubs"))
        open(os.path.join(work, "_stubs", "cupy.pxd"), "w").close()
        _patch_module_pyx(os.path.join(work, "soket", "nn", "module.pyx"))
        open(os.path.join(work, "_build.py"), "w").write(_SETUP)
        subprocess.run([sys.executable, "_build.py"], cwd=work, check=True,
                       stdout=subprocess.DEVNULL)
        # install: extension modules + python files only
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        for root, _, files in os.walk(os.path.join(work, "soket")):
            rel = os.path.relpath(root, work)
            for f in files:
                if f.endswith(".so") or f.endswith(".py"):
                    os.makedirs(os.path.join(OUT, rel), exist_ok=True)
                    shutil.copy2(os.path.join(root, f), os.path.join(OUT, rel, f))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("oracle/_ref:", p)
