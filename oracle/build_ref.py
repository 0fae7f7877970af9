#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds the *unmodified* reference rasterizer into oracle/_ref/.

The reference hot path (submodules/gaussian_rasterization_ch3 of yjb6/SaRO-GS) is CUDA
code that compiles from five of its own source files.  This recipe compiles those files
WHERE THEY LIE under /root/reference (nothing is copied into the repo) and writes only a
binary: oracle/_ref/_C.cpython-*.so, the reference's own pybind module ``_C``
(ext.cpp:15-19) with ``rasterize_gaussians``, ``rasterize_gaussians_backward`` and
``mark_visible``.

Flags mirror the stock ``setup.py`` build (setup.py:19-29: no -O/-fmad/-use_fast_math
flags beyond torch's defaults, glm on the include path).  Two deviations, both needed
on this toolchain and neither touching arithmetic:
  * ``--pre-include cstdint``  (rasterizer_impl.h:24,40-61 use std::uintptr_t/uint32_t
    without including <cstdint>; gcc 13 no longer leaks it transitively)
  * ``-gencode arch=compute_100a,code=sm_100a`` (the reference ships no sm_100 arch).

oracle/_ref/ is git-ignored but NOT gpurun-ignored: the .so travels to the GPU box, where
/root/reference does not exist.  Only tests/, __graft_entry__.smoke() and bench.py's
reference/cpu_baseline legs may load it.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SARO_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "submodules", "gaussian_rasterization_ch3")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")


def ref_so_path():
    return os.path.join(OUT, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=True):
    so = ref_so_path()
    if os.path.exists(so) and not force:
        return so
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[build_ref] {SRC} absent: cannot build reference (prebuilt .so expected)")
        return None
    from torch.utils import cpp_extension as ce
    os.makedirs(OBJ, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"],
                                         os.path.join(SRC, "third_party", "glm"), SRC]:
        inc += ["-I", p]
    common = ["-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1", "-std=c++17"]
    nvcc = ["nvcc", "-c", "--pre-include", "cstdint",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "--compiler-options", "-fPIC"] + ce.COMMON_NVCC_FLAGS + common + inc
    gxx = ["g++", "-c", "-fPIC", "-O2", "-include", "cstdint"] + common + inc
    units = [
        (nvcc, "cuda_rasterizer/rasterizer_impl.cu"),
        (nvcc, "cuda_rasterizer/forward.cu"),
        (nvcc, "cuda_rasterizer/backward.cu"),
        (nvcc, "rasterize_points.cu"),
        (gxx, "ext.cpp"),
    ]

    def one(u):
        cmd, rel = u
        o = os.path.join(OBJ, os.path.basename(rel) + ".o")
        full = cmd + [os.path.join(SRC, rel), "-o", o]
        r = subprocess.run(full, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"[build_ref] failed: {' '.join(full)}\n{r.stderr[-4000:]}")
        return o

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(one, units))
    libs = []
    for p in ce.library_paths("cuda"):
        libs += ["-L", p, f"-Wl,-rpath,{p}"]
    link = ["g++", "-shared", "-o", so] + objs + libs + \
           ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"[build_ref] link failed:\n{r.stderr[-4000:]}")
    if verbose:
        print(f"[build_ref] built {so}")
    return so


if __name__ == "__main__":
    build(force="--force" in sys.argv)
