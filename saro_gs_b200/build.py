"""In-tree build of libsaro_gs_b200.so (hand-written sm_100a CUDA + C ABI).

Plain nvcc, no torch headers: the boundary is a C ABI (include/saro_gs_b200.h).
The .so lands next to this file so it travels with the repo snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsaro_gs_b200.so")

SOURCES = [
    "sgs_api.cu",
    "sgs_preprocess.cu",
    "sgs_binning.cu",
    "sgs_render_fwd.cu",
    "sgs_render_bwd.cu",
    "sgs_preprocess_bwd.cu",
    "sgs_loss.cu",
    "sgs_deform.cu",
    "sgs_densify.cu",
    "sgs_plane.cu",
]
# every header a translation unit may include feeds the per-object rebuild digest (a stale object after a header edit
# would let two kernels disagree on a shared record layout)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
    [os.path.join("..", "..", "include", "saro_gs_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # NOTE: deliberately no --use_fast_math / -fmad=false: the reference build uses nvcc
    # defaults and tile counts must be bit-identical (see DESIGN.md "bit-exact preprocess").
]


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def lib_path():
    return LIB


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src + ".o")
        stamp = op + ".sha"
        dig = _digest([sp] + hdrs)
        objs.append(op)
        fresh = os.path.exists(op) and os.path.exists(stamp) and open(stamp).read() == dig
        if force or not fresh:
            jobs.append((sp, op, stamp, dig))

    def one(job):
        sp, op, stamp, dig = job
        cmd = ["nvcc", "-c"] + NVCC_FLAGS + ["-I", os.path.join(HERE, "..", "include"), sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(dig)
        return r.stderr

    if jobs:
        if verbose:
            print(f"[saro_gs_b200.build] compiling {len(jobs)} unit(s) for sm_100a ...", flush=True)
        with ThreadPoolExecutor(max_workers=min(6, len(jobs))) as ex:
            for msg in ex.map(one, jobs):
                if verbose and msg.strip():
                    print(msg, file=sys.stderr)
    if jobs or not os.path.exists(LIB):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[saro_gs_b200.build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
