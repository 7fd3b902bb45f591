"""Build recipe of liborb_b200.so (sm_100a only, in-tree so the .so travels with gpurun snapshots)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liborb_b200.so")
SOURCES = ["orb_api.cu", "orb_ingest.cu", "orb_lk.cu", "orb_png.cpp"]
DEPS = ["orb_api.cu", "orb_ingest.cu", "orb_lk.cu", "orb_internal.h", "orb_png.cpp", "orb_png.h", "orb_kernels.cuh", "orb_match_tc.cuh", "orb_ingest_kernels.cuh", "orb_lk_kernels.cuh", "orb_math.cuh", "orb_plan.h",
        os.path.join("..", "..", "include", "orb_b200.h"), os.path.join("..", "..", "include", "orb_brief_pattern.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other targets, no fallback arch
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # the float paths restate IEEE sequences of the reference (no FMA contraction)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default",
    "-shared",
]


def nvcc_path():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS) or os.path.getmtime(__file__) > t


def build(force=False, verbose=False, extra=(), out=None):
    """Compile liborb_b200.so in-tree (or a variant build, e.g. -DORB_BOUNDS_CHECK, to `out`)."""
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    extra = list(extra) + os.environ.get("ORB_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                out=sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None))
