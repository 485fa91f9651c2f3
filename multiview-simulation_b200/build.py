"""Builds libmvsim.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python multiview-simulation_b200/build.py [--force] [--verbose]

Objects go to build/ (git-ignored), the library to multiview-simulation_b200/libmvsim.so (git-ignored,
but shipped to the GPU box by gpurun).  No torch, no cmake: plain nvcc, translation units in parallel.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# MVSIM_PACKED_FFT=0 builds the butterflies on scalar FADD/FFMA instead of Blackwell's packed FP32x2 pipe (A/B measurements)
PACKED = os.environ.get("MVSIM_PACKED_FFT", "1") != "0"
# MVSIM_PACKED_X=1 builds the x passes on the packed pipe too (A/B measurements: libmvsim_px.so, loaded with MVSIM_LIB=...)
PACKED_X = os.environ.get("MVSIM_PACKED_X", "0") == "1"
OBJ = os.path.join(ROOT, ("build_px" if PACKED_X else "build") if PACKED else "build_scalar")
LIB = os.path.join(PKG, ("libmvsim_px.so" if PACKED_X else "libmvsim.so") if PACKED else "libmvsim_scalar.so")
# generic A/B builds: MVSIM_VARIANT=<name> MVSIM_EXTRA_FLAGS="-DMVSIM_EXP_FOO=1" -> build_<name>/, libmvsim_<name>.so (MVSIM_LIB=... loads it)
VARIANT = os.environ.get("MVSIM_VARIANT", "")
EXTRA = os.environ.get("MVSIM_EXTRA_FLAGS", "").split()
if VARIANT:
    OBJ = os.path.join(ROOT, "build_" + VARIANT)
    LIB = os.path.join(PKG, f"libmvsim_{VARIANT}.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
         f"-DMVSIM_PACKED_FFT={1 if PACKED else 0}", f"-DMVSIM_PACKED_X={1 if PACKED_X else 0}"] + EXTRA

UNITS = [("stages", "stages.cu", []), ("phantom", "phantom.cu", []), ("conv", "conv.cu", []), ("capi", "capi.cu", [])] + \
        [(f"fft_g{g}_t{t}", os.path.join("fft", "fft_group.cu"), [f"-DMVSIM_GROUP={g}", f"-DMVSIM_LANES={t}"])
         for g in (4, 3, 2, 1, 0) for t in (8,)] + \
        [(f"fft_dec_g{g}_t8", os.path.join("fft", "fft_group.cu"), [f"-DMVSIM_GROUP={g}", "-DMVSIM_LANES=8", "-DMVSIM_DEC_UNIT=1"]) for g in (2, 1)] + \
        [(f"fft_poly_g{g}_t8", os.path.join("fft", "fft_group.cu"), [f"-DMVSIM_GROUP={g}", "-DMVSIM_LANES=8", "-DMVSIM_POLY_UNIT=1"]) for g in (2, 1)]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libmvsim.so cannot be built (there is no CPU fallback)")


def _sources_mtime():
    m = os.path.getmtime(os.path.join(ROOT, "include", "mvsim.h"))
    for d, _, files in os.walk(CSRC):
        for f in files:
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    newest = _sources_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    nvcc = _nvcc()
    # the image exports CC/CXX=/opt/gcc/bin/*; let nvcc use the PATH g++ it was validated with
    env = dict(os.environ)

    def compile_unit(u):
        name, src, extra = u
        obj = os.path.join(OBJ, name + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            return obj, ""
        cmd = [nvcc] + ARCH + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
