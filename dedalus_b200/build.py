"""Build the CUDA extension in-tree:  dedalus_b200/libdedalus_b200.so  (nvcc, sm_100a, -lineinfo).

    python -m dedalus_b200.build            # product library (needs nvcc; cross-compiles without a GPU)
    python -m dedalus_b200.build --emu      # TEST-ONLY CPU emulation of the same kernels (tests/emu/)
"""
import os, subprocess, sys, pathlib, shutil

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
SOURCES = ["core.cu", "fft.cu", "rfft_regs.cu", "pencil.cu", "pointwise.cu"]
LIB = PKG / "libdedalus_b200.so"
EMU_DIR = ROOT / "tests" / "emu"
EMU_LIB = EMU_DIR / "libdedalus_b200_emu.so"


def _newer(target, deps):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(pathlib.Path(d).stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [str(CSRC / s) for s in SOURCES]
    deps = srcs + [str(CSRC / "db_common.cuh"), str(CSRC / "tw96.inc"), str(ROOT / "include" / "dedalus_b200.h")]
    if not force and not _newer(LIB, deps):
        return LIB
    if not pathlib.Path(nvcc).exists():
        raise RuntimeError("nvcc not found: cannot build libdedalus_b200.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", str(LIB)] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB


def build_emu(force=False):
    srcs = [str(CSRC / s) for s in SOURCES] + [str(EMU_DIR / "cuda_emu.cpp")]
    deps = srcs + [str(CSRC / "db_common.cuh"), str(ROOT / "include" / "dedalus_b200.h"), str(EMU_DIR / "cuda_emu.h")]
    if not force and not _newer(EMU_LIB, deps):
        return EMU_LIB
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DDB_EMU", "-I", str(EMU_DIR), "-x", "c++"] + srcs + ["-o", str(EMU_LIB)]
    subprocess.run(cmd, check=True)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force=True))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
