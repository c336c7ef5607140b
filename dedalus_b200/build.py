"""Build the CUDA extension in-tree:  dedalus_b200/libdedalus_b200.so  (nvcc, sm_100a, -lineinfo).

    python -m dedalus_b200.build            # product library (needs nvcc; cross-compiles without a GPU)
    python -m dedalus_b200.build --emu      # TEST-ONLY CPU emulation of the same kernels (tests/emu/)

Every source is compiled to its own object (dedalus_b200/csrc/_obj/, git-ignored) and only stale objects are rebuilt;
the objects are compiled concurrently and linked into the shared library through a temporary file + rename, so that
concurrent test workers never load a half-written library.
"""
import os, subprocess, sys, pathlib, shutil

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
SOURCES = ["core.cu", "fft.cu", "rfft_regs.cu", "pencil.cu", "pointwise.cu", "banded.cu", "dense.cu"]
LIB = PKG / "libdedalus_b200.so"
OBJ = CSRC / "_obj"
EMU_DIR = ROOT / "tests" / "emu"
EMU_LIB = EMU_DIR / "libdedalus_b200_emu.so"
HEADERS = [CSRC / "db_common.cuh", CSRC / "tw96.inc", ROOT / "include" / "dedalus_b200.h"]


def source_hash():
    """sha256 over every CUDA source and header of the library: the built .so carries it in a side file so that a stale
    library (sources edited after the last build; file times do not survive every copy) is detected when it is loaded."""
    import hashlib
    h = hashlib.sha256()
    for f in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(f.name.encode()); h.update(pathlib.Path(f).read_bytes())
    return h.hexdigest()


def lib_is_current():
    stamp = pathlib.Path(str(LIB) + ".hash")
    return LIB.exists() and stamp.exists() and stamp.read_text().strip() == source_hash()


def _newer(target, deps):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(pathlib.Path(d).stat().st_mtime > t for d in deps)


def _compile_all(jobs):
    """jobs: list of (command, object path); run the stale ones concurrently."""
    procs = [(subprocess.Popen(cmd), obj) for cmd, obj in jobs]
    failed = [str(obj) for p, obj in procs if p.wait() != 0]
    if failed:
        raise RuntimeError(f"compilation failed: {failed}")


def _link(cmd, target):
    tmp = target.with_suffix(f".tmp{os.getpid()}.so")
    subprocess.run(cmd + ["-o", str(tmp)], check=True)
    os.replace(tmp, target)


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    OBJ.mkdir(exist_ok=True)
    jobs, objs = [], []
    for s in SOURCES:
        obj = OBJ / (s + ".o")
        objs.append(obj)
        if force or _newer(obj, [CSRC / s] + HEADERS):
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                   "-Xcompiler", "-fPIC", "-c", str(CSRC / s), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((cmd, obj))
    if not jobs and not _newer(LIB, objs) and lib_is_current():
        return LIB
    if not pathlib.Path(nvcc).exists():
        raise RuntimeError("nvcc not found: cannot build libdedalus_b200.so")
    _compile_all(jobs)
    _link([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + [str(o) for o in objs], LIB)
    pathlib.Path(str(LIB) + ".hash").write_text(source_hash() + "\n")
    return LIB


def build_emu(force=False):
    OBJ.mkdir(exist_ok=True)
    jobs, objs = [], []
    for src in [CSRC / s for s in SOURCES] + [EMU_DIR / "cuda_emu.cpp"]:
        obj = OBJ / (src.name + ".emu.o")
        objs.append(obj)
        if force or _newer(obj, [src, EMU_DIR / "cuda_emu.h"] + HEADERS):
            # no _FORTIFY_SOURCE: its longjmp check rejects the (deliberate) jumps between fiber stacks of tests/emu/cuda_emu.cpp
            jobs.append((["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-DDB_EMU", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-I", str(EMU_DIR), "-x", "c++", "-c", str(src),
                          "-o", str(obj)], obj))
    if not jobs and not _newer(EMU_LIB, objs):
        return EMU_LIB
    _compile_all(jobs)
    _link(["g++", "-shared", "-fPIC"] + [str(o) for o in objs], EMU_LIB)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
