"""The kernel emulation schedules the threads of a block cooperatively, lowest thread first.  Code that is correct only because of
that order -- a missing __syncthreads / __syncwarp between a producer and a consumer -- would still pass; DB_EMU_ORDER=reverse runs
the threads in descending order instead, so such code fails in one of the two orders.  This re-runs the tests of the kernels that have
not been on hardware (dense shell systems, pair combinations with radial symbols, matsolver plugins) and of the banded kernels in
reverse order; the whole CPU suite passes that way too (run it with the variable set)."""
import os, subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_kernels_pass_with_reversed_thread_order():
    env = dict(os.environ, DB_EMU_ORDER="reverse")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
           str(ROOT / "tests" / "test_emu_shell.py"), str(ROOT / "tests" / "test_emu_sphere.py"), str(ROOT / "tests" / "test_emu_plugins.py"),
           "-k", "a_sbdf2 or banded or sw16 or matsolver or transforms"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, env=env, cwd=str(ROOT))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout


def test_alternative_kernel_variants():
    """The variants behind environment switches -- the dense solve with its columns in global memory (taken for n > 800), the banded
    kernels with the cp.async ring -- stay correct."""
    for env_extra, select in ((dict(DB_DENSE_SOLVE_SMEM="0"), "dense or a_sbdf2"), (dict(DB_BANDED_MODE="0"), "banded or sw16")):
        cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
               str(ROOT / "tests" / "test_emu_shell.py"), str(ROOT / "tests" / "test_emu_sphere.py"), "-k", select]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, env=dict(os.environ, **env_extra), cwd=str(ROOT))
        assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
