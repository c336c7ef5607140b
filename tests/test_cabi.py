"""The CUDA library builds for sm_100a on the GPU-less container, loads, and exports every symbol declared in
include/dedalus_b200.h (no compute calls here).  Also: the product refuses to run without a CUDA device."""
import re, pathlib, ctypes, pytest
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_library_builds_loads_and_exports_declared_symbols():
    from dedalus_b200 import build, lib
    so = build.build()
    cdll = ctypes.CDLL(str(so))
    header = (ROOT / "include" / "dedalus_b200.h").read_text()
    declared = set(re.findall(r"\b(db_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(cdll, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.SIGNATURES), "ctypes signature table and header disagree"
    bound = lib.bind(so)
    assert bound.version() == 100
    assert b"" == bound._raw_db_last_error() or isinstance(bound._raw_db_last_error(), bytes)


def test_sass_is_sm100a_only():
    import subprocess, shutil
    from dedalus_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "--list-elf", str(build.build())], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import dedalus_b200 as d3
    from dedalus_b200 import examples
    from dedalus_b200.lib import DedalusB200Error
    pk = examples.kdv_burgers(N=16)
    solver = pk['problem'].build_solver(d3.SBDF1)
    with pytest.raises(DedalusB200Error):
        solver.step(1e-3)
    with pytest.raises(DedalusB200Error):
        pk['u']['g']


def test_product_does_not_import_oracle_or_reference():
    for path in (ROOT / "dedalus_b200").glob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
        assert "/root/reference" not in text, path
