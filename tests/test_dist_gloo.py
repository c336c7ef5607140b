"""N>1 path on CPU: world_size-2 gloo run of the distributed solver (tests/dist_worker.py)."""
import subprocess, sys, pathlib, os, socket, pytest
ROOT = pathlib.Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("which", ["rb3d_8.npz", "rb2d_16x16.npz", "rb3d_16.npz", "blocked:16x32", "sphere:sw16", "shell:a_sbdf2", "tasks:cartesian", "tasks:shell", "plugin:transpose", "strong3d", "shell:forced", "staged2d"])
def test_two_rank_solver_matches_reference(which):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_worker.py"), which]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
