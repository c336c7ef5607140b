"""Sphere (S2) path on the GPU through the C ABI: SphereBasis field transforms, per-m pencil matrices, the banded pencil
kernels and the shallow-water IVP of BASELINE config 4 against data produced by the unmodified reference
(tests/golden/sphere.npz), up to the config's own size (512 x 256, Lmax = 254)."""
import ctypes as C
import numpy as np, pytest
import sphere_cases as S

pytestmark = pytest.mark.gpu


class _CudaArrays:
    @property
    def lib(self):
        from dedalus_b200.lib import get_lib
        return get_lib()

    @property
    def stream(self):
        from dedalus_b200.lib import current_stream
        return current_stream()

    def dev(self, a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def ptr(self, t):
        return C.c_void_p(t.data_ptr())

    def host(self, t):
        return t.cpu().numpy()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sphere_field_transforms(golden, tag):
    S.check_field_transforms(golden("sphere.npz"), tag)


def test_banded_kernels_against_dense_solves():
    S.check_banded_kernels(_CudaArrays())


def test_banded_factor_flags_singular_system():
    S.check_banded_singular(_CudaArrays())


def test_sphere_pencil_matrices(golden):
    S.check_pencil_matrices(golden("sphere.npz"))


@pytest.mark.parametrize("tag,scheme", [("sw16", "RK222"), ("sw32", "RK222"), ("sw32sbdf2", "SBDF2")])
def test_shallow_water_matches_reference(golden, tag, scheme):
    sw, solver = S.check_shallow_water(golden("sphere.npz"), tag, scheme)
    assert solver.bset.last_verify < 1e-12


def test_shallow_water_config4_size(golden):
    """BASELINE config 4 at its stated size: Nphi, Ntheta = 512, 256 (Lmax = 254), 3 RK222 steps."""
    solver = S.check_config4_size(golden("sphere.npz"))
    assert solver.bset.last_verify < 1e-12


@pytest.mark.parametrize("Nphi,Ntheta,scheme", [(64, 32, "RK443"), (128, 64, "RK222"), (64, 32, "SBDF2")])
def test_shallow_water_with_timestep_changes_matches_oracle(Nphi, Ntheta, scheme):
    """Sizes without reference fixtures, other schemes, and time-step changes (each one refactorises all per-m systems), against
    the oracle (pinned to the reference by tests/test_oracle.py)."""
    dt = 600 / 3600 * min(1.0, 128 / Ntheta)
    solver = S.check_against_oracle(Nphi, Ntheta, scheme, [dt, dt, dt / 2, dt / 2, dt])
    assert solver.bset.last_verify < 1e-12
