"""Complex-dtype pencil path (T3: complex128 IVPs on ComplexFourier^n x Jacobi; dedalus_b200/complex_path.py and the real / imaginary
plane embedding of dedalus_b200/pencils.py) through the CPU emulation of the kernels; GPU versions: tests/test_gpu_t3_complex.py."""
import numpy as np, pytest
import dedalus_b200 as d3
from emu import emu_lib as E
import complex_cases as CC


@pytest.fixture(autouse=True)
def emulation():
    E.install()
    yield
    E.uninstall()


@pytest.mark.parametrize("timestepper", list(d3.schemes.keys()))
def test_heat_periodic_complex_every_timestepper(timestepper):
    CC.check_heat_periodic(timestepper)


@pytest.mark.parametrize("tag,scheme", [("rk222", "RK222"), ("sbdf2", "SBDF2")])
def test_complex_ginzburg_landau_matches_reference(golden, tag, scheme):
    solver = CC.check_ginzburg_landau(golden("complex_cgl.npz"), tag, scheme)
    assert solver.bset.last_verify < 1e-10


def test_complex_embedding_of_pencil_matrices():
    """The real embedding [[Re, -Im], [Im, Re]] of one pencil's LHS reproduces the complex matrix-vector product."""
    from dedalus_b200 import examples
    from dedalus_b200.pencils import PencilSystemBuilder
    pb = examples.complex_ginzburg_landau(8, 6)
    builder = PencilSystemBuilder(pb['problem'])
    assert builder.complex
    cls = builder.find_class((3,))
    A = builder.class_matrix(cls, 'L', (3,), restrict=False).toarray()
    assert np.isrealobj(A)
    # natural ordering per owner: [re rows | im rows]; unknowns u (6), tau1, tau2 -> complex size 8, real size 16
    assert A.shape == (16, 16)
    def embed_index(n_per_owner):
        re, im, off = [], [], 0
        for n in n_per_owner:
            re += list(range(off, off + n)); im += list(range(off + n, off + 2 * n)); off += 2 * n
        return np.array(re), np.array(im)
    re, im = embed_index([6, 1, 1])
    Ac = A[np.ix_(re, re)] + 1j * A[np.ix_(im, re)]
    assert np.allclose(A[np.ix_(im, im)], Ac.real) and np.allclose(A[np.ix_(re, im)], -Ac.imag)
    assert np.abs(Ac.imag).max() > 0          # the complex diffusion / advection coefficients are really there
