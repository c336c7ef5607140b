"""Shared bodies of the plugin-contract tests: the reference's Matsolver and Transpose plugin interfaces served by the device kernels
(emulation: tests/test_emu_plugins.py, GPU: tests/test_gpu_t9_plugins.py)."""
import numpy as np
from scipy import sparse
from scipy.sparse import linalg as spla


def bordered_banded(n, kl, ku, nborder, seed):
    """Pencil-like matrix: a band plus dense boundary rows and tau columns."""
    rng = np.random.default_rng(seed)
    A = sparse.diags([rng.standard_normal(n - abs(k)) for k in range(-kl, ku + 1)], list(range(-kl, ku + 1)), format='lil')
    A.setdiag(A.diagonal() + 4.0)
    if nborder:
        A[n - nborder:, :] = rng.standard_normal((nborder, n))
        A[:, n - nborder:] = rng.standard_normal((n, nborder))
    return A.tocsr()


def check_matsolvers():
    """matsolvers[name](matrix).solve(vector) as libraries/matsolvers.py:126-194 define it, against SuperLU."""
    from dedalus_b200.matsolvers import matsolvers
    assert {'b200dense', 'b200banded'} <= set(matsolvers)
    rng = np.random.default_rng(3)
    for name, A in (("b200dense", bordered_banded(70, 3, 2, 4, 1)), ("b200dense", bordered_banded(33, 1, 1, 0, 2)),
                    ("b200banded", bordered_banded(90, 4, 3, 0, 3)), ("b200banded", bordered_banded(40, 1, 6, 0, 4))):
        solver = matsolvers[name](A, solver=None)
        lu = spla.splu(A.tocsc())
        for shape in ((A.shape[0],), (A.shape[0], 1), (A.shape[0], 5), (A.shape[0], 37)):
            b = rng.standard_normal(shape)
            x = solver.solve(b)
            assert x.shape == b.shape and x is not b
            ref = lu.solve(b)
            assert np.allclose(x, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max()), (name, shape, np.abs(x - ref).max())
    singular = sparse.csr_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))
    from dedalus_b200.lib import DedalusB200Error
    for name in ("b200dense", "b200banded"):
        try:
            matsolvers[name](singular)
        except DedalusB200Error:
            pass
        else:
            raise AssertionError("singular matrix not reported")


def check_transpose_single_rank():
    """One rank: both directions are copies (the reference builds no transposes then, distributor.py:131-175)."""
    from dedalus_b200.transposes import B200Transpose
    shape = (3, 8, 6, 5)
    plan = B200Transpose(shape, (1, 2, 1, 1), np.float64, 1, None)
    rng = np.random.default_rng(0)
    RL = rng.standard_normal(shape); CL = np.zeros(shape)
    plan.localize_columns(RL, CL)
    assert np.array_equal(RL, CL)
    back = np.zeros(shape)
    plan.localize_rows(CL, back)
    assert np.array_equal(back, RL)
