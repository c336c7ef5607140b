"""Edge cases of the C-ABI entry points under the test-only CPU emulation: empty inputs, single lines, ragged tiles,
sizes outside every specialised path (the reference's own tests sweep odd / prime sizes: tests/test_transforms.py)."""
import numpy as np, pytest
from dedalus_b200 import jacobi
from emu import emu_lib as E


def test_empty_inputs_launch_nothing():
    lib = E.emu(); plan = E.EmuPlan(24, 'real'); cplan = E.EmuPlan(24, 'complex')
    z = np.zeros(0)
    n0 = lib.launches
    lib.call("db_rfft_backward", plan.ref(), E.ptr(z), E.ptr(z), 0, 16, 32, 0, 0.0, None)
    lib.call("db_rfft_forward", plan.ref(), E.ptr(z), E.ptr(z), 3, 16, 0, None)
    lib.call("db_cfft_forward", cplan.ref(), E.ptr(z), E.ptr(z), 0, 16, 4, None)
    lib.call("db_cheb_backward", plan.ref(), E.ptr(z), E.ptr(z), 0, 16, 1, None, 0, None, 0, None)
    lib.call("db_cheb_forward", plan.ref(), E.ptr(z), E.ptr(z), 0, 16, 1, None, 0, None)
    lib.call("db_band_lines", E.ptr(z), E.ptr(z), 0, 16, None, 0, None, 0, 1, None)
    lib.call("db_pointwise", E.ptr(z), E.ptr(z), 0, 2, 1, None, None, None, None, 0, None)
    assert lib.launches == n0 + 7                      # calls counted, all returned 0 without touching memory


@pytest.mark.parametrize("lines", [1, 2, 15, 17])
def test_chebyshev_register_kernels_partial_tiles(lines):
    from oracle import transforms_oracle as T
    lib = E.emu(); N, M = 48, 32; plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(lines)
    c = rng.standard_normal((lines, M)); g = np.full((lines + 1, N), np.nan)      # one guard line after the output
    lib.call("db_cheb_backward", plan.ref(), E.ptr(c), E.ptr(g), lines, M, 1, None, 0, None, 0, None)
    assert np.isnan(g[lines]).all()                                                  # nothing written past the last line
    assert np.allclose(g[:lines], T.cheb_backward_fft(c, N, 1), rtol=0, atol=1e-12)
    back = np.full((lines + 1, M), np.nan)
    lib.call("db_cheb_forward", plan.ref(), E.ptr(np.ascontiguousarray(g[:lines])), E.ptr(back), lines, M, 1, None, 0, None)
    assert np.isnan(back[lines]).all()
    assert np.allclose(back[:lines], c, rtol=0, atol=1e-12)


@pytest.mark.parametrize("M,N", [(14, 21), (10, 15), (22, 33), (26, 39)])
def test_sizes_outside_every_specialised_path(M, N):
    """Odd grid sizes with factors 3, 5, 7, 11, 13: generic radix passes of csrc/fft.cu against the matrix transforms."""
    from oracle import transforms_oracle as T
    lib = E.emu(); plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(N)
    c = rng.standard_normal((3, M, 5)); c[:, 1, :] = 0
    g = np.zeros((3, N, 5))
    lib.call("db_rfft_backward", plan.ref(), E.ptr(c), E.ptr(g), 3, M, 5, 0, 0.0, None)
    Fm, Bm = T.rf_matrices(N, M)
    assert np.allclose(g, T.apply_along(Bm, c, 1), rtol=0, atol=1e-12)
    back = np.zeros_like(c)
    lib.call("db_rfft_forward", plan.ref(), E.ptr(g), E.ptr(back), 3, M, 5, None)
    assert np.allclose(back, T.apply_along(Fm, g, 1), rtol=0, atol=1e-12)
    cz = rng.standard_normal((4, M)); gz = np.zeros((4, N))
    lib.call("db_cheb_backward", plan.ref(), E.ptr(cz), E.ptr(gz), 4, M, 1, None, 0, None, 0, None)
    Fj, Bj = T.jacobi_matrices(N, M, -0.5, -0.5, -0.5, -0.5)
    assert np.allclose(gz, cz @ Bj.T if Bj.shape == (N, M) else T.apply_along(Bj, cz, 1), rtol=0, atol=1e-11)


def test_ragged_system_counts_in_tiles():
    """Batches whose system count is not a multiple of the 64-system tile (and smaller than one tile) go through the fused
    kernels with padded lanes: whole-solver run at a size where every batch is ragged."""
    import dedalus_b200 as d3
    from dedalus_b200 import examples
    from oracle import rb_oracle
    E.install()
    try:
        pb = examples.rayleigh_benard(dim=3, Nh=6, Nz=6, Rayleigh=1e5)
        solver = pb['problem'].build_solver(d3.RK222)
        examples.rayleigh_benard_initial_condition(pb['b'], pb['bases'], pb['Lz'])
        b0 = pb['b']['c'].copy()
        for _ in range(2):
            solver.step(0.01)
        ref = rb_oracle.run(dim=3, Nh=6, Nz=6, Ra=1e5, b0_c=b0, steps=2, dt=0.01, scheme="RK222")
        for name in ("p", "b", "u"):
            assert np.allclose(pb[name]['c'], ref[name], rtol=1e-8, atol=1e-12), name
    finally:
        E.uninstall()
