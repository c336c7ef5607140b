"""Pin the oracle (oracle/) to outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np, pytest
from scipy import sparse
from oracle import transforms_oracle as T
from oracle import rb_oracle, kdv_oracle

TOL = dict(rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (16, 21), (32, 48), (12, 18), (10, 15)])
def test_real_fourier(golden, M, N):
    g = golden("transforms.npz")
    fwd, bwd = T.rf_matrices(N, M)
    for ref in ("matrix", "scipy"):
        assert np.allclose(T.apply_along(bwd, g[f"rf_{ref}_{M}_{N}_cin"], 1), g[f"rf_{ref}_{M}_{N}_gout"], **TOL)
        assert np.allclose(T.apply_along(fwd, g[f"rf_{ref}_{M}_{N}_gin"], 1), g[f"rf_{ref}_{M}_{N}_cout"], **TOL)
        assert np.allclose(T.rf_backward_fft(g[f"rf_{ref}_{M}_{N}_cin"], N, 1), g[f"rf_{ref}_{M}_{N}_gout"], **TOL)
        assert np.allclose(T.rf_forward_fft(g[f"rf_{ref}_{M}_{N}_gin"], M, 1), g[f"rf_{ref}_{M}_{N}_cout"], **TOL)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (15, 22), (12, 18)])
def test_complex_fourier(golden, M, N):
    g = golden("transforms.npz")
    fwd, bwd = T.cf_matrices(N, M)
    assert np.allclose(T.apply_along(bwd, g[f"cf_scipy_{M}_{N}_cin"], 1), g[f"cf_scipy_{M}_{N}_gout"], **TOL)
    assert np.allclose(T.apply_along(fwd, g[f"cf_scipy_{M}_{N}_gin"], 1), g[f"cf_scipy_{M}_{N}_cout"], **TOL)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (15, 22), (15, 15), (32, 48)])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_chebyshev(golden, M, N, alpha):
    g = golden("transforms.npz"); a = alpha - 0.5
    fwd, bwd = T.jacobi_matrices(N, M, a, a, -0.5, -0.5)
    for ref in ("matrix", "scipy_dct"):
        key = f"ch_{ref}_{M}_{N}_{alpha}"
        assert np.allclose(T.apply_along(bwd, g[key + "_cin"], 2), g[key + "_gout"], rtol=1e-11, atol=1e-11)
        assert np.allclose(T.apply_along(fwd, g[key + "_gin"], 2), g[key + "_cout"], rtol=1e-11, atol=1e-11)
        assert np.allclose(T.cheb_backward_fft(g[key + "_cin"], N, 2, a, a), g[key + "_gout"], rtol=1e-11, atol=1e-11)
        assert np.allclose(T.cheb_forward_fft(g[key + "_gin"], M, 2, a, a), g[key + "_cout"], rtol=1e-11, atol=1e-11)


def _golden_matrix(g, tag, name):
    shape = tuple(g[f"{tag}_{name}_shape"])
    return sparse.coo_matrix((g[f"{tag}_{name}_val"], (g[f"{tag}_{name}_row"], g[f"{tag}_{name}_col"])), shape=shape).tocsr()


@pytest.mark.parametrize("fname,groups", [("rb3d_8.npz", [(0, 0), (0, 2), (3, 0), (1, 2)]), ("rb2d_16x16.npz", [(0,), (1,), (5,)])])
def test_rb_pencil_matrices(golden, fname, groups):
    g = golden(fname)
    orc = rb_oracle.RBOracle(int(g['dim']), int(g['Nh']), int(g['Nz']), float(g['Ra']))
    for grp in groups:
        tag = "pen_" + "_".join(str(k) for k in grp)
        M, L, vr, vc, _, _ = orc.pencil_matrices(grp)
        assert np.array_equal(vr, g[f"{tag}_valid_rows"]) and np.array_equal(vc, g[f"{tag}_valid_cols"])
        for name, mine in (("M", M), ("L", L)):
            ref = _golden_matrix(g, tag, name)
            assert abs(mine - ref).max() <= 1e-11 * max(1.0, abs(ref).max()), (grp, name, abs(mine - ref).max())


@pytest.mark.parametrize("fname,transforms", [("rb3d_8.npz", "fft"), ("rb3d_8.npz", "matrix"), ("rb2d_16x16.npz", "fft"),
                                              ("rb3d_8x8x12_sbdf2.npz", "fft"), ("rb3d_16.npz", "fft")])
def test_rb_states(golden, fname, transforms):
    g = golden(fname)
    st = rb_oracle.run(int(g['dim']), int(g['Nh']), int(g['Nz']), float(g['Ra']), g['b0_c'], int(g['steps']), float(g['dt']),
                       scheme=str(g['scheme']), transforms=transforms)
    for name in ("p", "b", "u", "tau_b1", "tau_b2", "tau_u1", "tau_u2"):
        ref = g[f"{name}_c"]
        got = np.asarray(st[name]).reshape(ref.shape)
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-12), (name, np.abs(got - ref).max())


@pytest.mark.parametrize("prefix", ["kdv_", "kdv443_"])
def test_kdv(golden, prefix):
    g = golden("kdv.npz")
    orc = kdv_oracle.KdVOracle(int(g[prefix + "N"]))
    u = orc.run(g[prefix + "u0_c"], int(g[prefix + "steps"]), float(g[prefix + "dt"]), scheme=str(g[prefix + "scheme"]))
    assert np.allclose(u, g[prefix + "u_c"], rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_swsh_oracle_matches_reference(golden, tag):
    """oracle/transforms_oracle.py swsh_* pinned to the reference's SWSHColatitudeTransform (tests/golden/swsh.npz)."""
    from oracle import transforms_oracle as T
    g = golden("swsh.npz")
    Nphi, Ntheta, Lmax, Gp, Gt, Cp, Ce = (int(v) for v in g[tag + "_meta"])
    mm = g[tag + "_m_maps"]
    for s in (0, 1, -1, 2, -2):
        c = np.zeros_like(g[f"{tag}_s{s}_cout"]); T.swsh_forward(g[f"{tag}_s{s}_gin"], c, mm, Gt, Lmax, s)
        assert np.allclose(c, g[f"{tag}_s{s}_cout"], rtol=1e-12, atol=1e-13)
        gg = np.full_like(g[f"{tag}_s{s}_gout"], np.nan); T.swsh_backward(g[f"{tag}_s{s}_cin"], gg, mm, Gt, Lmax, s)
        assert np.allclose(gg, g[f"{tag}_s{s}_gout"], rtol=1e-12, atol=1e-12, equal_nan=True)
    for s in (0, 2):
        for m in (0, 3):
            F, B = T.swsh_matrices(Gt, Lmax, m, s)
            assert np.allclose(F, g[f"{tag}_s{s}_m{m}_fwdmat"], rtol=1e-12, atol=1e-13)
            assert np.allclose(B, g[f"{tag}_s{s}_m{m}_bwdmat"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag,scheme", [("sw16", "RK222"), ("sw32", "RK222"), ("sw32sbdf2", "SBDF2")])
def test_sphere_oracle_matches_reference(golden, tag, scheme):
    """oracle/sphere_oracle.py (complex per-m formulation, dense solves) vs the reference's shallow-water states."""
    from oracle import sphere_oracle
    g = golden("sphere.npz")
    Nphi, Ntheta, dealias, steps, dt = g[f"{tag}_meta"]
    out = sphere_oracle.run(int(Nphi), int(Ntheta), g[f"{tag}_u0"], g[f"{tag}_h0"], int(steps), float(dt), scheme, dealias=float(dealias))
    for name in ("u", "h"):
        ref = g[f"{tag}_{name}1"]
        assert np.allclose(out[name], ref, rtol=1e-9, atol=1e-13 * np.abs(ref).max()), name


def test_rb_oracle_with_strong_flow_matches_reference(golden):
    """The Rayleigh-Benard oracle against the reference with O(1) velocities (advection as large as the linear terms): the
    noise-started fixtures above keep |u| << |b| and would not see a wrong order of truncations in the nonlinear terms.  The bench's
    parity gate and the GPU tests at sizes without fixtures lean on this oracle at O(1) velocity."""
    from oracle import rb_oracle
    g = golden("bc_data.npz")
    ref = rb_oracle.run(dim=2, Nh=16, Nz=16, Ra=2e5, b0_c=g["strong_b0"], steps=5, dt=0.01, u0_c=g["strong_u0"])
    for name in ("p", "b", "u"):
        want = g["strong_" + name]
        assert np.allclose(ref[name], want, rtol=1e-8, atol=1e-11 * np.abs(want).max()), (name, np.abs(ref[name] - want).max())


def test_rb3d_oracle_with_strong_flow_matches_reference(golden):
    """Same in 3-D (the benchmark's problem, 8^3)."""
    from oracle import rb_oracle
    g = golden("bc_data.npz")
    ref = rb_oracle.run(dim=3, Nh=8, Nz=8, Ra=1e6, b0_c=g["strong3d_b0"], steps=3, dt=0.01, u0_c=g["strong3d_u0"])
    for name in ("p", "b", "u"):
        want = g["strong3d_" + name]
        assert np.allclose(ref[name], want, rtol=1e-8, atol=1e-11 * np.abs(want).max()), (name, np.abs(ref[name] - want).max())
