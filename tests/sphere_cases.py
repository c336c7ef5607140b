"""Shared bodies of the sphere (S2) parity tests: the same checks run through the CPU emulation of the kernels
(tests/test_emu_sphere.py) and on the GPU (tests/test_gpu_t2_sphere.py).  All expected values come from the unmodified
reference (tests/golden/sphere.npz, written by tests/golden/make_golden.py gen_sphere)."""
import numpy as np
import dedalus_b200 as d3
from dedalus_b200 import examples


def check_field_transforms(g, tag):
    """Scalar / vector / rank-2 fields: grid -> coefficients (reference packing) -> grid, vs the reference chain
    (core/basis.py:3062-3138: azimuthal FFT, spin recombination, SWSH colatitude transform)."""
    Nphi, Ntheta, dealias = g[f"f{tag}_meta"]
    coords = d3.S2Coordinates('phi', 'theta')
    dist = d3.Distributor(coords, dtype=np.float64)
    basis = d3.SphereBasis(coords, (int(Nphi), int(Ntheta)), radius=1.7, dealias=float(dealias), dtype=np.float64)
    phi, theta = dist.local_grids(basis, scales=(dealias, dealias))
    assert np.allclose(phi.ravel(), g[f"f{tag}_phi"], rtol=0, atol=1e-14)
    assert np.allclose(theta.ravel(), g[f"f{tag}_theta"], rtol=0, atol=1e-14)
    for name, f in (("s", dist.Field(bases=basis)), ("v", dist.VectorField(coords, bases=basis)),
                    ("t", dist.TensorField((coords, coords), bases=basis))):
        f.preset_scales(dealias)
        f['g'] = g[f"f{tag}_{name}_gin"]
        c = f['c'].copy()
        assert c.shape == g[f"f{tag}_{name}_c"].shape
        assert np.allclose(c, g[f"f{tag}_{name}_c"], rtol=1e-12, atol=1e-13), (tag, name, "forward")
        assert np.allclose(f['g'], g[f"f{tag}_{name}_g2"], rtol=1e-12, atol=1e-12), (tag, name, "backward")


def run_shallow_water(g, tag, scheme="RK222"):
    Nphi, Ntheta, dealias, steps, timestep = g[f"{tag}_meta"]
    sw = examples.shallow_water(int(Nphi), int(Ntheta), dealias=float(dealias))
    assert np.isclose(sw['timestep'], timestep, rtol=1e-15)
    solver = sw['problem'].build_solver(getattr(d3, scheme))
    examples.shallow_water_initial_condition(sw['u'], sw['h'], sw['basis'], sw['units'])
    return sw, solver, int(steps), float(timestep)


def check_shallow_water(g, tag, scheme="RK222"):
    """K steps from the analytic jet + perturbation vs the reference state: np.allclose(rtol 1e-8, atol 1e-12 max|ref|)."""
    sw, solver, steps, dt = run_shallow_water(g, tag, scheme)
    for name in ('u', 'h'):
        ref = g[f"{tag}_{name}0"]
        assert np.allclose(sw[name]['c'], ref, rtol=1e-11, atol=1e-14 * np.abs(ref).max()), (name, "initial condition")
    for _ in range(steps):
        solver.step(dt)
    for name in ('u', 'h'):
        ref = g[f"{tag}_{name}1"]
        got = sw[name]['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-12 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())
    return sw, solver


def check_pencil_matrices(g, tag="sw16"):
    """M and L of three per-m pencils against the reference's subproblem matrices in natural ordering
    (core/subsystems.py:497-602): variable-major, component, (cos | -sin), degree."""
    from dedalus_b200.sphere import lhs_blocks
    sw, solver, steps, dt = run_shallow_water(g, tag)
    basis = sw['basis']
    variables = sw['problem'].variables
    Lmax = basis.Lmax
    for m in (0, 1, 5):
        NL = Lmax + 1 - m
        comps = [(iv, c) for iv, v in enumerate(variables) for c in range(max(v.ncomp, 1))]
        n = len(comps) * 2 * NL
        mats = {0: np.zeros((n, n)), 1: np.zeros((n, n))}
        row0 = 0
        for eq in sw['problem'].equations:
            blocks = lhs_blocks(eq['LHS'], variables, basis, m)
            ncomp_eq = max(int(np.prod([cs.dim for cs in eq['tensorsig']], dtype=int)), 1)
            for t, d in blocks.items():
                for (co, iv, ci), B in d.items():
                    B = B.toarray()
                    cv = comps.index((iv, ci))
                    r = (row0 + co) * 2 * NL; c = cv * 2 * NL
                    # reference natural order of a group: component, then (cos, -sin), then degree
                    mats[t][r:r + NL, c:c + NL] += B.real; mats[t][r:r + NL, c + NL:c + 2 * NL] += -B.imag
                    mats[t][r + NL:r + 2 * NL, c:c + NL] += B.imag; mats[t][r + NL:r + 2 * NL, c + NL:c + 2 * NL] += B.real
            row0 += ncomp_eq
        for t, name in ((1, 'M'), (0, 'L')):
            ref = g[f"{tag}_m{m}_{name}"]
            # the reference's natural matrices keep the modes that do not exist (l < |s|; -sin part of l = 0) as zero rows and
            # columns, and follow the MEMORY order of the degrees: descending for the folded wavenumbers
            valid = _valid_natural(basis, variables, m)
            mine = mats[t] * np.outer(valid, valid)
            _, cols = basis.mode_columns(m)
            if len(cols) > 1 and cols[1] < cols[0]:
                perm = (np.arange(n).reshape(-1, NL)[:, ::-1]).ravel()
                mine = mine[np.ix_(perm, perm)]
            assert mine.shape == ref.shape, (m, name, mine.shape, ref.shape)
            assert np.allclose(mine, ref, rtol=1e-12, atol=1e-14 * max(np.abs(ref).max(), 1e-300)), (m, name, np.abs(mine - ref).max())


def _valid_natural(basis, variables, m):
    Lmax = basis.Lmax
    out = []
    for v in variables:
        spins = [int(s) for s in basis.spin_weights(v.tensorsig).reshape(-1)] or [0]
        for s in spins:
            for part in (0, 1):
                for l in range(m, Lmax + 1):
                    out.append(l >= max(m, abs(s)) and not (l == 0 and part == 1 and len(v.tensorsig) <= 1))
    return np.array(out)


def check_config4_size(g):
    """512 x 256 (Lmax = 254): 3 RK222 steps vs the reference's checksums and four coefficient rows."""
    sw, solver = None, None
    tag = "sw512"
    sw, solver, steps, dt = run_shallow_water(g, tag)
    for _ in range(steps):
        solver.step(dt)
    for name in ('u', 'h'):
        got = sw[name]['c']
        assert np.isclose(np.sum(got.astype(np.longdouble)**2), float(g[f"{tag}_{name}1_sumsq"]), rtol=1e-10), name
        assert np.isclose(np.abs(got).max(), float(g[f"{tag}_{name}1_absmax"]), rtol=1e-10), name
        ref = g[f"{tag}_{name}1_rows"]
        assert np.allclose(got[..., 20:24, :], ref, rtol=1e-8, atol=1e-12 * float(g[f"{tag}_{name}1_absmax"])), name
    return solver


def check_banded_kernels(B):
    """db_banded_combine / factor / solve / matvec on a ragged batch (sizes 1 .. 100, 1 .. 9 right-hand sides, and one system
    with more columns than a CTA's chunk) against dense numpy solves.  B: array backend (emulation: numpy; GPU: torch)."""
    import ctypes as C
    from dedalus_b200 import lib as dlib
    rng = np.random.default_rng(0)
    kl, ku = 5, 3
    ns, nrhs = [1, 7, 40, 33, 64, 100, 24], [1, 3, 1, 9, 2, 1, 19]
    ld0, ldf = kl + ku + 1, 2 * kl + ku + 1
    sysarr = (dlib.BandedSys * len(ns))()
    op_off = lu_off = piv_off = vec_off = 0
    for i, (n, r) in enumerate(zip(ns, nrhs)):
        s = sysarr[i]
        s.n, s.nrhs, s.op_off, s.lu_off, s.piv_off, s.vec_off = n, r, op_off, lu_off, piv_off, vec_off
        op_off += n * ld0; lu_off += n * ldf; piv_off += n; vec_off += n * r
    m_ab, l_ab = np.zeros(op_off), np.zeros(op_off)
    Ms, Ls = [], []
    for i, n in enumerate(ns):
        for ab, store in ((m_ab, Ms), (l_ab, Ls)):
            A = np.zeros((n, n))
            for a in range(n):
                for b in range(max(0, a - kl), min(n, a + ku + 1)):
                    A[a, b] = rng.standard_normal()
                    ab[sysarr[i].op_off + b * ld0 + ku + a - b] = A[a, b]
            store.append(A)
    a0, b0 = 1.0, 0.37
    v1h, v2h = rng.standard_normal(vec_off), rng.standard_normal(vec_off)
    sysb = B.dev(np.frombuffer(bytes(sysarr), dtype=np.uint8).copy())
    m_d, l_d = B.dev(m_ab), B.dev(l_ab)
    lu, ipiv, info = B.dev(np.full(lu_off, np.nan)), B.dev(np.zeros(piv_off, dtype=np.int32)), B.dev(np.full(len(ns), -1, dtype=np.int32))
    v1, v2, x, ym, yl = B.dev(v1h), B.dev(v2h), B.dev(np.zeros(vec_off)), B.dev(np.zeros(vec_off)), B.dev(np.zeros(vec_off))
    P = B.ptr
    B.lib.call("db_banded_combine", P(sysb), len(ns), kl, ku, a0, P(m_d), b0, P(l_d), P(lu), B.stream)
    B.lib.call("db_banded_factor", P(sysb), len(ns), kl, ku, P(lu), P(ipiv), P(info), B.stream)
    assert not B.host(info).any()
    vc = dlib.VecComb(); vc.nvec = 2
    vc.vec[0], vc.vec[1] = P(v1).value, P(v2).value
    vc.coef[0], vc.coef[1] = 2.0, -0.5
    B.lib.call("db_banded_solve", P(sysb), len(ns), kl, ku, max(ns), max(nrhs), P(lu), P(ipiv), C.byref(vc), P(x), B.stream)
    B.lib.call("db_banded_matvec", P(sysb), len(ns), kl, ku, P(m_d), P(l_d), P(x), P(ym), P(yl), B.stream)
    xh, ymh, ylh = B.host(x), B.host(ym), B.host(yl)
    for i, (n, r) in enumerate(zip(ns, nrhs)):
        o = sysarr[i].vec_off
        b = (2.0 * v1h[o:o + n * r] - 0.5 * v2h[o:o + n * r]).reshape(n, r)
        xr = np.linalg.solve(a0 * Ms[i] + b0 * Ls[i], b)
        xg = xh[o:o + n * r].reshape(n, r)
        assert np.abs(xg - xr).max() <= 1e-10 * np.abs(xr).max(), (n, r, np.abs(xg - xr).max())
        assert np.allclose(ymh[o:o + n * r].reshape(n, r), Ms[i] @ xg, rtol=1e-12, atol=1e-12)
        assert np.allclose(ylh[o:o + n * r].reshape(n, r), Ls[i] @ xg, rtol=1e-12, atol=1e-12)
    # index gather / scatter
    arena_h = rng.standard_normal(50)
    idx_h = np.array([3, -1, 7, 49, -1, 0], dtype=np.int64)
    arena, idx, vec = B.dev(arena_h.copy()), B.dev(idx_h), B.dev(np.full(6, np.nan))
    B.lib.call("db_index_move", P(idx), 6, P(arena), P(vec), 1, B.stream)
    assert np.array_equal(B.host(vec), np.where(idx_h >= 0, arena_h[np.maximum(idx_h, 0)], 0.0))
    vec2 = B.dev(np.arange(6.0) + 100)
    B.lib.call("db_index_move", P(idx), 6, P(arena), P(vec2), 0, B.stream)
    exp = arena_h.copy(); exp[idx_h[idx_h >= 0]] = (np.arange(6.0) + 100)[idx_h >= 0]
    assert np.array_equal(B.host(arena), exp)


def check_banded_singular(B):
    from dedalus_b200 import lib as dlib
    sysarr = (dlib.BandedSys * 1)()
    sysarr[0].n, sysarr[0].nrhs = 4, 1
    sysb = B.dev(np.frombuffer(bytes(sysarr), dtype=np.uint8).copy())
    lu, ipiv, info = B.dev(np.zeros(16)), B.dev(np.zeros(4, dtype=np.int32)), B.dev(np.zeros(1, dtype=np.int32))
    B.lib.call("db_banded_factor", B.ptr(sysb), 1, 1, 1, B.ptr(lu), B.ptr(ipiv), B.ptr(info), B.stream)
    assert B.host(info)[0] == 4


def check_against_oracle(Nphi, Ntheta, scheme, dts, seed=7):
    """Random (smooth) initial state, a time-step sequence with changes (every change refactorises all per-m systems), against
    oracle/sphere_oracle.py -- an independent complex, dense formulation pinned to the reference in tests/test_oracle.py."""
    from oracle import sphere_oracle
    sw = examples.shallow_water(Nphi, Ntheta)
    solver = sw['problem'].build_solver(getattr(d3, scheme))
    examples.shallow_water_initial_condition(sw['u'], sw['h'], sw['basis'], sw['units'])
    rng = np.random.default_rng(seed)
    phi, theta = sw['dist'].local_grids(sw['basis'])
    sw['h']['g'] += 1e-5 * np.cos(3 * phi) * np.sin(theta)**3 * rng.standard_normal()
    u0, h0 = sw['u']['c'].copy(), sw['h']['c'].copy()
    for dt in dts:
        solver.step(dt)
    ref = sphere_oracle.run(Nphi, Ntheta, u0, h0, len(dts), dts, scheme)
    for name in ('u', 'h'):
        got = sw[name]['c']
        assert np.allclose(got, ref[name], rtol=1e-8, atol=1e-12 * np.abs(ref[name]).max()), (name, np.abs(got - ref[name]).max())
    return solver


def check_analysis_tasks(g, tag="sw16"):
    """The stock script's output tasks on the reference's final state: a field, the vorticity -div(skew(u)) and a Laplacian,
    evaluated on the device through a dictionary handler (reference Future.evaluate / Handler.add_task)."""
    sw = examples.shallow_water(*(int(v) for v in g[f"{tag}_meta"][:2]))
    solver = sw['problem'].build_solver(d3.RK222)
    sw['u']['c'] = g[f"{tag}_u1"]; sw['h']['c'] = g[f"{tag}_h1"]
    handler = solver.evaluator.add_dictionary_handler(iter=1)
    handler.add_task(sw['h'], layout='c', name='height')
    handler.add_task(-d3.div(d3.skew(sw['u'])), layout='c', name='vorticity')
    handler.add_task(d3.lap(sw['h']), layout='c', name='lap_h')
    solver.evaluator.evaluate_handlers(iteration=0, wall_time=0.0, sim_time=0.0, timestep=0.0)
    assert np.array_equal(handler.fields['height'], g[f"{tag}_h1"])
    for name, key in (('vorticity', 'vort1'), ('lap_h', 'laph1')):
        ref = g[f"{tag}_{key}"]
        assert np.allclose(handler.fields[name], ref, rtol=1e-13, atol=1e-15 * np.abs(ref).max()), name


def check_cfl_curvilinear(g):
    """extras.flow_tools.CFL on a sphere and in a shell: the device reduction equals the grid maximum of the reference's
    AdvectiveCFL operator (core/basis.py:6156-6212) for the same random velocity."""
    sw = examples.shallow_water(32, 16)
    sw['basis'].radius = 2.5            # only the CFL spacing reads it here: the fixture's sphere has radius 2.5
    solver = sw['problem'].build_solver(d3.RK222)
    solver._init_device()
    sw['u']['c'] = g['sphere_u_c']
    cfl = d3.CFL(solver, initial_dt=1.0, cadence=1)
    cfl.add_velocity(sw['u'])
    cfl._on_step(solver)
    assert np.isclose(float(cfl.max_freq.item()), float(g['sphere_fmax']), rtol=1e-12)
    sc = examples.shell_convection(32, 16, 12)
    solver = sc['problem'].build_solver(d3.SBDF2)
    solver._init_device()
    sc['u']['c'] = g['shell_u_c']
    cfl = d3.CFL(solver, initial_dt=1.0, cadence=1)
    cfl.add_velocity(sc['u'])
    cfl._on_step(solver)
    assert np.isclose(float(cfl.max_freq.item()), float(g['shell_fmax']), rtol=1e-12)


def check_balanced_shallow_water(g, tag):
    """The stock script start to finish (examples/ivp_sphere_shallow_water/shallow_water.py:45-86): zonal jet -> LBVP for the
    balanced height ("g*lap(h) + c = - div(u@grad(u) + 2*Omega*zcross(u))", "ave(h) = 0") -> perturbation -> RK222 steps.
    Expected values: tests/golden/sphere_lbvp.npz (make_golden.py shallow_water_balanced, unmodified reference)."""
    Nphi, Ntheta, dealias, steps, timestep = g[f"{tag}_meta"]
    Nphi, Ntheta, steps = int(Nphi), int(Ntheta), int(steps)
    meter = 1 / 6.37122e6; hour = 1; second = hour / 3600
    R = 6.37122e6 * meter; Omega = 7.292e-5 / second; nu = 1e5 * meter**2 / second / 32**2
    g_ = 9.80616 * meter / second**2; H = 1e4 * meter
    coords = d3.S2Coordinates('phi', 'theta')
    dist = d3.Distributor(coords, dtype=np.float64)
    basis = d3.SphereBasis(coords, (Nphi, Ntheta), radius=R, dealias=float(dealias), dtype=np.float64)
    u = dist.VectorField(coords, name='u', bases=basis)
    h = dist.Field(name='h', bases=basis)
    zcross = lambda A: d3.MulCosine(d3.skew(A))
    phi, theta = dist.local_grids(basis)
    lat = np.pi / 2 - theta + 0*phi
    umax = 80 * meter / second
    lat0 = np.pi / 7; lat1 = np.pi / 2 - lat0
    en = np.exp(-4 / (lat1 - lat0)**2)
    jet = (lat0 <= lat) * (lat <= lat1)
    u_jet = umax / en * np.exp(1 / (lat[jet] - lat0) / (lat[jet] - lat1))
    u['g'][0][jet] = u_jet
    c = dist.Field(name='c')
    ns = dict(g=g_, Omega=Omega, zcross=zcross, u=u, h=h, c=c, nu=nu, H=H)
    problem = d3.LBVP([h, c], namespace=ns)
    problem.add_equation("g*lap(h) + c = - div(u@grad(u) + 2*Omega*zcross(u))")
    problem.add_equation("ave(h) = 0")
    solver = problem.build_solver()
    solver.solve()
    ref = g[f"{tag}_h_bal"]
    got = h['c']
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max()), ("balanced height", np.abs(got - ref).max(), np.abs(ref).max())
    assert abs(float(np.asarray(c['c']).ravel()[0])) <= 1e-12 * np.abs(ref).max()
    assert np.allclose(u['c'], g[f"{tag}_u_bal"], rtol=1e-11, atol=1e-14)          # the right-hand side field is left untouched
    lat2 = np.pi / 4; hpert = 120 * meter; alpha = 1 / 3; beta = 1 / 15
    h.change_scales(1); u.change_scales(1)
    h['g'] += hpert * np.cos(lat) * np.exp(-(phi/alpha)**2) * np.exp(-((lat2-lat)/beta)**2)
    problem = d3.IVP([u, h], namespace=ns)
    problem.add_equation("dt(u) + nu*lap(lap(u)) + g*grad(h) + 2*Omega*zcross(u) = - u@grad(u)")
    problem.add_equation("dt(h) + nu*lap(lap(h)) + H*div(u) = - div(h*u)")
    solver = problem.build_solver(d3.RK222)
    for _ in range(steps):
        solver.step(float(timestep))
    for name, f in (('u', u), ('h', h)):
        ref = g[f"{tag}_{name}1"]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())
