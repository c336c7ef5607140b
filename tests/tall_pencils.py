"""Shared by the emulated and the GPU tests: state parity against the oracle for TALL pencils (few horizontal modes, the
benchmark's Nz), with an O(1) velocity in the initial condition so that every variable of every pencil class -- including the
decoupled mean / shear-flow Helmholtz components (kx = 0 or ky = 0) whose static pivot order was unstable in round 1 --
carries signal well above the comparison's atol."""
import numpy as np
import dedalus_b200 as d3
from dedalus_b200 import examples

# Tolerance: b, u: np.allclose(rtol=1e-8, atol=1e-12), the golden-state tests' tolerance.  p: atol = 1e-10 * max|p| (SURVEY.md
# section 8c: "pencil solve <= 1e-10 relative (conditioning)").  The pencil systems' condition number is ~1e6 at Nz = 256 and
# the pressure is the Lagrange multiplier of the (initially non-solenoidal) O(1) velocity, p ~ div(u) / dt: two backward-stable
# solvers with different pivot orders agree on it to cond * eps ~ 1e-10 only.  Measured at 4 x 4 x 256, dt = 2.5e-3, 2 steps,
# against an extended-precision refinement of every solve: this path 6e-11 * max|p|, the oracle's SuperLU 5e-12 * max|p|
# (random right-hand sides: static order 7e-12, SuperLU 4e-12, LAPACK partial pivoting 1e-11 relative forward error).
RTOL, ATOL = 1e-8, 1e-12


def run_and_compare(Nh, Nz, dts, scheme="RK222", Ra=1e6, dim=3):
    from oracle import rb_oracle
    pb = examples.rayleigh_benard(dim=dim, Nh=Nh, Nz=Nz, Rayleigh=Ra)
    solver = pb['problem'].build_solver(getattr(d3, scheme))
    examples.rayleigh_benard_initial_condition(pb['b'], pb['bases'], pb['Lz'])
    u = pb['u']
    u.fill_random('g', seed=7, distribution='normal', scale=1.0)
    uc = np.array(u['c']); uc[..., Nz // 2:] = 0; u['c'] = uc          # resolved in z
    b0, u0 = np.array(pb['b']['c']), np.array(pb['u']['c'])
    assert np.abs(u0).max() > 0.05
    for dt in dts:
        solver.step(dt)
    ref = rb_oracle.run(dim=dim, Nh=Nh, Nz=Nz, Ra=Ra, b0_c=b0, steps=len(dts), dt=list(dts), scheme=scheme, u0_c=u0)
    worst = {}
    for name in ("p", "b", "u"):
        got = np.array(pb[name]['c'])
        assert np.abs(ref[name]).max() > 1e-3, name
        worst[name] = float(np.abs(got - ref[name]).max() / np.abs(ref[name]).max())
        assert np.allclose(got, ref[name], rtol=RTOL, atol=1e-10 * np.abs(ref[name]).max() if name == "p" else ATOL), (name, worst[name])
    return solver, worst
