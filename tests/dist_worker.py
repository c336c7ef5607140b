"""Worker for tests/test_dist_gloo.py: 2-rank (gloo, CPU, test-only kernel emulation) run of the distributed
pencil path -- block decomposition, transpose hops, rank-local pencil batches -- checked against the reference
state of the SAME global problem (tests/golden/rb3d_8.npz), i.e. multi-rank == single-rank to round-off."""
import os, sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.distributed as dist


def main():
    backend = os.environ.get("DB_DIST_BACKEND", "gloo")
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
        from emu import emu_lib as E
        E.install()                      # CPU run: test-only kernel emulation
    rank, world = dist.get_rank(), dist.get_world_size()
    import dedalus_b200 as d3
    from dedalus_b200 import examples
    which = sys.argv[1] if len(sys.argv) > 1 else "rb3d_8.npz"
    if which.startswith("blocked:"):
        # blocked transposes (the Fourier passes write / read the all-to-all buffers directly) against the pack /
        # unpack path on the SAME distributed problem: sizes chosen so the register-resident kernels cover both passes
        Nh, Nz = (int(v) for v in which.split(":")[1].split("x"))
        states = []
        for mode in ("1", "0"):
            os.environ["DB_BLOCKED_TRANSPOSE"] = mode
            pb = examples.rayleigh_benard(dim=3, Nh=Nh, Nz=Nz, Rayleigh=1e5, mesh=(world,))
            solver = pb['problem'].build_solver(d3.RK222)
            examples.rayleigh_benard_initial_condition(pb['b'], pb['bases'], pb['Lz'])
            pb['u'].fill_random('g', seed=43, distribution='normal', scale=0.5)      # O(1) flow: the products matter
            pb['u'].low_pass_filter(scales=0.5)
            for i in range(2):
                solver.step(1e-3)
            if mode == "1":
                assert solver.rhs_plan._blocked_bwd_ok() and solver.rhs_plan._blocked_fwd_ok(), "blocked path not taken"
            states.append([np.array(pb[name]['c']) for name in ('p', 'b', 'u')])
        ok = all(np.allclose(a, b, rtol=1e-11, atol=1e-13) and np.isfinite(a).all() for a, b in zip(*states))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which == "staged2d":
        # staged right-hand sides (grid function + derivative of a product) on this mesh against the single-rank reference
        import bc_cases
        g = np.load(ROOT / "tests" / "golden" / "bc_data.npz")
        res = bc_cases.rb2d_conservative(d3, mesh=(world,))
        ok = True
        for name, f in res.items():
            full = g["cons_" + name]
            rows = f.dist.coeff_local_slice(0, f.bases[0])
            ok = ok and bool(np.allclose(f['c'], full[..., rows, :], rtol=1e-8, atol=1e-11 * np.abs(full).max()))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which == "shell:forced":
        # general curvilinear right-hand sides (grid function on the RHS) on this mesh, strong-flow start, against the reference
        g = np.load(ROOT / "tests" / "golden" / "shell_strong.npz")
        Nphi, Ntheta, Nr, steps, dtv = g["strong_meta"]
        sc = examples.shell_convection(int(Nphi), int(Ntheta), int(Nr), rhs_b_extra="0.05*sin(3*b)")
        solver = sc['problem'].build_solver(d3.SBDF2)
        rows = sc['dist'].coeff_local_slice(0, sc['shell'])
        sc['b']['c'] = g["strong_b0"][rows]; sc['u']['c'] = g["strong_u0"][:, rows]
        for _ in range(int(steps)):
            solver.step(float(dtv))
        ok = type(solver.rhs_plan).__name__ == "GenericCurvilinearRHS"
        for name in ('p', 'b', 'u'):
            full = g[f"forced_{name}1"]
            ok = ok and bool(np.allclose(sc[name]['c'], full[..., rows, :, :], rtol=1e-8, atol=1e-10 * np.abs(full).max()))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which == "strong3d":
        # the benchmark's 3-D problem with O(1) velocities on this mesh against the single-rank reference: the noise-started fixtures
        # keep |u| ~ 1e-7 |b| and would not notice a wrong nonlinear term in the distributed layout
        import bc_cases
        g = np.load(ROOT / "tests" / "golden" / "bc_data.npz")
        res, init = bc_cases.rb3d_strong(d3, mesh=(world,))
        ok = True
        for name, f in res.items():
            full = g["strong3d_" + name]
            rows = f.dist.coeff_local_slice(0, f.bases[0])
            ok = ok and bool(np.allclose(f['c'], full[..., rows, :, :], rtol=1e-8, atol=1e-11 * np.abs(full).max()))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which == "plugin:transpose":
        # the reference's Transpose plugin contract (core/transposes.pyx:22-246) on this mesh: the same global array on every rank
        # (seeded), row-local block in, column-local block out and back
        from dedalus_b200.transposes import B200Transpose
        ok = True
        for dtype, shape, axis in ((np.float64, (3, 8, 12, 5), 1), (np.float64, (16, 6), 0), (np.complex128, (2, 4, 8, 3), 1)):
            rng = np.random.default_rng(5)
            G = rng.standard_normal(shape).astype(dtype)
            if dtype is np.complex128:
                G = G + 1j * rng.standard_normal(shape)
            chunk = [1] * len(shape); chunk[axis] = 2
            plan = B200Transpose(shape, chunk, dtype, axis, None)
            n1, n2 = shape[axis], shape[axis + 1]
            rs = slice(rank * n1 // world, (rank + 1) * n1 // world)
            cs = slice(rank * n2 // world, (rank + 1) * n2 // world)
            idx_r = [slice(None)] * len(shape); idx_r[axis] = rs
            idx_c = [slice(None)] * len(shape); idx_c[axis + 1] = cs
            RL = np.ascontiguousarray(G[tuple(idx_r)]); CL = np.zeros_like(np.ascontiguousarray(G[tuple(idx_c)]))
            plan.localize_columns(RL, CL)
            ok = ok and bool(np.array_equal(CL, G[tuple(idx_c)]))
            back = np.zeros_like(RL)
            plan.localize_rows(CL, back)
            ok = ok and bool(np.array_equal(back, RL))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which.startswith("tasks:"):
        # stand-alone output expressions on a distributed domain: every rank holds its block of the reference's grid / coefficient data
        kind = which.split(":")[1]
        ok = True
        if kind == "cartesian":
            import expression_cases as X
            g = np.load(ROOT / "tests" / "golden" / "expressions.npz")
            Nx, Nz = (int(v) for v in g['meta'])
            coords = d3.CartesianCoordinates('x', 'z')
            dd = d3.Distributor(coords, dtype=np.float64, mesh=(world,))
            xb = d3.RealFourier(coords['x'], size=Nx, bounds=(0, 4), dealias=3/2)
            zb = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, 1), dealias=3/2)
            u = dd.VectorField(coords, name='u', bases=(xb, zb)); b = dd.Field(name='b', bases=(xb, zb))
            cs = dd.coeff_local_slice(0, xb)
            u['c'] = g['u_c'][:, cs]; b['c'] = g['b_c'][cs]
            tasks = dict(vorticity=-d3.div(d3.skew(u)), Re=np.sqrt(u@u)/0.37, sinb_b=np.sin(b)*b + b)
            for name, op in tasks.items():
                f = op.evaluate()
                ref_c = g[f"{name}_c"]
                ok = ok and bool(np.allclose(f['c'], ref_c[cs], rtol=1e-10, atol=1e-12 * np.abs(ref_c).max()))
                f.change_scales(1)
                got = np.asarray(f['g'])
                gs = tuple(dd.grid_local_slice(ax, bb, 1) for ax, bb in enumerate((xb, zb)))
                ref_g = g[f"{name}_g"]
                ok = ok and bool(np.allclose(got, ref_g[gs], rtol=1e-10, atol=1e-12 * np.abs(ref_g).max()))
        if kind == "shell":
            g = np.load(ROOT / "tests" / "golden" / "shell_tasks.npz")
            Ri, Ro, dealias = 14, 15, 3/2
            coords = d3.SphericalCoordinates('phi', 'theta', 'r')
            dd = d3.Distributor(coords, dtype=np.float64, mesh=(world,))
            shell = d3.ShellBasis(coords, shape=(16, 8, 6), radii=(Ri, Ro), dealias=dealias, dtype=np.float64)
            b = dd.Field(name='b', bases=shell); u = dd.VectorField(coords, name='u', bases=shell)
            rows = dd.coeff_local_slice(0, shell)
            b['c'] = g['b_c'][rows]; u['c'] = g['u_c'][:, rows]
            kappa = nu = 3500 ** (-1/2)
            er = dd.VectorField(coords, bases=shell.radial_basis); er['g'][2] = 1
            flux = er @ (-kappa*d3.grad(b) + u*b)
            tasks = dict(bmid=(b(r=(Ri+Ro)/2), dealias), flux_r_outer=(flux(r=Ro), dealias), flux_phi_end=(flux(phi=3*np.pi/2), dealias),
                         Re=(np.sqrt(u@u)/nu, 1), flux=(flux, 1))
            for name, (op, scales) in tasks.items():
                f = op.evaluate()
                f.change_scales(scales)
                got, ref = np.asarray(f['g']), g[f"{name}_g"]
                th = dd.grid_local_slice(1, shell, scales)
                ok = ok and got.shape == ref[:, th].shape and bool(np.allclose(got, ref[:, th], rtol=1e-9, atol=1e-11 * np.abs(ref).max()))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which.startswith("sphere:"):
        # S2 shallow water distributed over the azimuthal pairs (coefficients) / colatitude (grid) against the reference state of
        # the same global problem
        tag = which.split(":")[1]
        g = np.load(ROOT / "tests" / "golden" / "sphere.npz")
        Nphi, Ntheta, dealias, steps, dt = g[f"{tag}_meta"]
        sw = examples.shallow_water(int(Nphi), int(Ntheta), dealias=float(dealias))
        scheme = "SBDF2" if "sbdf2" in tag else "RK222"
        solver = sw['problem'].build_solver(getattr(d3, scheme))
        examples.shallow_water_initial_condition(sw['u'], sw['h'], sw['basis'], sw['units'])
        rows = sw['dist'].coeff_local_slice(0, sw['basis'])
        ok = True
        for name in ('u', 'h'):
            ref = g[f"{tag}_{name}0"][..., rows, :]
            ok = ok and bool(np.allclose(sw[name]['c'], ref, rtol=1e-11, atol=1e-14 * np.abs(g[f"{tag}_{name}0"]).max()))
        for _ in range(int(steps)):
            solver.step(float(dt))
        for name in ('u', 'h'):
            full = g[f"{tag}_{name}1"]
            ok = ok and bool(np.allclose(sw[name]['c'], full[..., rows, :], rtol=1e-8, atol=1e-12 * np.abs(full).max()))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    if which.startswith("shell:"):
        # shell convection (config 5's problem) distributed over the azimuthal pairs / colatitude against the reference state
        tag = which.split(":")[1]
        g = np.load(ROOT / "tests" / "golden" / "shell_ivp.npz")
        Nphi, Ntheta, Nr, steps, dt = g[f"{tag}_meta"]
        sc = examples.shell_convection(int(Nphi), int(Ntheta), int(Nr))
        solver = sc['problem'].build_solver(d3.SBDF2 if "sbdf2" in tag else d3.RK222)
        examples.shell_convection_initial_condition(sc['b'], sc['shell'], sc['Ri'], sc['Ro'])
        rows = sc['dist'].coeff_local_slice(0, sc['shell'])
        ok = bool(np.allclose(sc['b']['c'], g[f"{tag}_b0"][rows], rtol=1e-11, atol=1e-13))
        for _ in range(int(steps)):
            solver.step(float(dt))
        floor = 1e-13 * np.abs(g[f"{tag}_b1"]).max()
        for name in ('p', 'b', 'u'):
            full = g[f"{tag}_{name}1"]
            ok = ok and bool(np.allclose(sc[name]['c'], full[..., rows, :, :], rtol=1e-8, atol=1e-10 * np.abs(full).max() + floor))
        for name in ('tau_b1', 'tau_b2', 'tau_u1', 'tau_u2'):
            full = g[f"{tag}_{name}1"]
            ok = ok and bool(np.allclose(sc['taus'][name]['c'], full[..., rows, :, :], rtol=1e-6, atol=1e-4 * np.abs(full).max() + 1e-20))
        flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
        dist.destroy_process_group()
        return
    g = np.load(ROOT / "tests" / "golden" / which)
    dim, Nh, Nz = int(g['dim']), int(g['Nh']), int(g['Nz'])
    pb = examples.rayleigh_benard(dim=dim, Nh=Nh, Nz=Nz, Rayleigh=float(g['Ra']), mesh=(world,))
    solver = pb['problem'].build_solver(getattr(d3, str(g['scheme'])))
    examples.rayleigh_benard_initial_condition(pb['b'], pb['bases'], pb['Lz'])
    xs = pb['dist'].coeff_local_slice(0, pb['bases'][0])
    ok = bool(np.allclose(pb['b']['c'], g['b0_c'][xs], rtol=1e-12, atol=1e-14))
    nsteps = 2
    for i in range(nsteps):
        solver.step(float(g['dt']))
        if i == 0:
            for name in ('p', 'b', 'u'):
                ref = g[f"{name}_c_step1"]
                ref = ref[:, xs] if name == 'u' else ref[xs]
                ok = ok and bool(np.allclose(pb[name]['c'], ref, rtol=1e-8, atol=1e-12))
    # grid-layout round trip of a distributed field
    b = pb['b']; c0 = b['c'].copy()
    b.change_scales(1)
    _ = b['g']
    ok = ok and bool(np.allclose(b['c'], c0, rtol=1e-11, atol=1e-12))
    flag = torch.tensor([1 if ok else 0], device='cuda' if backend == "nccl" else 'cpu')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if int(flag.item()) == 1 else "DIST_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
