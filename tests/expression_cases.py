"""Shared body of the stand-alone expression tests (emulation: tests/test_emu_expressions.py, GPU: tests/test_gpu_t8_expressions.py).
Expected values: tests/golden/expressions.npz (make_golden.py gen_expressions, unmodified reference)."""
import numpy as np
import dedalus_b200 as d3


def setup(g):
    Nx, Nz = (int(v) for v in g['meta'])
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64)
    xb = d3.RealFourier(coords['x'], size=Nx, bounds=(0, 4), dealias=3/2)
    zb = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, 1), dealias=3/2)
    u = dist.VectorField(coords, name='u', bases=(xb, zb))
    b = dist.Field(name='b', bases=(xb, zb))
    u['c'] = g['u_c']; b['c'] = g['b_c']
    nu = 0.37
    tasks = dict(vorticity=-d3.div(d3.skew(u)), Re=np.sqrt(u@u)/nu, ke=0.5*(u@u), sinb_b=np.sin(b)*b + b,
                 grad_mag=np.sqrt(d3.grad(b)@d3.grad(b)), absdiv=np.abs(d3.div(u)) * 2.0)
    return dist, u, b, tasks


def check_expressions(g):
    """The output-task expressions of the stock Rayleigh-Benard script (rayleigh_benard.py:93-103) and other grid-function
    expressions: grid data at scale 1 and coefficient data as the reference's Future.evaluate returns them."""
    dist, u, b, tasks = setup(g)
    for name, op in tasks.items():
        for rep in range(2):                    # the second evaluation reuses the compiled program
            f = op.evaluate()
            ref_c = g[f"{name}_c"]
            assert np.allclose(f['c'], ref_c, rtol=1e-10, atol=1e-12 * np.abs(ref_c).max()), (name, "c", np.abs(f['c'] - ref_c).max())
            f.change_scales(1)
            ref_g = g[f"{name}_g"]
            assert np.allclose(f['g'], ref_g, rtol=1e-10, atol=1e-12 * np.abs(ref_g).max()), (name, "g", np.abs(f['g'] - ref_g).max())
    # reductions of expressions: volume integral, horizontal average (a profile), mid-plane value, a point value ...
    coords = dist.coordsys
    x, z = coords['x'], coords['z']
    reductions = dict(int_bb=d3.Integrate(d3.Integrate(b*b, x), z), prof_b=d3.Average(b, x), mid_uu=(u@u)(z=0.3),
                      avg_speed=2.0*d3.Average(np.sqrt(u@u), x), point=b(x=1.3)(z=0.6), vol_avg=d3.Average(d3.Average(b, z), x),
                      int_dzb=d3.Integrate(d3.Differentiate(b, z), z))
    for name, op in reductions.items():
        f = op.evaluate()
        f.change_scales(1)
        ref = g[f"red_{name}_g"]
        got = np.asarray(f['g'])
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-12 * max(np.abs(ref).max(), 1e-3)), (name, np.abs(got - ref).max(), np.abs(ref).max())
    # numpy scalars on the left keep dispatching to the operand overloads
    e = np.float64(2.0) * b + np.float64(1.0) * b
    assert np.allclose(e.evaluate()['c'], 3 * g['b_c'], rtol=1e-13, atol=1e-14)


def check_flow_property_reductions(g):
    """GlobalFlowProperty with an expression property: max / min / grid_average of the scale-1 grid values and the volume integral
    (reference extras/flow_tools.py:92-130) against numpy reductions of the reference's task output."""
    from dedalus_b200 import examples
    dist, u, b, tasks = setup(g)
    pb = examples.rayleigh_benard(dim=2, Nh=16, Nz=16, Rayleigh=1e5)          # any solver: the properties hang on its step hooks
    solver = pb['problem'].build_solver(d3.RK222)
    flow = d3.GlobalFlowProperty(solver, cadence=1)
    flow.add_property(tasks['Re'], name='Re')
    ref = g['Re_g']
    assert np.isclose(flow.max('Re'), ref.max(), rtol=1e-10)
    assert np.isclose(flow.min('Re'), ref.min(), rtol=1e-10, atol=1e-13)
    assert np.isclose(flow.grid_average('Re'), ref.mean(), rtol=1e-10)
    flow.add_property(b * b, name='bb')
    assert np.isclose(flow.volume_integral('bb'), float(g['red_int_bb_g'].ravel()[0]), rtol=1e-10)


def check_field_helpers(g):
    """fill_random + low_pass_filter reproduce the reference's fields (the fixture's inputs were made that way); norms, high-pass
    filter and global data access against the reference's values (core/field.py:782-986)."""
    Nx, Nz = (int(v) for v in g['meta'])
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64)
    xb = d3.RealFourier(coords['x'], size=Nx, bounds=(0, 4), dealias=3/2)
    zb = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, 1), dealias=3/2)
    u = dist.VectorField(coords, name='u', bases=(xb, zb)); b = dist.Field(name='b', bases=(xb, zb))
    u.fill_random('g', seed=11, distribution='normal', scale=1.0); u.low_pass_filter(scales=0.75)
    b.fill_random('g', seed=12, distribution='normal', scale=1.0); b.low_pass_filter(scales=0.75)
    assert np.allclose(u['c'], g['u_c'], rtol=1e-12, atol=1e-14) and np.allclose(b['c'], g['b_c'], rtol=1e-12, atol=1e-14)
    got = [b.allreduce_L2_norm(), b.allreduce_L2_norm(normalize_volume=False), u.allreduce_L2_norm(),
           b.allreduce_data_norm('c', 2), b.allreduce_data_max('g')]
    assert np.allclose(got, g['norms'], rtol=1e-11), (got, g['norms'])
    hp = b.copy(); hp.high_pass_filter(shape=(16, 8))
    assert np.allclose(hp['c'], g['b_highpass_c'], rtol=1e-12, atol=1e-14)
    full = b.allgather_data('g')
    c = dist.Field(bases=(xb, zb)); c.preset_scales(b.scales); c.load_from_global_grid_data(full)
    assert np.allclose(c['c'], b['c'], rtol=1e-12, atol=1e-14)
    assert b.evaluate() is b
