"""Shared body of the spherical-shell transform tests (emulation: tests/test_emu_shell.py, GPU: tests/test_gpu_t6_shell.py)."""
import numpy as np
import dedalus_b200 as d3


def check_shell_field_transforms(g, tag):
    """Scalar / vector / rank-2 fields on a ShellBasis: grid -> coefficients (regularity components, reference packing) -> grid,
    vs the reference chain (core/basis.py:4474-4508: radial factor, regularity recombination Q(l), radial Jacobi transform; plus
    the sphere chain of core/basis.py:3062-3138 with three spin components)."""
    *shape, dealias, k = g[f"{tag}_meta"]
    shape = tuple(int(v) for v in shape)
    coords = d3.SphericalCoordinates('phi', 'theta', 'r')
    dist = d3.Distributor(coords, dtype=np.float64)
    shell = d3.ShellBasis(coords, shape=shape, radii=(1.2, 2.7), dealias=float(dealias), dtype=np.float64, k=int(k))
    grids = dist.local_grids(shell, scales=(dealias,) * 3)
    for got, name in zip(grids, ("phi", "theta", "r")):
        assert np.allclose(got.ravel(), g[f"{tag}_{name}"], rtol=0, atol=1e-14), name
    for name, f in (("s", dist.Field(bases=shell)), ("v", dist.VectorField(coords, bases=shell)),
                    ("t", dist.TensorField((coords, coords), bases=shell))):
        f.preset_scales(dealias)
        f['g'] = g[f"{tag}_{name}_gin"]
        c = f['c'].copy()
        assert c.shape == g[f"{tag}_{name}_c"].shape
        assert np.allclose(c, g[f"{tag}_{name}_c"], rtol=1e-12, atol=1e-13), (tag, name, "forward", np.abs(c - g[f"{tag}_{name}_c"]).max())
        assert np.allclose(f['g'], g[f"{tag}_{name}_g2"], rtol=1e-12, atol=1e-12), (tag, name, "backward")


def check_intertwiner_orthogonal():
    """Q(l) is orthogonal on the allowed components and zero on the forbidden ones (l < rank)."""
    from dedalus_b200.shell import Intertwiner
    for rank in (1, 2):
        for ell in range(0, 12):
            Q = Intertwiner(ell).matrix(rank)
            G = Q.T @ Q
            d = np.diag(G)
            assert np.allclose(G, np.diag(d), atol=1e-13)
            assert np.all((np.abs(d - 1) < 1e-13) | (np.abs(d) < 1e-13))
            if ell >= rank:
                assert np.allclose(d, 1)
