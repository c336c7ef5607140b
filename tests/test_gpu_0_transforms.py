"""GPU parity of the transform plugins (dedalus_b200/transforms.py -> csrc/fft.cu through the C ABI) against the
reference outputs in tests/golden/transforms.npz, mirroring the reference's own cross-implementation tests
(dedalus/tests/test_transforms.py:18-57 real Fourier, 117-158 Chebyshev) plus round trips at benchmark sizes."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-12, atol=1e-12)   # fp64: max|d| <= 1e-13*log2(N) scale, stated in SURVEY.md section 8c


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (16, 21), (32, 48), (12, 18), (10, 15)])
def test_real_fourier_vs_reference(golden, M, N):
    import torch
    from dedalus_b200.transforms import RealFourierTransform
    g = golden("transforms.npz"); plan = RealFourierTransform(N, M)
    for ref in ("matrix", "scipy"):
        cin, gout = g[f"rf_{ref}_{M}_{N}_cin"], g[f"rf_{ref}_{M}_{N}_gout"]
        out = torch.zeros(gout.shape, dtype=torch.float64, device='cuda')
        plan.backward(_t(cin), out, 1)
        assert np.allclose(out.cpu().numpy(), gout, **TOL)
        gin, cout = g[f"rf_{ref}_{M}_{N}_gin"], g[f"rf_{ref}_{M}_{N}_cout"]
        out = torch.zeros(cout.shape, dtype=torch.float64, device='cuda')
        plan.forward(_t(gin), out, 1)
        assert np.allclose(out.cpu().numpy(), cout, **TOL)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (15, 22), (12, 18)])
def test_complex_fourier_vs_reference(golden, M, N):
    import torch
    from dedalus_b200.transforms import ComplexFourierTransform
    g = golden("transforms.npz"); plan = ComplexFourierTransform(N, M)
    for ref in ("matrix", "scipy"):
        cin, gout = g[f"cf_{ref}_{M}_{N}_cin"], g[f"cf_{ref}_{M}_{N}_gout"]
        out = torch.zeros(gout.shape, dtype=torch.complex128, device='cuda')
        plan.backward(_t(cin), out, 1)
        assert np.allclose(out.cpu().numpy(), gout, **TOL)
        gin, cout = g[f"cf_{ref}_{M}_{N}_gin"], g[f"cf_{ref}_{M}_{N}_cout"]
        out = torch.zeros(cout.shape, dtype=torch.complex128, device='cuda')
        plan.forward(_t(gin), out, 1)
        assert np.allclose(out.cpu().numpy(), cout, **TOL)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (15, 22), (15, 15), (32, 48)])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_chebyshev_vs_reference(golden, M, N, alpha):
    import torch
    from dedalus_b200.transforms import FastChebyshevTransform
    g = golden("transforms.npz"); a = b = alpha - 0.5
    plan = FastChebyshevTransform(N, M, a, b, -0.5, -0.5)
    for ref in ("matrix", "scipy_dct"):
        key = f"ch_{ref}_{M}_{N}_{alpha}"
        out = torch.zeros(g[key + "_gout"].shape, dtype=torch.float64, device='cuda')
        plan.backward(_t(g[key + "_cin"]), out, 2)
        assert np.allclose(out.cpu().numpy(), g[key + "_gout"], **TOL)
        out = torch.zeros(g[key + "_cout"].shape, dtype=torch.float64, device='cuda')
        plan.forward(_t(g[key + "_gin"]), out, 2)
        assert np.allclose(out.cpu().numpy(), g[key + "_cout"], **TOL)


@pytest.mark.parametrize("M,N", [(16, 24), (12, 12)])
@pytest.mark.parametrize("ab", [(0.0, 0.0), (1.0, 0.5)])
def test_jacobi_matrix_transform_vs_reference(golden, M, N, ab):
    import torch
    from dedalus_b200.transforms import JacobiMatrixTransform
    g = golden("transforms.npz"); a, b = ab
    plan = JacobiMatrixTransform(N, M, a, b, a, b)
    assert np.allclose(plan.forward_matrix, g[f"jac_{M}_{N}_{a}_{b}_fwdmat"], rtol=1e-11, atol=1e-12)
    assert np.allclose(plan.backward_matrix, g[f"jac_{M}_{N}_{a}_{b}_bwdmat"], rtol=1e-11, atol=1e-12)
    rng = np.random.default_rng(0)
    c = rng.standard_normal((3, M, 5)); out = torch.zeros((3, N, 5), dtype=torch.float64, device='cuda')
    plan.backward(_t(c), out, 1)
    assert np.allclose(out.cpu().numpy(), np.einsum('ij,ojr->oir', g[f"jac_{M}_{N}_{a}_{b}_bwdmat"], c), **TOL)
    back = torch.zeros((3, M, 5), dtype=torch.float64, device='cuda')
    plan.forward(out, back, 1)
    assert np.allclose(back.cpu().numpy(), c, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_round_trips_at_benchmark_sizes(axis):
    """Size-independent property at the 256^3 / 3/2-dealias line length: backward then forward is the identity."""
    import torch
    from dedalus_b200.transforms import RealFourierTransform, FastChebyshevTransform
    torch.manual_seed(0)
    M, N = 256, 384
    shape = [40, 36, 33]; shape[axis] = M
    c = torch.randn(shape, dtype=torch.float64, device='cuda')
    gshape = list(shape); gshape[axis] = N
    g = torch.empty(gshape, dtype=torch.float64, device='cuda'); back = torch.empty_like(c)
    cheb = FastChebyshevTransform(N, M, 1.5, 1.5, -0.5, -0.5)
    cheb.backward(c, g, axis); cheb.forward(g, back, axis)
    assert torch.allclose(back, c, rtol=1e-9, atol=1e-9)       # conversion back-solve conditioning ~1e3
    rf = RealFourierTransform(N, M)
    idx = [slice(None)] * 3; idx[axis] = 1
    c[tuple(idx)] = 0                                            # -sin(0 x) slot carries no data
    rf.backward(c, g, axis); rf.forward(g, back, axis)
    assert torch.allclose(back, c, rtol=1e-12, atol=1e-12)


def test_fused_derivatives_match_matrices():
    import torch
    from dedalus_b200.transforms import RealFourierTransform, FastChebyshevTransform
    from dedalus_b200 import jacobi
    rng = np.random.default_rng(3)
    M, N = 24, 36
    c = rng.standard_normal((5, M))
    # Chebyshev: d/dz fused == (D c) transformed from the derivative basis
    stretch = 0.5
    plan = FastChebyshevTransform(N, M, -0.5, -0.5, -0.5, -0.5, stretch=stretch)
    for d in (1, 2):
        out = torch.empty((5, N), dtype=torch.float64, device='cuda')
        plan.backward(_t(c), out, 1, deriv=d)
        dc = c.copy(); a = -0.5
        for j in range(d):
            dc = (jacobi.differentiation_matrix(M, a + j, a + j) / stretch @ dc.T).T
        z = jacobi.gauss_grid(N, -0.5, -0.5)[0]
        ref = dc @ jacobi.polynomials(M, a + d, a + d, z)
        assert np.allclose(out.cpu().numpy(), ref, rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("M,N", [(256, 384), (128, 192), (64, 96), (32, 48), (16, 24), (512, 768), (100, 192)])
@pytest.mark.parametrize("deriv", [0, 1, 2])
def test_register_resident_real_fourier(M, N, deriv):
    """csrc/rfft_regs.cu (the path the 256^3 benchmark takes on its two Fourier axes) against the CPU oracle."""
    import torch
    from dedalus_b200.transforms import RealFourierTransform
    from dedalus_b200.lib import get_lib
    from oracle import transforms_oracle as T
    rng = np.random.default_rng(11 + deriv)
    inner, outer = 48, 3
    kscale = 2 * np.pi / 3.0
    rf = RealFourierTransform(N, M, kscale=kscale)
    c = rng.standard_normal((outer, M, inner)); c[:, 1, :] = 0
    cd = c.copy()
    for _ in range(deriv):
        k = (np.arange(M) // 2) * kscale
        a, b = cd[:, 0::2, :].copy(), cd[:, 1::2, :].copy()
        cd[:, 0::2, :] = -b * k[0::2, None]; cd[:, 1::2, :] = a * k[1::2, None]
    g = torch.full((outer, N, inner), float('nan'), dtype=torch.float64, device='cuda')
    served = get_lib().rfft_regs_launches()
    rf.backward(_t(c), g, 1, deriv=deriv)
    assert get_lib().rfft_regs_launches() == served + 1
    ref = T.rf_backward_fft(cd, N, 1)
    assert np.allclose(g.cpu().numpy(), ref, rtol=0, atol=1e-13 * max(1.0, np.abs(ref).max()))
    if deriv == 0:
        gr = rng.standard_normal((outer, N, inner))
        out = torch.full((outer, M, inner), float('nan'), dtype=torch.float64, device='cuda')
        rf.forward(_t(gr), out, 1)
        assert np.allclose(out.cpu().numpy(), T.rf_forward_fft(gr, M, 1), rtol=0, atol=1e-13)


@pytest.mark.parametrize("M,N", [(256, 384), (128, 192), (64, 96), (32, 48), (16, 24), (200, 384)])
@pytest.mark.parametrize("alpha", [0, 2])
def test_register_resident_chebyshev(M, N, alpha):
    """csrc/rfft_regs.cu k_chbwd_regs / k_chfwd_regs (the z passes of the 256^3 benchmark) against the CPU oracle."""
    import torch
    from dedalus_b200.transforms import FastChebyshevTransform
    from dedalus_b200.lib import get_lib
    from oracle import transforms_oracle as T
    rng = np.random.default_rng(3 + alpha)
    a = -0.5 + alpha
    plan = FastChebyshevTransform(N, M, a, a, -0.5, -0.5)
    plain = FastChebyshevTransform(N, M, -0.5, -0.5, -0.5, -0.5)
    c = rng.standard_normal((7, 3, M))
    g = torch.full((7, 3, N), float('nan'), dtype=torch.float64, device='cuda')
    served = get_lib().rfft_regs_launches()
    plain.backward(_t(c), g, 2)
    assert get_lib().rfft_regs_launches() == served + 1
    ref = T.cheb_backward_fft(c, N, 2)
    assert np.allclose(g.cpu().numpy(), ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
    gr = rng.standard_normal((7, 3, N))
    out = torch.full((7, 3, M), float('nan'), dtype=torch.float64, device='cuda')
    plan.forward(_t(gr), out, 2)
    assert get_lib().rfft_regs_launches() == served + 2
    ref = T.cheb_forward_fft(gr, M, 2, a, a)
    assert np.allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("M,N", [(256, 384), (64, 96), (16, 24)])
def test_chebyshev_derivative_fused_scan_matches_matrices(M, N):
    """backward(deriv=1) on contiguous lines takes db_cheb_backward_scan (derivative + back-conversion by a warp scan on
    the staged lines + DCT-III in one kernel); reference: differentiation matrix then the plain oracle transform."""
    import torch
    from dedalus_b200.transforms import FastChebyshevTransform
    from dedalus_b200 import jacobi
    rng = np.random.default_rng(M)
    stretch = 0.5
    plan = FastChebyshevTransform(N, M, -0.5, -0.5, -0.5, -0.5, stretch=stretch)
    c = rng.standard_normal((37, M))
    out = torch.full((37, N), float('nan'), dtype=torch.float64, device='cuda')
    plan.backward(_t(c), out, 1, deriv=1)
    dc = (jacobi.differentiation_matrix(M, -0.5, -0.5) / stretch @ c.T).T
    z = jacobi.gauss_grid(N, -0.5, -0.5)[0]
    ref = dc @ jacobi.polynomials(M, 0.5, 0.5, z)
    assert np.allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_register_kernels_match_reference_at_benchmark_lengths(golden):
    """The UNMODIFIED reference's own transforms (ScipyRealFFT, ScipyFastChebyshev; tests/golden/transforms_bench.npz) at the
    256 -> 384 and 128 -> 192 line lengths through the plugins, in the layouts the register-resident kernels take."""
    import torch
    from dedalus_b200.transforms import RealFourierTransform, FastChebyshevTransform
    from dedalus_b200.lib import get_lib
    g = golden("transforms_bench.npz")
    for (M, N) in [(256, 384), (128, 192)]:
        rf = RealFourierTransform(N, M)
        served = get_lib().rfft_regs_launches()
        out = torch.full(g[f"rf_{M}_{N}_gout"].shape, float('nan'), dtype=torch.float64, device='cuda')
        rf.backward(_t(g[f"rf_{M}_{N}_cin"]), out, 1)
        assert np.allclose(out.cpu().numpy(), g[f"rf_{M}_{N}_gout"], **TOL)
        out = torch.full(g[f"rf_{M}_{N}_cout"].shape, float('nan'), dtype=torch.float64, device='cuda')
        rf.forward(_t(g[f"rf_{M}_{N}_gin"]), out, 1)
        assert np.allclose(out.cpu().numpy(), g[f"rf_{M}_{N}_cout"], **TOL)
        assert get_lib().rfft_regs_launches() == served + 2
        for alpha in ((0, 2) if M == 256 else (2,)):
            key = f"ch_{M}_{N}_{alpha}"
            a = alpha - 0.5
            plan = FastChebyshevTransform(N, M, a, a, -0.5, -0.5)
            served = get_lib().rfft_regs_launches()
            gout = g[key + "_gout"]
            out = torch.full(gout.shape, float('nan'), dtype=torch.float64, device='cuda')
            plan.backward(_t(g[key + "_cin"]), out, 1)
            assert np.allclose(out.cpu().numpy(), gout, rtol=1e-10, atol=1e-10 * np.abs(gout).max())
            cout = g[key + "_cout"]
            out = torch.full(cout.shape, float('nan'), dtype=torch.float64, device='cuda')
            plan.forward(_t(g[key + "_gin"]), out, 1)
            assert np.allclose(out.cpu().numpy(), cout, rtol=1e-11, atol=1e-11 * np.abs(cout).max())
            assert get_lib().rfft_regs_launches() == served + 2


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("s", [0, 1, -2])
def test_swsh_colatitude_transform_matches_reference(golden, tag, s):
    """T5 on the GPU: SWSHColatitudeTransform plugin vs the UNMODIFIED reference's plan on its own m_maps (swsh.npz)."""
    import torch
    from dedalus_b200.transforms import SWSHColatitudeTransform
    g = golden("swsh.npz")
    Nphi, Ntheta, Lmax, Gp, Gt, Cp, Ce = (int(v) for v in g[tag + "_meta"])
    plan = SWSHColatitudeTransform(Gt, Lmax, [tuple(r) for r in g[tag + "_m_maps"]], s)
    c = torch.zeros(g[f"{tag}_s{s}_cout"].shape, dtype=torch.float64, device='cuda')
    plan.forward(_t(g[f"{tag}_s{s}_gin"]), c, 2)
    assert np.allclose(c.cpu().numpy(), g[f"{tag}_s{s}_cout"], rtol=1e-12, atol=1e-13)
    gg = torch.full(g[f"{tag}_s{s}_gout"].shape, float('nan'), dtype=torch.float64, device='cuda')
    plan.backward(_t(g[f"{tag}_s{s}_cin"]), gg, 2)
    assert np.allclose(gg.cpu().numpy(), g[f"{tag}_s{s}_gout"], rtol=1e-12, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("s", [0, 2])
def test_swsh_round_trip_at_config4_size(s):
    """BASELINE config 4's resolution (Nphi = 512, Ntheta = 256 -> Lmax = 254, 3/2 dealiasing: 384 colatitude points): backward
    then forward is the identity on band-limited coefficients (l >= max(|m|, |s|)), and matches the CPU oracle for one m."""
    import torch
    from dedalus_b200.transforms import SWSHColatitudeTransform
    from oracle import transforms_oracle as T
    Lmax, Nt = 254, 384
    maps = [(m, 2 * m, 2 * m + 2, 2 * m, 2 * m + 2, m, Lmax + 1, 1) for m in range(Lmax + 1)]
    plan = SWSHColatitudeTransform(Nt, Lmax, maps, s)
    rng = np.random.default_rng(s)
    c = rng.standard_normal((3, 2 * (Lmax + 1), Lmax + 1))
    for m in range(Lmax + 1):
        c[:, 2 * m:2 * m + 2, :max(m, abs(s))] = 0
    g = torch.full((3, 2 * (Lmax + 1), Nt), float('nan'), dtype=torch.float64, device='cuda')
    plan.backward(_t(c), g, 2)
    back = torch.zeros(c.shape, dtype=torch.float64, device='cuda')
    plan.forward(g, back, 2)
    assert np.allclose(back.cpu().numpy(), c, rtol=1e-10, atol=1e-11)
    m = 37
    _, B = T.swsh_matrices(Nt, Lmax, m, s)
    ref = np.einsum('tk,oik->oit', B, c[:, 2 * m:2 * m + 2, m:])
    assert np.allclose(g.cpu().numpy()[:, 2 * m:2 * m + 2], ref, rtol=1e-9, atol=1e-10)       # two independent recurrences at degree 254
