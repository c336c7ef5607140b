"""Kernel logic of csrc/fft.cu executed under the test-only CPU emulation (tests/emu) against the reference
outputs in tests/golden/transforms.npz (ScipyRealFFT / RealFourierMMT / ScipyComplexFFT / ScipyFastChebyshev /
JacobiMMT dumped by make_golden.py).  The same comparisons run on the real GPU in test_gpu_0_transforms.py."""
import numpy as np, pytest, ctypes as C
from dedalus_b200 import jacobi
from emu import emu_lib as E


def _axis_view(shape, axis):
    outer = int(np.prod(shape[:axis], dtype=int)); inner = int(np.prod(shape[axis + 1:], dtype=int))
    return outer, inner


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (16, 21), (32, 48), (12, 18), (10, 15)])
def test_rfft(golden, M, N):
    g = golden("transforms.npz"); lib = E.emu(); plan = E.EmuPlan(N, 'real')
    for ref in ("matrix", "scipy"):
        cin, gout = g[f"rf_{ref}_{M}_{N}_cin"], g[f"rf_{ref}_{M}_{N}_gout"]
        outer, inner = _axis_view(cin.shape, 1)
        out = np.zeros_like(gout)
        lib.call("db_rfft_backward", plan.ref(), E.ptr(np.ascontiguousarray(cin)), E.ptr(out), outer, M, inner, 0, 0.0, None)
        assert np.allclose(out, gout, rtol=1e-13, atol=1e-13)
        gin, cout = g[f"rf_{ref}_{M}_{N}_gin"], g[f"rf_{ref}_{M}_{N}_cout"]
        out = np.zeros_like(cout)
        lib.call("db_rfft_forward", plan.ref(), E.ptr(np.ascontiguousarray(gin)), E.ptr(out), outer, M, inner, None)
        assert np.allclose(out, cout, rtol=1e-13, atol=1e-13)


def test_rfft_last_axis_and_derivative():
    lib = E.emu(); N, M = 24, 16; plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(5)
    c = rng.standard_normal((5, 3, M)); c[..., 1] = 0
    out = np.zeros((5, 3, N))
    kscale = 2 * np.pi / 4.0
    lib.call("db_rfft_backward", plan.ref(), E.ptr(c), E.ptr(out), 15, M, 1, 1, kscale, None)
    x = np.arange(N) / N * 4.0
    k = np.arange(M // 2) * kscale
    ref = np.zeros_like(out)
    for kk in range(1, (M - 1) // 2 + 1):
        a, b = c[..., 2 * kk], c[..., 2 * kk + 1]
        ref += (-a * k[kk])[..., None] * np.sin(k[kk] * x) + (-b * k[kk])[..., None] * np.cos(k[kk] * x)
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (15, 22), (12, 18)])
def test_cfft(golden, M, N):
    g = golden("transforms.npz"); lib = E.emu(); plan = E.EmuPlan(N, 'complex')
    for ref in ("matrix", "scipy"):
        cin, gout = g[f"cf_{ref}_{M}_{N}_cin"], g[f"cf_{ref}_{M}_{N}_gout"]
        outer, inner = _axis_view(cin.shape, 1)
        out = np.zeros_like(gout)
        lib.call("db_cfft_backward", plan.ref(), E.ptr(np.ascontiguousarray(cin)), E.ptr(out), outer, M, inner, 0, 0.0, None)
        assert np.allclose(out, gout, rtol=1e-13, atol=1e-13)
        gin, cout = g[f"cf_{ref}_{M}_{N}_gin"], g[f"cf_{ref}_{M}_{N}_cout"]
        out = np.zeros_like(cout)
        lib.call("db_cfft_forward", plan.ref(), E.ptr(np.ascontiguousarray(gin)), E.ptr(out), outer, M, inner, None)
        assert np.allclose(out, cout, rtol=1e-13, atol=1e-13)


def _conv_diags(M, N, a, b):
    """Upper diagonals of the (-1/2,-1/2)->(a,b) conversion restricted to the first M rows (row length M)."""
    K = max(M, N)
    Cm = jacobi.conversion_matrix(K, -0.5, -0.5, a, b).toarray()
    nd = int(round((a + 0.5) + (b + 0.5))) + 1
    d = np.zeros((nd, M))
    for k in range(nd):
        for i in range(M):
            if i + k < K:
                d[k, i] = Cm[i, i + k]
    return np.ascontiguousarray(d)


@pytest.mark.parametrize("M,N", [(16, 8), (16, 16), (16, 24), (15, 22), (15, 15), (32, 48)])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_chebyshev(golden, M, N, alpha):
    g = golden("transforms.npz"); lib = E.emu(); plan = E.EmuPlan(N, 'real')
    a = b = alpha - 0.5
    diags = _conv_diags(M, N, a, b) if alpha else None
    nd = diags.shape[0] if alpha else 0
    sdiags = None
    if alpha:
        sdiags = diags.copy(); sdiags[0] = 1.0 / sdiags[0]      # solve contract: reciprocal diagonal
    for ref in ("matrix", "scipy_dct"):
        key = f"ch_{ref}_{M}_{N}_{alpha}"
        cin, gout, gin, cout = g[key + "_cin"], g[key + "_gout"], g[key + "_gin"], g[key + "_cout"]
        outer, inner = _axis_view(cin.shape, 2)
        out = np.zeros_like(gout)
        lib.call("db_cheb_backward", plan.ref(), E.ptr(np.ascontiguousarray(cin)), E.ptr(out), outer, M, inner,
                 None, 0, E.ptr(sdiags), nd, None)
        assert np.allclose(out, gout, rtol=1e-12, atol=1e-12), (ref, "backward")
        out = np.zeros_like(cout)
        lib.call("db_cheb_forward", plan.ref(), E.ptr(np.ascontiguousarray(gin)), E.ptr(out), outer, M, inner,
                 E.ptr(diags), nd, None)
        assert np.allclose(out, cout, rtol=1e-12, atol=1e-12), (ref, "forward")


def test_chebyshev_strided_axis_matches_last_axis():
    lib = E.emu(); N, M = 24, 16; plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(3)
    c = rng.standard_normal((3, M, 37))
    out = np.zeros((3, N, 37))
    lib.call("db_cheb_backward", plan.ref(), E.ptr(c), E.ptr(out), 3, M, 37, None, 0, None, 0, None)
    c2 = np.ascontiguousarray(np.moveaxis(c, 1, 2)); out2 = np.zeros((3, 37, N))
    lib.call("db_cheb_backward", plan.ref(), E.ptr(c2), E.ptr(out2), 3 * 37, M, 1, None, 0, None, 0, None)
    assert np.allclose(np.moveaxis(out, 1, 2), out2, rtol=0, atol=1e-14)
    back = np.zeros_like(c)
    lib.call("db_cheb_forward", plan.ref(), E.ptr(out), E.ptr(back), 3, M, 37, None, 0, None)
    assert np.allclose(back, c, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("M,N", [(256, 384), (128, 192), (64, 96), (32, 48), (512, 768)])
def test_benchmark_line_lengths_use_static_passes(M, N):
    """The compile-time specialised FFT passes (nc = 192, 96, 48, 24->generic, 384) against the oracle transforms."""
    from oracle import transforms_oracle as T
    lib = E.emu(); plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(7)
    inner = 18                                       # > 8 lines so the tile width stays 16
    c = rng.standard_normal((2, M, inner)); c[:, 1, :] = 0
    g = np.zeros((2, N, inner))
    lib.call("db_rfft_backward", plan.ref(), E.ptr(c), E.ptr(g), 2, M, inner, 0, 0.0, None)
    assert np.allclose(g, T.rf_backward_fft(c, N, 1), rtol=1e-12, atol=1e-12)
    back = np.zeros_like(c)
    lib.call("db_rfft_forward", plan.ref(), E.ptr(g), E.ptr(back), 2, M, inner, None)
    assert np.allclose(back, c, rtol=1e-12, atol=1e-12)
    cz = rng.standard_normal((20, M)); gz = np.zeros((20, N))
    lib.call("db_cheb_backward", plan.ref(), E.ptr(cz), E.ptr(gz), 20, M, 1, None, 0, None, 0, None)
    assert np.allclose(gz, T.cheb_backward_fft(cz, N, 1), rtol=1e-11, atol=1e-11)
    bz = np.zeros_like(cz)
    lib.call("db_cheb_forward", plan.ref(), E.ptr(gz), E.ptr(bz), 20, M, 1, None, 0, None)
    assert np.allclose(bz, cz, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("M,N", [(256, 384), (128, 192), (64, 96), (32, 48), (16, 24), (512, 768), (100, 192), (2, 24)])
@pytest.mark.parametrize("deriv", [0, 1, 2, 3])
def test_register_resident_rfft(M, N, deriv):
    """csrc/rfft_regs.cu (two-stage register FFT, two real lines per complex line) against the oracle transforms:
    strided axis with the line count a multiple of 16, dealiased sizes, coefficient-space derivative fused."""
    from oracle import transforms_oracle as T
    lib = E.emu(); plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(11 + deriv)
    inner, outer = 48, 5                          # 15 tiles over at most 8 persistent CTAs: ragged multi-tile walks
    c = rng.standard_normal((outer, M, inner)); c[:, 1, :] = 0
    kscale = 2 * np.pi / 3.0
    cd = c.copy()
    for _ in range(deriv):                         # d/dx on cos/-sin pairs: (a, b) -> k (-b, a)
        k = (np.arange(M) // 2) * kscale
        a, b = cd[:, 0::2, :].copy(), cd[:, 1::2, :].copy()
        cd[:, 0::2, :] = -b * k[0::2, None]; cd[:, 1::2, :] = a * k[1::2, None]
    g = np.full((outer, N, inner), np.nan)
    served = lib.rfft_regs_launches()
    lib.call("db_rfft_backward", plan.ref(), E.ptr(c), E.ptr(g), outer, M, inner, deriv, kscale, None)
    assert lib.rfft_regs_launches() == served + 1      # not the generic fallback
    ref = T.rf_backward_fft(cd, N, 1)
    scale = max(1.0, np.abs(ref).max())
    assert np.allclose(g, ref, rtol=0, atol=1e-13 * scale)
    if deriv == 0:
        back = np.full_like(c, np.nan)
        lib.call("db_rfft_forward", plan.ref(), E.ptr(g), E.ptr(back), outer, M, inner, None)
        assert np.allclose(back, c, rtol=0, atol=1e-13)
        gr = rng.standard_normal((outer, N, inner))
        out = np.full_like(c, np.nan)
        lib.call("db_rfft_forward", plan.ref(), E.ptr(gr), E.ptr(out), outer, M, inner, None)
        assert np.allclose(out, T.rf_forward_fft(gr, M, 1), rtol=0, atol=1e-13)


@pytest.mark.parametrize("M", [256, 64, 50, 6])
@pytest.mark.parametrize("npre", [0, 2, 3])
def test_band_lines_warp_scan_matches_serial_recurrence(M, npre):
    """db_band_lines: compact even-diagonal storage (stride 2, warp suffix scan over affine maps) against the general
    one-thread-per-line recurrence and a dense numpy solve."""
    lib = E.emu()
    rng = np.random.default_rng(M + npre)
    lines = 11
    c = rng.standard_normal((lines, M))
    Cm = jacobi.conversion_matrix(M, -0.5, -0.5, 0.5, 0.5).toarray()          # diagonals 0 and 2
    full = np.zeros((3, M)); full[0] = np.diag(Cm); full[2, :M - 2] = np.diag(Cm, 2)
    pre = np.zeros((max(npre, 1), M))
    P = np.eye(M)
    if npre:
        P = np.zeros((M, M))
        for d in range(npre):
            pre[d, :M - d] = rng.standard_normal(M - d)
            P += np.diag(pre[d, :M - d], d)
    ref = np.linalg.solve(Cm, (P @ c.T)).T
    sol_full = full.copy(); sol_full[0] = 1.0 / full[0]
    sol_compact = np.ascontiguousarray(sol_full[[0, 2]])
    out_serial = np.full_like(c, np.nan); out_scan = np.full_like(c, np.nan)
    pre_ptr = E.ptr(pre) if npre else None
    lib.call("db_band_lines", E.ptr(c), E.ptr(out_serial), lines, M, pre_ptr, npre, E.ptr(sol_full), 3, 1, None)
    lib.call("db_band_lines", E.ptr(c), E.ptr(out_scan), lines, M, pre_ptr, npre, E.ptr(sol_compact), 2, 2, None)
    scale = np.abs(ref).max()
    assert np.allclose(out_serial, ref, rtol=0, atol=1e-12 * scale)
    assert np.allclose(out_scan, ref, rtol=0, atol=1e-12 * scale)


@pytest.mark.parametrize("M,N", [(256, 384), (128, 192), (64, 96), (32, 48), (16, 24), (200, 384), (384, 384)])
@pytest.mark.parametrize("alpha", [0, 2])
def test_register_resident_chebyshev(M, N, alpha):
    """csrc/rfft_regs.cu k_chbwd_regs / k_chfwd_regs (contiguous lines, two lines per complex FFT, conversion fused into
    the forward store) against the oracle transforms; 21 lines = one full tile + a partial one."""
    from oracle import transforms_oracle as T
    lib = E.emu(); plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(3 + alpha)
    lines = 21
    a = -0.5 + alpha
    c = rng.standard_normal((lines, M))
    g = np.full((lines, N), np.nan)
    served = lib.rfft_regs_launches()
    lib.call("db_cheb_backward", plan.ref(), E.ptr(c), E.ptr(g), lines, M, 1, None, 0, None, 0, None)
    assert lib.rfft_regs_launches() == served + 1
    ref = T.cheb_backward_fft(c, N, 1)                       # plain DCT-III of (-1/2, -1/2) coefficients
    assert np.allclose(g, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
    gr = rng.standard_normal((lines, N))
    out = np.full((lines, M), np.nan)
    if alpha:
        Cm = jacobi.conversion_matrix(max(M, N), -0.5, -0.5, a, a)
        nd = 2 * alpha + 1
        from dedalus_b200.transforms import banded_upper_diags
        dg = banded_upper_diags(Cm, M, nd)
        lib.call("db_cheb_forward", plan.ref(), E.ptr(gr), E.ptr(out), lines, M, 1, E.ptr(dg), nd, None)
    else:
        lib.call("db_cheb_forward", plan.ref(), E.ptr(gr), E.ptr(out), lines, M, 1, None, 0, None)
    assert lib.rfft_regs_launches() == served + 2
    ref = T.cheb_forward_fft(gr, M, 1, a, a)
    assert np.allclose(out, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("M,N", [(256, 384), (64, 96), (16, 24), (50, 96)])
@pytest.mark.parametrize("npre", [0, 2])
def test_chebyshev_backward_with_fused_derivative_scan(M, N, npre):
    """db_cheb_backward_scan == db_band_lines (stride-2 scan) followed by db_cheb_backward, bit for bit up to rounding."""
    lib = E.emu(); plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(M + npre)
    lines = 19
    c = rng.standard_normal((lines, M))
    Cm = jacobi.conversion_matrix(M, -0.5, -0.5, 0.5, 0.5).toarray()
    sol2 = np.zeros((2, M)); sol2[0] = 1.0 / np.diag(Cm); sol2[1, :M - 2] = np.diag(Cm, 2)
    pre = np.zeros((max(npre, 1), M))
    for d in range(npre):
        pre[d, :M - d] = rng.standard_normal(M - d)
    pre_ptr = E.ptr(pre) if npre else None
    tmp = np.full_like(c, np.nan); ref = np.full((lines, N), np.nan); out = np.full((lines, N), np.nan)
    lib.call("db_band_lines", E.ptr(c), E.ptr(tmp), lines, M, pre_ptr, npre, E.ptr(sol2), 2, 2, None)
    lib.call("db_cheb_backward", plan.ref(), E.ptr(tmp), E.ptr(ref), lines, M, 1, None, 0, None, 0, None)
    assert lib.call_optional("db_cheb_backward_scan", plan.ref(), E.ptr(c), E.ptr(out), lines, M, pre_ptr, npre, E.ptr(sol2), None)
    assert np.allclose(out, ref, rtol=0, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("M,N,P", [(16, 24, 2), (64, 96, 4), (256, 384, 8)])
def test_blocked_row_addressing_equals_pack_transform_unpack(M, N, P):
    """db_rfft_*_blocked: writing / reading the per-peer blocks of an all-to-all buffer directly gives what the plain
    transform gives after / before the pack and unpack copies (reference: FFTWTranspose buffer copies,
    core/transposes.pyx:146-246).  Two fields share the buffers (blk_stride spans both)."""
    lib = E.emu(); plan = E.EmuPlan(N, 'real')
    rng = np.random.default_rng(P)
    outer, inner, nf = 3, 32, 2
    kscale = 1.3

    def blocked(a, rows, rpb):
        # a: (nf, outer, rows, inner) -> (rows // rpb, nf, outer, rpb, inner)
        return np.ascontiguousarray(a.reshape(nf, outer, rows // rpb, rpb, inner).transpose(2, 0, 1, 3, 4))

    c = rng.standard_normal((nf, outer, M, inner)); c[:, :, 1, :] = 0
    g_ref = np.zeros((nf, outer, N, inner))
    for f in range(nf):
        lib.call("db_rfft_backward", plan.ref(), E.ptr(np.ascontiguousarray(c[f])), E.ptr(g_ref[f]), outer, M, inner, 1, kscale, None)
    # backward: blocked input (receive buffer of M / P rows per peer) and blocked output (send buffer of N / P rows per peer)
    cin = blocked(c, M, M // P); gout = np.full((P, nf, outer, N // P, inner), np.nan)
    per_field_in, per_field_out = outer * (M // P) * inner, outer * (N // P) * inner
    for f in range(nf):
        assert lib.call_optional("db_rfft_backward_blocked", plan.ref(), E.ptr(cin.reshape(-1)[f * per_field_in:]), E.ptr(gout.reshape(-1)[f * per_field_out:]),
                                 outer, M, inner, 1, kscale, M // P, nf * per_field_in, N // P, nf * per_field_out, None)
    assert np.array_equal(gout, blocked(g_ref, N, N // P))
    # forward likewise
    g = rng.standard_normal((nf, outer, N, inner))
    c_ref = np.zeros((nf, outer, M, inner))
    for f in range(nf):
        lib.call("db_rfft_forward", plan.ref(), E.ptr(np.ascontiguousarray(g[f])), E.ptr(c_ref[f]), outer, M, inner, None)
    gin = blocked(g, N, N // P); cout = np.full((P, nf, outer, M // P, inner), np.nan)
    for f in range(nf):
        assert lib.call_optional("db_rfft_forward_blocked", plan.ref(), E.ptr(gin.reshape(-1)[f * per_field_out:]), E.ptr(cout.reshape(-1)[f * per_field_in:]),
                                 outer, M, inner, N // P, nf * per_field_out, M // P, nf * per_field_in, None)
    assert np.array_equal(cout, blocked(c_ref, M, M // P))


def test_register_kernels_match_reference_at_benchmark_lengths(golden):
    """The UNMODIFIED reference's own transforms (ScipyRealFFT, ScipyFastChebyshev; tests/golden/transforms_bench.npz) at the
    256 -> 384 and 128 -> 192 line lengths, in the layouts the register-resident kernels take."""
    from dedalus_b200.transforms import banded_upper_diags
    g = golden("transforms_bench.npz"); lib = E.emu()
    for (M, N) in [(256, 384), (128, 192)]:
        plan = E.EmuPlan(N, 'real')
        cin, gout = g[f"rf_{M}_{N}_cin"], g[f"rf_{M}_{N}_gout"]
        served = lib.rfft_regs_launches()
        out = np.full_like(gout, np.nan)
        lib.call("db_rfft_backward", plan.ref(), E.ptr(np.ascontiguousarray(cin)), E.ptr(out), cin.shape[0], M, cin.shape[2], 0, 0.0, None)
        assert np.allclose(out, gout, rtol=1e-12, atol=1e-12)
        gin, cout = g[f"rf_{M}_{N}_gin"], g[f"rf_{M}_{N}_cout"]
        out = np.full_like(cout, np.nan)
        lib.call("db_rfft_forward", plan.ref(), E.ptr(np.ascontiguousarray(gin)), E.ptr(out), gin.shape[0], M, gin.shape[2], None)
        assert np.allclose(out, cout, rtol=1e-12, atol=1e-12)
        assert lib.rfft_regs_launches() == served + 2
        for alpha in ((0, 2) if M == 256 else (2,)):
            key = f"ch_{M}_{N}_{alpha}"
            a = alpha - 0.5
            cin, gout, gin, cout = (g[key + s] for s in ("_cin", "_gout", "_gin", "_cout"))
            lines = cin.shape[0]
            # backward from the (a, a) basis = back-conversion to Chebyshev-T coefficients (host, exact banded solve), then DCT-III
            served = lib.rfft_regs_launches()
            if alpha:
                Cm = jacobi.conversion_matrix(M, -0.5, -0.5, a, a).toarray()
                c0 = np.linalg.solve(Cm, cin.T).T
            else:
                c0 = cin
            out = np.full_like(gout, np.nan)
            lib.call("db_cheb_backward", plan.ref(), E.ptr(np.ascontiguousarray(c0)), E.ptr(out), lines, M, 1, None, 0, None, 0, None)
            assert np.allclose(out, gout, rtol=1e-10, atol=1e-10 * np.abs(gout).max())
            out = np.full_like(cout, np.nan)
            if alpha:
                nd = 2 * alpha + 1
                dg = banded_upper_diags(jacobi.conversion_matrix(max(M, N), -0.5, -0.5, a, a), M, nd)
                lib.call("db_cheb_forward", plan.ref(), E.ptr(np.ascontiguousarray(gin)), E.ptr(out), lines, M, 1, E.ptr(dg), nd, None)
            else:
                lib.call("db_cheb_forward", plan.ref(), E.ptr(np.ascontiguousarray(gin)), E.ptr(out), lines, M, 1, None, 0, None)
            assert np.allclose(out, cout, rtol=1e-11, atol=1e-11 * np.abs(cout).max())
            assert lib.rfft_regs_launches() == served + 2


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("s", [0, 1, -1, 2, -2])
def test_swsh_colatitude_transform_matches_reference(golden, tag, s):
    """T5: SWSHColatitudeTransform plugin (ragged per-m matrices, folded triangular packing) vs the UNMODIFIED reference's own
    plan on its own m_maps (tests/golden/swsh.npz), forward and backward, and the per-m matrices vs the reference's."""
    import torch
    from dedalus_b200.transforms import SWSHColatitudeTransform
    E.install()
    try:
        g = golden("swsh.npz")
        Nphi, Ntheta, Lmax, Gp, Gt, Cp, Ce = (int(v) for v in g[tag + "_meta"])
        plan = SWSHColatitudeTransform(Gt, Lmax, [tuple(r) for r in g[tag + "_m_maps"]], s)
        if s in (0, 2):
            for m in (0, 3):
                F, B = plan.matrices(m)
                assert np.allclose(F, g[f"{tag}_s{s}_m{m}_fwdmat"], rtol=1e-12, atol=1e-13)
                assert np.allclose(B, g[f"{tag}_s{s}_m{m}_bwdmat"], rtol=1e-12, atol=1e-13)
        gin = torch.from_numpy(np.ascontiguousarray(g[f"{tag}_s{s}_gin"]))
        c = torch.zeros(g[f"{tag}_s{s}_cout"].shape, dtype=torch.float64)
        plan.forward(gin, c, 2)
        assert np.allclose(c.numpy(), g[f"{tag}_s{s}_cout"], rtol=1e-12, atol=1e-13)
        cin = torch.from_numpy(np.ascontiguousarray(g[f"{tag}_s{s}_cin"]))
        gg = torch.full(g[f"{tag}_s{s}_gout"].shape, float('nan'), dtype=torch.float64)
        plan.backward(cin, gg, 2)
        assert np.allclose(gg.numpy(), g[f"{tag}_s{s}_gout"], rtol=1e-12, atol=1e-12, equal_nan=True)
    finally:
        E.uninstall()
