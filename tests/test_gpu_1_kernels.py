"""Direct GPU checks, through the C ABI, of the kernels that the whole-solver tests only exercise indirectly: the grid-space
product programs (P1: core/arithmetic.py:246-251, 666-674, 855-866 in the reference), the transpose pack / unpack permutes
(X1: core/transposes.pyx:106-113, 211-246) and the dense matrix transform (T4: tools/array.py:104-129), against numpy."""
import numpy as np, pytest

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _dp(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("npts", [2 * 148 * 1024 + 2, 4098, 2])
def test_pointwise_pairs_matches_numpy(npts):
    """RB's stage program shape: 15 inputs, 4 outputs, out_j = -sum_i u_i * d_i f_j (plus a linear term and a constant-free
    single factor), bit-comparable up to the order of the three-term sums."""
    import torch
    from dedalus_b200.lib import get_lib, current_stream
    rng = np.random.default_rng(5)
    n_in, n_out = 15, 4
    x = rng.standard_normal((n_in, npts))
    terms, ptr = [], [0]
    for j in range(n_out):
        for i in range(3):
            terms.append((-1.0, i, 3 + 3 * j + i))
        if j == 1:
            terms.append((0.5, 7, -1))                     # single-factor term
        ptr.append(len(terms))
    rec = np.zeros(len(terms), dtype=np.dtype([('coef', '<f8'), ('a', '<i4'), ('b', '<i4')]))
    for t, (c, a, b) in enumerate(terms):
        rec[t] = (c, a, b)
    xin, out = _t(x), torch.full((n_out, npts), float('nan'), dtype=torch.float64, device='cuda')
    tp = _t(np.array(ptr, dtype=np.int32)); pr = _t(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy())
    get_lib().call("db_pointwise_pairs", _dp(xin), _dp(out), npts, n_in, n_out, _dp(tp), _dp(pr), current_stream())
    ref = np.zeros((n_out, npts))
    for j in range(n_out):
        for c, a, b in terms[ptr[j]:ptr[j + 1]]:
            ref[j] += c * x[a] * (x[b] if b >= 0 else 1.0)
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("npts", [100001, 1000])
def test_pointwise_general_program_matches_numpy(npts):
    import torch
    from dedalus_b200.lib import get_lib, current_stream
    rng = np.random.default_rng(2)
    x = rng.standard_normal((7, npts))
    term_ptr = np.array([0, 3, 5], dtype=np.int32); coef = np.array([-1.0, -1.0, -1.0, 2.0, 0.5])
    fac_ptr = np.array([0, 2, 4, 6, 9, 10], dtype=np.int32); fac = np.array([0, 3, 1, 4, 2, 5, 6, 6, 0, 1], dtype=np.int32)
    out = torch.full((2, npts), float('nan'), dtype=torch.float64, device='cuda')
    xin, tp, cf, fp, fc = _t(x), _t(term_ptr), _t(coef), _t(fac_ptr), _t(fac)
    get_lib().call("db_pointwise", _dp(xin), _dp(out), npts, 7, 2, _dp(tp), _dp(cf), _dp(fp), _dp(fc), len(fac), current_stream())
    o = out.cpu().numpy()
    assert np.allclose(o[0], -x[0] * x[3] - x[1] * x[4] - x[2] * x[5], rtol=1e-14, atol=1e-14)
    assert np.allclose(o[1], 2 * x[6] * x[6] * x[0] + 0.5 * x[1], rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("P,B,n1,n2,n3", [(4, 2, 8, 12, 5), (8, 3, 32, 48, 48), (2, 1, 6, 4, 1)])
def test_transpose_pack_unpack_is_the_global_permutation(P, B, n1, n2, n3):
    """pack -> exchange of per-peer blocks -> unpack == slicing the global array the other way (bitwise: a pure permutation),
    both directions, P emulated ranks on one GPU."""
    import torch
    from dedalus_b200.lib import get_lib, current_stream
    lib, st = get_lib(), current_stream()
    rng = np.random.default_rng(4)
    G = rng.standard_normal((B, n1, n2, n3))
    n1b, n2b = n1 // P, n2 // P
    sends = []
    for r in range(P):
        a = _t(G[:, r * n1b:(r + 1) * n1b]); s = torch.zeros(a.numel(), dtype=torch.float64, device='cuda')
        lib.call("db_transpose_pack", _dp(a), _dp(s), B, n1b, n2, n3, P, st)
        sends.append(s.view(P, -1))
    for r in range(P):
        recv = torch.stack([sends[src][r] for src in range(P)]).contiguous()
        out = torch.zeros((B, n1, n2b, n3), dtype=torch.float64, device='cuda')
        lib.call("db_transpose_unpack", _dp(recv), _dp(out), B, n1, n2b, n3, P, st)
        assert np.array_equal(out.cpu().numpy(), G[:, :, r * n2b:(r + 1) * n2b])
    sends = []
    for r in range(P):
        a = _t(G[:, :, r * n2b:(r + 1) * n2b]); s = torch.zeros(a.numel(), dtype=torch.float64, device='cuda')
        lib.call("db_transpose_pack_rev", _dp(a), _dp(s), B, n1, n2b, n3, P, st)
        sends.append(s.view(P, -1))
    for r in range(P):
        recv = torch.stack([sends[src][r] for src in range(P)]).contiguous()
        out = torch.zeros((B, n1b, n2, n3), dtype=torch.float64, device='cuda')
        lib.call("db_transpose_unpack_rev", _dp(recv), _dp(out), B, n1b, n2, n3, P, st)
        assert np.array_equal(out.cpu().numpy(), G[:, r * n1b:(r + 1) * n1b])


@pytest.mark.parametrize("m,n,outer,inner", [(70, 45, 3, 67), (384, 255, 2, 130), (5, 3, 1, 1),
                                             (70, 46, 3, 66), (384, 256, 2, 130), (255, 384, 2, 64), (8, 4, 1, 2), (200, 128, 1, 1000)])
def test_dense_matrix_transform_matches_numpy(m, n, outer, inner):
    """Odd n / inner: FMA-pipe tiles (k_mmt); even: FP64 tensor-core tiles (k_mmt_dmma, DMMA.8x8x4), including partial tiles."""
    import torch
    from dedalus_b200.lib import get_lib, current_stream
    rng = np.random.default_rng(9)
    mat = rng.standard_normal((m, n)); x = rng.standard_normal((outer, n, inner))
    out = torch.full((outer, m, inner), float('nan'), dtype=torch.float64, device='cuda')
    tm, tx = _t(mat), _t(x)
    get_lib().call("db_mmt_apply", _dp(tm), m, n, _dp(tx), _dp(out), outer, inner, current_stream())
    ref = np.einsum('ij,ojr->oir', mat, x)
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-13, atol=1e-13 * np.sqrt(n))


def test_absmax_and_residual_probe():
    import torch
    from dedalus_b200.lib import get_lib, current_stream
    v = np.random.default_rng(1).standard_normal(500001); v[1234] = -9.5
    res = torch.zeros(1, dtype=torch.float64, device='cuda'); tv = _t(v)
    get_lib().call("db_absmax", _dp(tv), v.size, _dp(res), current_stream())
    assert float(res.item()) == 9.5
