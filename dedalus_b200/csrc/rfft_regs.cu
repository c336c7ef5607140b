// Register-resident real Fourier transforms for the dealiased (3/2-rule) sizes on a strided axis (T1 hot path).
//
// The generic kernel in fft.cu runs log-many shared-memory passes and is bound by instruction issue, not HBM.  Here a
// line of n = 3 * Q * NB grid points is transformed in TWO register stages with ONE shared-memory exchange:
//   * two adjacent real lines (2p, 2p+1) are packed as one complex line z = g1 + i g2 ("two for one"), so every global
//     access is a 16-byte load / store covering both lines, and the real <-> half-complex pre / post processing of the
//     generic kernel collapses to four additions per mode;
//   * length n = NA * NB with NA = 3Q: stage A is a register DFT of size NA, stage B one of size NB; the twiddle between
//     them comes from the plan's exp(-2 pi i j / n) table;
//   * dealias padding is structural: coefficient modes k >= n/3 are zero, which is exactly the middle third of the
//     first radix-3 split of stage A (backward) / the discarded third of its outputs (forward), so that third of the
//     butterflies is never computed and the padded zeros / truncated modes are never touched in memory.
// One CTA = 16 lines (8 complex pairs, 128-byte rows) x the whole line; 48 KB of shared memory.
// Conventions as in fft.cu (reference core/transforms.py:469-509, 537-565): cos / -sin interleaved coefficients,
// forward scaled by 2/n (1/n for k = 0), sin(0) slot zero, Nyquist dropped.
#include "db_common.cuh"
#include <cstdlib>
#include <utility>
#include "tw96.inc"

namespace {

struct RegArgs {
    const double* in;
    double* out;
    const double* twn;      // exp(-2 pi i j / n), j < n, interleaved
    int64_t inner;          // stride between line elements (lines are adjacent along inner)
    int32_t M;              // coefficient rows (even, <= 2n/3)
    int32_t deriv;
    double kscale;
    // Blocked row addressing (pencil transposes without pack / unpack kernels, X1): with rpb > 0 row r of outer index o
    // lives at (r / rpb) * blk_stride + (o * rpb + r % rpb) * inner, i.e. the rows are grouped into the per-peer blocks
    // of an all-to-all send buffer (output side) or receive buffer (input side).  rpb = 0: plain (o * rows + r) * inner.
    int32_t in_rpb, out_rpb;
    int64_t in_blk_stride, out_blk_stride;
    // Peer exchange (X1 without a library collective): with out_peers > 0 output block b is not at out + b * out_blk_stride
    // but at out_blk_ptr[b] -- the receive buffer of GPU b, mapped into this process (NVLink peer memory): the transform's
    // own 16-byte stores ARE the all-to-all, overlapped tile by tile with its arithmetic.
    int32_t out_peers;
    double* out_blk_ptr[8];
};

__device__ __forceinline__ int64_t row_offset(int64_t o, int r, int rows, int rpb, int64_t blk_stride, int64_t inner)
{
    if (rpb == 0) return (o * rows + r) * inner;
    const int blk = r / rpb;
    return (int64_t)blk * blk_stride + (o * rpb + (r - blk * rpb)) * inner;
}

// address of output row r of outer index o (plain / blocked / blocked into peer memory)
__device__ __forceinline__ double* out_row_ptr(const RegArgs& a, double* gout, int64_t o, int r, int rows, int64_t inner)
{
    if (a.out_peers > 0) {
        const int blk = r / a.out_rpb;
        return a.out_blk_ptr[blk] + (gout - a.out) + (o * a.out_rpb + (r - blk * a.out_rpb)) * inner;
    }
    return gout + row_offset(o, r, rows, a.out_rpb, a.out_blk_stride, inner);
}

__device__ __forceinline__ double2 cadd2(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub2(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// a * exp(-/+ 2 pi i NUM / DEN) with the twiddle a compile-time constant (forward: minus sign, INV: plus)
template <int NUM, int DEN, bool INV> __device__ __forceinline__ double2 mul_tw(double2 a)
{
    static_assert(96 % DEN == 0, "twiddle denominators must divide 96");
    constexpr int j = ((NUM % DEN) * (96 / DEN)) % 96;
    if constexpr (j == 0) return a;
    else if constexpr (j == 24) return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
    else if constexpr (j == 48) return make_double2(-a.x, -a.y);
    else if constexpr (j == 72) return INV ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
    else {
        constexpr double c = tw96_cos(j);
        constexpr double s = INV ? tw96_sin(j) : -tw96_sin(j);
        return make_double2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

// natural-order in-register DFT of a power-of-two length (radix-2 decimation in time down to length 4 / 2)
template <int N, bool INV> struct DftP2 {
    template <int... K>
    static __device__ __forceinline__ void combine(double2 (&v)[N], const double2 (&e)[N / 2], const double2 (&o)[N / 2],
                                                   std::integer_sequence<int, K...>)
    {
        ((void)([&] {
            const double2 t = mul_tw<K, N, INV>(o[K]);
            v[K] = cadd2(e[K], t);
            v[K + N / 2] = csub2(e[K], t);
        }()), ...);
    }
    static __device__ __forceinline__ void run(double2 (&v)[N])
    {
        double2 e[N / 2], o[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        DftP2<N / 2, INV>::run(e);
        DftP2<N / 2, INV>::run(o);
        combine(v, e, o, std::make_integer_sequence<int, N / 2>{});
    }
};
template <bool INV> struct DftP2<4, INV> {
    static __device__ __forceinline__ void run(double2 (&v)[4])
    {
        const double2 a = cadd2(v[0], v[2]), b = csub2(v[0], v[2]), c = cadd2(v[1], v[3]), d = csub2(v[1], v[3]);
        const double2 jd = INV ? make_double2(-d.y, d.x) : make_double2(d.y, -d.x);
        v[0] = cadd2(a, c); v[2] = csub2(a, c);
        v[1] = cadd2(b, jd); v[3] = csub2(b, jd);
    }
};
template <bool INV> struct DftP2<2, INV> {
    static __device__ __forceinline__ void run(double2 (&v)[2])
    {
        const double2 a = v[0], b = v[1];
        v[0] = cadd2(a, b); v[1] = csub2(a, b);
    }
};

// u[b] *= exp(-/+ 2 pi i b C / NA) for b = 0..Q-1
template <int C, int NA, bool INV, int Q, int... B>
__device__ __forceinline__ void twiddle_row(double2 (&u)[Q], std::integer_sequence<int, B...>)
{
    ((void)(u[B] = mul_tw<(B * C) % NA, NA, INV>(u[B])), ...);
}

__device__ __forceinline__ double2 ldtwn(const double* __restrict__ tw, int j) { return reinterpret_cast<const double2*>(tw)[j]; }
// a * conj(w)
__device__ __forceinline__ double2 cmulc2(double2 a, double2 w) { return make_double2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }
__device__ __forceinline__ double2 cmul2(double2 a, double2 w) { return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

#define RR_P 8          // complex line pairs per CTA tile (16 real lines)

// Persistent CTAs walk the tiles (tile = 16 adjacent lines x the whole line length) with the NEXT tile's input rows in
// flight (16-byte cp.async, L1-bypassing) while the current tile is transformed: two shared-memory buffers of
// n x 128 bytes; the stage A -> stage B exchange reuses the buffer the current tile's input was staged in.
struct TileWalk {
    int64_t tiles_per_outer, total;
};

// stage `rows` rows of 16 doubles (row stride `inner`) into buf[row][16]
__device__ __forceinline__ void stage_rows(double* buf, const double* __restrict__ src, int64_t o, int rows, int rpb,
                                           int64_t blk_stride, int64_t inner)
{
    // src already points at the tile's first column; row r of outer index o is at row_offset(...)
    for (int idx = threadIdx.x; idx < rows * 8; idx += blockDim.x) {
        const int row = idx >> 3, seg = idx & 7;
        db_cp_async16(buf + row * 16 + seg * 2, src + row_offset(o, row, rows, rpb, blk_stride, inner) + seg * 2);
    }
}

// Thread layout.  Both register stages are split so that no thread holds more than max(Q, NB/2) complex values:
//   stage A (size NA = 3Q): thread (c, n2, p) computes the Q outputs k1 = c + 3 kb of the radix-3 residue c
//   stage B (size NB):      thread (h, k1, p) computes the NB/2 outputs k2 = 2 m + h (first radix-2 split, DIF)
// (forward: the same two splits in the opposite order).  3 NB P threads work in stage A and 6 Q P in stage B: 384 = 384
// at n = 384, at ~80 registers, so two CTAs (24 warps) are resident per SM and every warp works in both stages.
template <int Q, int NB> struct RegGeom {
    static constexpr int NA = 3 * Q, N = NA * NB, P = RR_P;
    static constexpr int TA = 3 * NB * P, TB = 6 * Q * P;
    static constexpr int THREADS = TA > TB ? TA : TB;
    static constexpr int MINB = (2 * N * P * 16 + 1024) * 2 <= 226 * 1024 ? 2 : 1;
};

template <bool INV, int N, int... M>
__device__ __forceinline__ void twiddle_odd_half(double2 (&o)[N], std::integer_sequence<int, M...>)
{
    ((void)(o[M] = mul_tw<M, 2 * N, INV>(o[M])), ...);
}

// u[b] = lo[b] + W3^(2c) hi[b], then * W_NA^(b c)      (backward stage A, residue C)
template <int C, int Q, bool INV>
__device__ __forceinline__ void residue_combine(double2 (&u)[Q], const double2 (&lo)[Q], const double2 (&hi)[Q])
{
#pragma unroll
    for (int b = 0; b < Q; ++b) u[b] = cadd2(lo[b], mul_tw<(2 * C) % 3, 3, INV>(hi[b]));
    twiddle_row<C, 3 * Q, INV>(u, std::make_integer_sequence<int, Q>{});
}

// ---------------------------------------------------------------------------------------------------------
// backward: coefficients (M rows) -> grid (n rows)
// ---------------------------------------------------------------------------------------------------------
template <int Q, int NB, bool DERIV>
__global__ void __launch_bounds__(RegGeom<Q, NB>::THREADS, RegGeom<Q, NB>::MINB)
k_rbwd_regs(RegArgs a, TileWalk tw_)
{
    using G = RegGeom<Q, NB>;
    constexpr int NA = G::NA, N = G::N, P = G::P, H = NB / 2;
    DB_SMEM(double2, sm);                                   // 2 buffers of [N][P] double2
    const int p = threadIdx.x & (P - 1), r = threadIdx.x >> 3;
    const int64_t inner = a.inner;
    const int M = a.M, Kmax = (a.M - 1) / 2;
    const int deriv = DERIV ? a.deriv : 0;
    // (i k kscale)^deriv = (ur + i ui) (k kscale)^deriv
    const double ur = ((deriv & 3) == 0) ? 1.0 : ((deriv & 3) == 2) ? -1.0 : 0.0;
    const double ui = ((deriv & 3) == 1) ? 1.0 : ((deriv & 3) == 3) ? -1.0 : 0.0;
    const double* __restrict__ tw = a.twn;
    auto stage_tile = [&](double2* buf, int64_t t) { const int64_t o = t / tw_.tiles_per_outer, xt = t - o * tw_.tiles_per_outer;
        stage_rows(reinterpret_cast<double*>(buf), a.in + xt * (2 * P), o, M, a.in_rpb, a.in_blk_stride, inner); };
    const bool in_a = r < 3 * NB, in_b = r < 2 * NA;
    const int n2 = r % NB, c = r / NB;                       // stage A role
    const int k1 = r % NA, h = r / NA;                       // stage B role
    int64_t t = blockIdx.x;
    int cur = 0;
    if (t < tw_.total) stage_tile(sm, t);
    db_cp_commit();
    for (; t < tw_.total; t += gridDim.x, cur ^= 1) {
        double2* in = sm + cur * (N * P);
        const int64_t tn = t + gridDim.x;
        if (tn < tw_.total) stage_tile(sm + (cur ^ 1) * (N * P), tn);
        db_cp_commit();
        db_cp_wait<1>();
        __syncthreads();
        double2 u[Q];
        if (in_a) {
            // Z_k = X1_k + i X2_k (k <= Kmax), Z_{n-k} = conj X1_k + i conj X2_k, X_k = (c_2k + i c_2k+1) / 2, X_0 = c_0
            auto spec = [&](int k, bool mirrored) -> double2 {
                if (k > Kmax) return make_double2(0.0, 0.0);
                const double2 R = in[(2 * k) * P + p];                      // (re line 1, re line 2)
                if (k == 0) return DERIV ? make_double2(0.0, 0.0) : R;
                const double2 I = in[(2 * k + 1) * P + p];                  // (im line 1, im line 2)
                if (DERIV) {
                    const double ks = a.kscale * k, ks2 = ks * ks;
                    const double f = 0.5 * ((deriv == 1) ? ks : (deriv == 2) ? ks2 : (deriv == 3) ? ks2 * ks : ks2 * ks2);
                    const double mr = ur * f, mi = ui * f;
                    const double2 R2 = make_double2(mr * R.x - mi * I.x, mr * R.y - mi * I.y);
                    const double2 I2 = make_double2(mi * R.x + mr * I.x, mi * R.y + mr * I.y);
                    return mirrored ? make_double2(R2.x + I2.y, R2.y - I2.x) : make_double2(R2.x - I2.y, I2.x + R2.y);
                }
                return mirrored ? make_double2(0.5 * (R.x + I.y), 0.5 * (R.y - I.x)) : make_double2(0.5 * (R.x - I.y), 0.5 * (I.x + R.y));
            };
            double2 lo[Q], hi[Q];
#pragma unroll
            for (int b = 0; b < Q; ++b) {
                lo[b] = spec(NB * b + n2, false);                  // Z[NB b + n2]
                hi[b] = spec(NB * (Q - b) - n2, true);             // Z[NB (2Q + b) + n2] = mirror of mode n - that
            }
            // stage A: size-NA DFT over n1 = Q a + b with the a = 1 third identically zero; this thread's residue c
            if (c == 0) residue_combine<0, Q, true>(u, lo, hi);
            else if (c == 1) residue_combine<1, Q, true>(u, lo, hi);
            else residue_combine<2, Q, true>(u, lo, hi);
        }
        __syncthreads();                                           // every input value is consumed: reuse `in` as exchange
        if (in_a) {
            DftP2<Q, true>::run(u);
            double2* dst = in + n2 * P + p;
#pragma unroll
            for (int kb = 0; kb < Q; ++kb) {
                const int kk = 3 * kb + c;                         // output k1
                dst[kk * NB * P] = cmulc2(u[kb], ldtwn(tw, n2 * kk));
            }
        }
        __syncthreads();
        if (in_b) {
            // stage B: size-NB DFT over n2 for fixed k1, outputs k2 = 2 m + h; grid index j = k1 + NA k2
            const double2* src = in + (k1 * NB) * P + p;
            double2 v[H];
            if (h == 0) {
#pragma unroll
                for (int m = 0; m < H; ++m) v[m] = cadd2(src[m * P], src[(m + H) * P]);
            } else {
#pragma unroll
                for (int m = 0; m < H; ++m) v[m] = csub2(src[m * P], src[(m + H) * P]);
                twiddle_odd_half<true>(v, std::make_integer_sequence<int, H>{});
            }
            DftP2<H, true>::run(v);
            const int64_t o = t / tw_.tiles_per_outer, xt = t - o * tw_.tiles_per_outer;
            double* __restrict__ gout = a.out + xt * (2 * P) + 2 * p;
#pragma unroll
            for (int m = 0; m < H; ++m)
                *reinterpret_cast<double2*>(out_row_ptr(a, gout, o, k1 + NA * (2 * m + h), N, inner)) = v[m];
        }
        __syncthreads();                                           // exchange reads done before the next prefetch lands here
    }
    db_cp_wait<0>();
}

// forward stage 2, residue C of the outputs: t[b] = sum_a' W3^(a' C) B[Q a' + b], times W_NA^(b C)
template <int C, int Q>
__device__ __forceinline__ void residue_gather(double2 (&u)[Q], const double2* src, int stride)
{
#pragma unroll
    for (int b = 0; b < Q; ++b) {
        const double2 b0 = src[b * stride], b1 = src[(Q + b) * stride], b2 = src[(2 * Q + b) * stride];
        u[b] = cadd2(b0, cadd2(mul_tw<C % 3, 3, false>(b1), mul_tw<(2 * C) % 3, 3, false>(b2)));
    }
    twiddle_row<C, 3 * Q, false>(u, std::make_integer_sequence<int, Q>{});
}

// ---------------------------------------------------------------------------------------------------------
// forward: grid (n rows) -> coefficients (M rows)
// ---------------------------------------------------------------------------------------------------------
template <int Q, int NB>
__global__ void __launch_bounds__(RegGeom<Q, NB>::THREADS, RegGeom<Q, NB>::MINB)
k_rfwd_regs(RegArgs a, TileWalk tw_)
{
    using G = RegGeom<Q, NB>;
    constexpr int NA = G::NA, N = G::N, P = G::P, H = NB / 2;
    DB_SMEM(double2, sm);                                   // 2 buffers of [N][P] double2
    const int p = threadIdx.x & (P - 1), r = threadIdx.x >> 3;
    const int64_t inner = a.inner;
    const int M = a.M, Kmax = (a.M - 1) / 2;
    const double* __restrict__ tw = a.twn;
    auto stage_tile = [&](double2* buf, int64_t t) { const int64_t o = t / tw_.tiles_per_outer, xt = t - o * tw_.tiles_per_outer;
        stage_rows(reinterpret_cast<double*>(buf), a.in + xt * (2 * P), o, N, a.in_rpb, a.in_blk_stride, inner); };
    const bool in_1 = r < 2 * NA, in_2 = r < 3 * NB;
    const int j1 = r % NA, h = r / NA;                       // stage 1 role: outputs k2 = 2 m + h of the size-NB DFT
    const int k2 = r % NB, c = r / NB;                       // stage 2 role: outputs k1 = c + 3 kb of the size-NA DFT
    int64_t t = blockIdx.x;
    int cur = 0;
    if (t < tw_.total) stage_tile(sm, t);
    db_cp_commit();
    for (; t < tw_.total; t += gridDim.x, cur ^= 1) {
        double2* in = sm + cur * (N * P);
        const int64_t tn = t + gridDim.x;
        if (tn < tw_.total) stage_tile(sm + (cur ^ 1) * (N * P), tn);
        db_cp_commit();
        db_cp_wait<1>();
        __syncthreads();
        double2 v[H];
        if (in_1) {
            // stage 1: size-NB DFT over j2 for fixed j1 (grid index j = j1 + NA j2), first radix-2 split (DIF)
            const double2* src = in + j1 * P + p;
            if (h == 0) {
#pragma unroll
                for (int m = 0; m < H; ++m) v[m] = cadd2(src[(NA * m) * P], src[(NA * (m + H)) * P]);
            } else {
#pragma unroll
                for (int m = 0; m < H; ++m) v[m] = csub2(src[(NA * m) * P], src[(NA * (m + H)) * P]);
                twiddle_odd_half<false>(v, std::make_integer_sequence<int, H>{});
            }
        }
        __syncthreads();                                           // inputs are consumed: reuse `in` as exchange
        if (in_1) {
            DftP2<H, false>::run(v);
            double2* dst = in + (j1 * NB + h) * P + p;
#pragma unroll
            for (int m = 0; m < H; ++m) {
                const int kk = 2 * m + h;                          // output k2
                dst[(2 * m) * P] = (kk == 0) ? v[m] : cmul2(v[m], ldtwn(tw, j1 * kk));
            }
        }
        __syncthreads();
        double2 z[Q];
        if (in_2) {
            // stage 2: size-NA DFT over j1 = Q a' + b for fixed k2; this thread's output residue c: k1 = c + 3 kb
            const double2* src = in + k2 * P + p;
            if (c == 0) residue_gather<0, Q>(z, src, NB * P);
            else if (c == 1) residue_gather<1, Q>(z, src, NB * P);
            else residue_gather<2, Q>(z, src, NB * P);
            DftP2<Q, false>::run(z);
        }
        __syncthreads();
        if (in_2) {
            // park the high third (k1 >= 2Q) as Zhi[k1 - 2Q][k2] for the mirror partners
            double2* dst = in + k2 * P + p;
#pragma unroll
            for (int kb = 0; kb < Q; ++kb) {
                const int kk = 3 * kb + c;
                if (kk >= 2 * Q) dst[(kk - 2 * Q) * NB * P] = z[kb];
            }
        }
        __syncthreads();
        if (in_2) {
            // Z_{n-k} for k = NB k1 + k2 sits at Zhi[(Q - 1 - k1), NB - k2] for k2 > 0, Zhi[Q - k1, 0] for k2 = 0
            const double sc = 1.0 / N;
            const int64_t o = t / tw_.tiles_per_outer, xt = t - o * tw_.tiles_per_outer;
            double* __restrict__ gout = a.out + xt * (2 * P) + 2 * p;
            const int k2p = (k2 == 0) ? 0 : NB - k2;
            const double2* src = in + k2p * P + p;
#pragma unroll
            for (int kb = 0; kb < Q; ++kb) {
                const int kk = 3 * kb + c;                         // k1
                const int k = NB * kk + k2;
                if (kk >= Q || k > Kmax) continue;
                double* row = out_row_ptr(a, gout, o, 2 * k, M, inner);
                double* row1 = out_row_ptr(a, gout, o, 2 * k + 1, M, inner);
                const double2 za = z[kb];
                if (k == 0) {
                    *reinterpret_cast<double2*>(row) = make_double2(za.x * sc, za.y * sc);
                    *reinterpret_cast<double2*>(row1) = make_double2(0.0, 0.0);
                } else {
                    const int kbp = (k2 == 0) ? Q - kk : Q - 1 - kk;
                    const double2 zb = src[kbp * NB * P];
                    *reinterpret_cast<double2*>(row) = make_double2((za.x + zb.x) * sc, (za.y + zb.y) * sc);
                    *reinterpret_cast<double2*>(row1) = make_double2((za.y - zb.y) * sc, (zb.x - za.x) * sc);
                }
            }
        }
        __syncthreads();
    }
    db_cp_wait<0>();
}

// =========================================================================================================
// Chebyshev (DCT-II / DCT-III) along CONTIGUOUS lines with the same two register stages (T2 on the benchmark path).
// A tile is 16 adjacent lines = one contiguous block of global memory; lines are staged in shared memory as
// X[line][n + 2] (the +2 spreads line pairs over the banks), two adjacent lines form one complex line, and the DCT runs
// through the full-length complex FFT of the even/odd reordered sequence (Makhoul): v[m] = g[2m] (m < n/2),
// v[m] = g[2(n-1-m)+1] otherwise;
//   backward: Z_k = w_k ((a1 + b2) + i (a2 - b1)), a = c^_k, b = c^_{n-k}, w_k = exp(i pi k / 2n), c^ = scaled, sign-
//             flipped coefficients (zero beyond the coefficient size);  v1 + i v2 = IDFT_n(Z)
//   forward : Z = DFT_n(v1 + i v2);  C1_k = s_k Re(q^k (Z_k + conj Z_{n-k})),  C2_k = s_k Re(q^k (Z_k - conj Z_{n-k}) / i),
//             q = conj w; truncation to the coefficient size and the banded ultraspherical conversion are applied while
//             the coefficient lines are still in shared memory.
// Conventions as fft.cu K_CHFWD / K_CHBWD (reference core/transforms.py:715-746, 771-902).
// =========================================================================================================
struct ChebArgs {
    const double* in;
    double* out;
    const double* twn;      // exp(-2 pi i j / n)
    const double* twq;      // exp(-i pi k / (2 n))
    const double* diags;    // forward: conversion diagonals [nd][M]
    const double* pre;      // backward: derivative diagonals [npre][M] applied first (npre <= 3), or null
    const double* sol2;     // backward: [2][M] reciprocal diagonal and second super-diagonal of the back-conversion, or null
    int64_t lines;
    int32_t M;              // coefficient size (even, <= n)
    int32_t nd;
    int32_t npre;
};

// Tile height.  The z kernels are one tile per CTA with full-CTA barriers between their phases, i.e. latency-bound (ncu,
// round 1: 2 CTAs / SM, long-scoreboard stalls, 1.1 - 1.9 TB/s): P = line pairs per tile is a template parameter so that
// smaller tiles (P = 4: 8 lines, 192 threads, 49 KB at n = 384 -> 4 CTAs / SM in different phases) can hide it.
template <int Q, int NB, int P> struct ChGeom {
    static constexpr int NA = 3 * Q, N = NA * NB, LINES = 2 * P;
    static constexpr int TA = 3 * NB * P, TB = 6 * Q * P;
    static constexpr int THREADS = TA > TB ? TA : TB;
    static constexpr int LOGP = P == 8 ? 3 : P == 4 ? 2 : P == 2 ? 1 : 0;
    static constexpr size_t SMEM = (size_t)LINES * (N + 2) * sizeof(double) + (size_t)N * P * sizeof(double2);
    static constexpr int MINB = (int)((226 * 1024) / (SMEM + 1024)) > 8 ? 8 : ((int)((226 * 1024) / (SMEM + 1024)) < 1 ? 1 : (int)((226 * 1024) / (SMEM + 1024)));
};

template <int Q, int NB, int P>
__global__ void __launch_bounds__((ChGeom<Q, NB, P>::THREADS), (ChGeom<Q, NB, P>::MINB))
k_chbwd_regs(ChebArgs a)
{
    using G = ChGeom<Q, NB, P>;
    constexpr int NA = G::NA, N = G::N, H = NB / 2, LX = N + 2, CH_LINES = G::LINES;
    DB_SMEM(double, X);                                      // [16][LX] doubles, then Y = [N][P] double2
    double2* Y = reinterpret_cast<double2*>(X + CH_LINES * LX);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int M = a.M;
    const int64_t l0 = (int64_t)blockIdx.x * CH_LINES;
    const int nl = (a.lines - l0 < CH_LINES) ? (int)(a.lines - l0) : CH_LINES;
    // ---- stage the coefficient lines (16-byte chunks); missing lines of a partial tile are zero
    {
        // one warp per line at a time, lanes along the line (16-byte chunks): no index division
        const int cpl = M / 2, lane = tid & 31, nw = nthreads >> 5;
        for (int l = tid >> 5; l < CH_LINES; l += nw) {
            double* dst = X + l * LX;
            if (l < nl) {
                const double* src = a.in + (l0 + l) * M;
                for (int ch = lane; ch < cpl; ch += 32) db_cp_async16(dst + 2 * ch, src + 2 * ch);
            } else {
                for (int ch = lane; ch < cpl; ch += 32) { dst[2 * ch] = 0.0; dst[2 * ch + 1] = 0.0; }
            }
        }
        db_cp_commit();
        db_cp_wait<0>();
    }
    __syncthreads();
    if (a.sol2 != nullptr) {
        // derivative + ultraspherical back-conversion in place on the staged lines (same suffix scan as k_band_scan2 in
        // fft.cu: x_i = r_i (t_i - u_i x_{i+2}), t = banded pre-apply of the coefficients), one warp per line
        const int lane = tid & 31, w = tid >> 5, nw = nthreads >> 5;
        const double* __restrict__ pre = a.pre;
        const double* __restrict__ sol = a.sol2;
        const int npre = a.npre, rounds = (M + 63) / 64;
        for (int l = w; l < CH_LINES; l += nw) {
            double* xl = X + l * LX;
            double carry0 = 0.0, carry1 = 0.0, nx0 = 0.0, nx1 = 0.0;     // nx: ORIGINAL c[i0+2], c[i0+3] of the round above
            for (int q = rounds - 1; q >= 0; --q) {
                const int i0 = 2 * (32 * q + lane);
                const double c0 = (i0 < M) ? xl[i0] : 0.0, c1 = (i0 + 1 < M) ? xl[i0 + 1] : 0.0;
                double c2 = __shfl_down_sync(0xffffffffu, c0, 1), c3 = __shfl_down_sync(0xffffffffu, c1, 1);
                if (lane == 31) { c2 = nx0; c3 = nx1; }
                nx0 = __shfl_sync(0xffffffffu, c0, 0); nx1 = __shfl_sync(0xffffffffu, c1, 0);
                double t0 = c0, t1 = c1;
                if (npre > 0) {
                    t0 = 0.0; t1 = 0.0;
                    if (i0 < M) {
                        t0 = pre[i0] * c0;
                        if (npre > 1) t0 = fma(pre[M + i0], c1, t0);
                        if (npre > 2) t0 = fma(pre[2 * M + i0], c2, t0);
                    }
                    if (i0 + 1 < M) {
                        t1 = pre[i0 + 1] * c1;
                        if (npre > 1) t1 = fma(pre[M + i0 + 1], c2, t1);
                        if (npre > 2) t1 = fma(pre[2 * M + i0 + 1], c3, t1);
                    }
                }
                double A0 = 0.0, B0 = 0.0, A1 = 0.0, B1 = 0.0;
                if (i0 < M) { const double rr = sol[i0]; B0 = rr * t0; A0 = (i0 + 2 < M) ? -rr * sol[M + i0] : 0.0; }
                if (i0 + 1 < M) { const double rr = sol[i0 + 1]; B1 = rr * t1; A1 = (i0 + 3 < M) ? -rr * sol[M + i0 + 1] : 0.0; }
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const double a0 = __shfl_down_sync(0xffffffffu, A0, off), b0 = __shfl_down_sync(0xffffffffu, B0, off);
                    const double a1 = __shfl_down_sync(0xffffffffu, A1, off), b1 = __shfl_down_sync(0xffffffffu, B1, off);
                    if (lane + off < 32) { B0 = fma(A0, b0, B0); A0 *= a0; B1 = fma(A1, b1, B1); A1 *= a1; }
                }
                const double x0 = fma(A0, carry0, B0), x1 = fma(A1, carry1, B1);
                if (i0 < M) xl[i0] = x0;
                if (i0 + 1 < M) xl[i0 + 1] = x1;
                carry0 = __shfl_sync(0xffffffffu, x0, 0); carry1 = __shfl_sync(0xffffffffu, x1, 0);
            }
        }
        __syncthreads();
    }
    // ---- spectrum of the packed line pairs: Y[k][p] = Z_k
    {
        const double c0 = 0.56418958354775628694807945156077;     // 1/sqrt(pi)
        const double c1 = 0.39894228040143267793994605993438;     // 1/sqrt(2 pi)
        // warp task = (block of 8 modes, half of the line pairs); lanes = 4 pairs x 8 modes (bank-friendly on both sides)
        const int lane = tid & 31, w = tid >> 5, nw = nthreads >> 5;
        constexpr int PL = P < 4 ? P : 4, KL = 32 / PL, PH = P / PL;       // lanes = PL pairs x KL modes; PH warp tasks per mode block
        const int kk = lane / PL;
        for (int wt = w; wt < PH * ((N + KL - 1) / KL); wt += nw) {
            const int k = (wt / PH) * KL + kk, pp = (lane % PL) + PL * (wt % PH);
            if (k >= N) continue;
            const double* x1 = X + (2 * pp) * LX;
            const double* x2 = x1 + LX;
            const int kb = N - k;
            const double sa = ((k == 0) ? c0 : c1) * ((k & 1) ? -1.0 : 1.0);
            const double sb = c1 * ((kb & 1) ? -1.0 : 1.0);
            const bool ha = k < M, hb = (k > 0) && (kb < M);
            const double a1 = ha ? x1[k] * sa : 0.0, a2 = ha ? x2[k] * sa : 0.0;
            const double b1 = hb ? x1[kb] * sb : 0.0, b2 = hb ? x2[kb] * sb : 0.0;
            const double2 wq = ldtwn(a.twq, k);              // exp(-i pi k / 2n): conj of w_k
            Y[k * P + pp] = cmulc2(make_double2(a1 + b2, a2 - b1), wq);
        }
    }
    __syncthreads();
    const int p = tid & (P - 1), r = tid >> G::LOGP;
    const bool in_a = r < 3 * NB, in_b = r < 2 * NA;
    const int n2 = r % NB, c = r / NB;
    const int k1 = r % NA, h = r / NA;
    const double* __restrict__ tw = a.twn;
    double2 u[Q];
    if (in_a) {
        const double2* src = Y + n2 * P + p;
        // stage A residue c: u[b] = sum_a W3^(a c) Z[NB (Q a + b) + n2], times W_NA^(b c)      (inverse signs)
        auto gather = [&](auto CC) {
            constexpr int C = decltype(CC)::value;
#pragma unroll
            for (int b = 0; b < Q; ++b) {
                const double2 z0 = src[(NB * b) * P], z1 = src[(NB * (Q + b)) * P], z2 = src[(NB * (2 * Q + b)) * P];
                u[b] = cadd2(z0, cadd2(mul_tw<C % 3, 3, true>(z1), mul_tw<(2 * C) % 3, 3, true>(z2)));
            }
            twiddle_row<C, NA, true>(u, std::make_integer_sequence<int, Q>{});
        };
        if (c == 0) gather(std::integral_constant<int, 0>{});
        else if (c == 1) gather(std::integral_constant<int, 1>{});
        else gather(std::integral_constant<int, 2>{});
    }
    __syncthreads();
    if (in_a) {
        DftP2<Q, true>::run(u);
        double2* dst = Y + n2 * P + p;
#pragma unroll
        for (int kb = 0; kb < Q; ++kb) {
            const int kq = 3 * kb + c;
            dst[kq * NB * P] = cmulc2(u[kb], ldtwn(tw, n2 * kq));
        }
    }
    __syncthreads();
    if (in_b) {
        const double2* src = Y + (k1 * NB) * P + p;
        double2 v[H];
        if (h == 0) {
#pragma unroll
            for (int m = 0; m < H; ++m) v[m] = cadd2(src[m * P], src[(m + H) * P]);
        } else {
#pragma unroll
            for (int m = 0; m < H; ++m) v[m] = csub2(src[m * P], src[(m + H) * P]);
            twiddle_odd_half<true>(v, std::make_integer_sequence<int, H>{});
        }
        DftP2<H, true>::run(v);
        // v[m] = (v1, v2)[j], j = k1 + NA (2 m + h): scatter to grid positions of the two lines
        double* o1 = X + (2 * p) * LX;
        double* o2 = o1 + LX;
#pragma unroll
        for (int m = 0; m < H; ++m) {
            const int j = k1 + NA * (2 * m + h);
            const int pos = (2 * j < N) ? 2 * j : 2 * (N - 1 - j) + 1;
            o1[pos] = v[m].x; o2[pos] = v[m].y;
        }
    }
    __syncthreads();
    // ---- store the grid lines (contiguous block)
    {
        const int cpl = N / 2, lane = tid & 31, nw = nthreads >> 5;
        for (int l = tid >> 5; l < nl; l += nw) {
            const double* src = X + l * LX;
            double* dst = a.out + (l0 + l) * N;
            for (int ch = lane; ch < cpl; ch += 32)
                *reinterpret_cast<double2*>(dst + 2 * ch) = *reinterpret_cast<const double2*>(src + 2 * ch);
        }
    }
}

template <int Q, int NB, int P>
__global__ void __launch_bounds__((ChGeom<Q, NB, P>::THREADS), (ChGeom<Q, NB, P>::MINB))
k_chfwd_regs(ChebArgs a)
{
    using G = ChGeom<Q, NB, P>;
    constexpr int NA = G::NA, N = G::N, H = NB / 2, LX = N + 2, CH_LINES = G::LINES;
    DB_SMEM(double, X);                                      // [16][LX] doubles, then Y = [N][P] double2
    double2* Y = reinterpret_cast<double2*>(X + CH_LINES * LX);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int M = a.M;
    const int64_t l0 = (int64_t)blockIdx.x * CH_LINES;
    const int nl = (a.lines - l0 < CH_LINES) ? (int)(a.lines - l0) : CH_LINES;
    {
        const int cpl = N / 2, lane = tid & 31, nw = nthreads >> 5;
        for (int l = tid >> 5; l < CH_LINES; l += nw) {
            double* dst = X + l * LX;
            if (l < nl) {
                const double* src = a.in + (l0 + l) * N;
                for (int ch = lane; ch < cpl; ch += 32) db_cp_async16(dst + 2 * ch, src + 2 * ch);
            } else {
                for (int ch = lane; ch < cpl; ch += 32) { dst[2 * ch] = 0.0; dst[2 * ch + 1] = 0.0; }
            }
        }
        db_cp_commit();
        db_cp_wait<0>();
    }
    __syncthreads();
    const int p = tid & (P - 1), r = tid >> G::LOGP;
    const bool in_1 = r < 2 * NA, in_2 = r < 3 * NB;
    const int j1 = r % NA, h = r / NA;
    const int k2 = r % NB, c = r / NB;
    const double* __restrict__ tw = a.twn;
    if (in_1) {
        // stage 1: v[j] = (g1, g2)[pos(j)], j = j1 + NA j2; size-NB DFT over j2, outputs k2 = 2 m + h
        const double* g1 = X + (2 * p) * LX;
        const double* g2 = g1 + LX;
        auto val = [&](int j2) -> double2 {
            const int j = j1 + NA * j2;
            const int pos = (2 * j < N) ? 2 * j : 2 * (N - 1 - j) + 1;
            return make_double2(g1[pos], g2[pos]);
        };
        double2 v[H];
        if (h == 0) {
#pragma unroll
            for (int m = 0; m < H; ++m) v[m] = cadd2(val(m), val(m + H));
        } else {
#pragma unroll
            for (int m = 0; m < H; ++m) v[m] = csub2(val(m), val(m + H));
            twiddle_odd_half<false>(v, std::make_integer_sequence<int, H>{});
        }
        DftP2<H, false>::run(v);
        double2* dst = Y + (j1 * NB + h) * P + p;
#pragma unroll
        for (int m = 0; m < H; ++m) {
            const int kq = 2 * m + h;
            dst[(2 * m) * P] = (kq == 0) ? v[m] : cmul2(v[m], ldtwn(tw, j1 * kq));
        }
    }
    __syncthreads();
    double2 z[Q];
    if (in_2) {
        const double2* src = Y + k2 * P + p;
        if (c == 0) residue_gather<0, Q>(z, src, NB * P);
        else if (c == 1) residue_gather<1, Q>(z, src, NB * P);
        else residue_gather<2, Q>(z, src, NB * P);
        DftP2<Q, false>::run(z);
    }
    __syncthreads();
    if (in_2) {
        // full spectrum in natural order: Y[k][p] = Z_k, k = NB k1 + k2, k1 = c + 3 kb
        double2* dst = Y + (k2 * P) + p;
#pragma unroll
        for (int kb = 0; kb < Q; ++kb) dst[(NB * (3 * kb + c)) * P] = z[kb];
    }
    __syncthreads();
    // ---- coefficients of both lines of every pair into X[line][k] (k < Kin), scaled, odd modes negated
    const int Kin = M < N ? M : N;
    {
        const double s0 = 0.5 / N * 1.7724538509055160272981674833411;      // sqrt(pi)/(2N)
        const double s1 = 1.0 / N * 1.2533141373155002512078826424055;      // sqrt(pi/2)/N
        const int lane = tid & 31, w = tid >> 5, nw = nthreads >> 5;
        constexpr int PL = P < 4 ? P : 4, KL = 32 / PL, PH = P / PL;
        const int kk = lane / PL;
        for (int wt = w; wt < PH * ((Kin + KL - 1) / KL); wt += nw) {
            const int k = (wt / PH) * KL + kk, pp = (lane % PL) + PL * (wt % PH);
            if (k >= Kin) continue;
            double* x1 = X + (2 * pp) * LX;
            double* x2 = x1 + LX;
            const double2 za = Y[k * P + pp];
            const double2 zb = Y[((k == 0) ? 0 : N - k) * P + pp];
            const double2 S = make_double2(za.x + zb.x, za.y - zb.y);            // Z_k + conj Z_{n-k}
            const double2 D = make_double2(za.y + zb.y, zb.x - za.x);            // (Z_k - conj Z_{n-k}) / i
            const double2 q = ldtwn(a.twq, k);
            const double sc = ((k == 0) ? s0 : s1) * ((k & 1) ? -1.0 : 1.0);
            x1[k] = sc * (q.x * S.x - q.y * S.y);
            x2[k] = sc * (q.x * D.x - q.y * D.y);
        }
    }
    __syncthreads();
    // ---- store with the banded conversion applied: out[i] = sum_d diag[d][i] cof[i + d]
    {
        // one warp per line at a time, two adjacent outputs per lane (16-byte stores).  All diagonal values of an output pair
        // are loaded BEFORE the multiply-adds (ncu, round 2: the former loop -- one dependent global load of diag[d][i] in
        // front of every FMA, runtime trip count -- took 67 % of this kernel's stall samples and 61 % of its instructions)
        constexpr int MAXD = 8;
        const int nd = a.nd, lane = tid & 31, nw = nthreads >> 5;
        const double* __restrict__ dg = a.diags;
        for (int l = tid >> 5; l < nl; l += nw) {
            const double* cf = X + l * LX;
            double* dst = a.out + (l0 + l) * M;
            for (int i = 2 * lane; i < M; i += 64) {
                double acc0 = 0.0, acc1 = 0.0;
                if (nd == 0) {
                    acc0 = (i < Kin) ? cf[i] : 0.0; acc1 = (i + 1 < Kin) ? cf[i + 1] : 0.0;
                } else if (nd <= MAXD) {
                    double2 dv[MAXD];
                    double cv[MAXD + 1];
#pragma unroll
                    for (int d = 0; d < MAXD; ++d)
                        dv[d] = (d < nd) ? __ldg(reinterpret_cast<const double2*>(dg + (int64_t)d * M + i)) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int d = 0; d <= MAXD; ++d) cv[d] = (d <= nd && i + d < Kin) ? cf[i + d] : 0.0;
#pragma unroll
                    for (int d = 0; d < MAXD; ++d) { acc0 = fma(dv[d].x, cv[d], acc0); acc1 = fma(dv[d].y, cv[d + 1], acc1); }
                } else {
                    for (int d = 0; d < nd; ++d) {
                        if (i + d < Kin) acc0 = fma(dg[(int64_t)d * M + i], cf[i + d], acc0);
                        if (i + 1 + d < Kin) acc1 = fma(dg[(int64_t)d * M + i + 1], cf[i + 1 + d], acc1);
                    }
                }
                *reinterpret_cast<double2*>(dst + i) = make_double2(acc0, acc1);
            }
        }
    }
}

template <int Q, int NB, int P>
int launch_cheb_p(bool fwd, const ChebArgs& a, void* stream)
{
    using G = ChGeom<Q, NB, P>;
    const int64_t blocks = (a.lines + G::LINES - 1) / G::LINES;
#ifndef DB_EMU
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_chbwd_regs<Q, NB, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_MAX_SMEM);
        cudaFuncSetAttribute(k_chfwd_regs<Q, NB, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_MAX_SMEM);
        cudaFuncSetAttribute(k_chbwd_regs<Q, NB, P>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_chfwd_regs<Q, NB, P>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr = true;
    }
#endif
    if (fwd) DB_LAUNCH((k_chfwd_regs<Q, NB, P>), dim3((unsigned)blocks), dim3(G::THREADS), G::SMEM, stream, a);
    else DB_LAUNCH((k_chbwd_regs<Q, NB, P>), dim3((unsigned)blocks), dim3(G::THREADS), G::SMEM, stream, a);
    return db_check_launch(fwd ? "cheb_forward(regs)" : "cheb_backward(regs)");
}

static int cheb_tile_pairs()
{
    static int p = 0;
    if (p == 0) { const char* e = getenv("DB_CHEB_P"); p = e ? atoi(e) : 4; if (p != 8 && p != 4 && p != 2) p = 4; }
    return p;
}

template <int Q, int NB>
int launch_cheb(bool fwd, const ChebArgs& a, void* stream)
{
    // small transforms keep the tall tile (their CTAs are small anyway); the tile height of the long ones is tunable
    if (3 * Q * NB < 192) return launch_cheb_p<Q, NB, 8>(fwd, a, stream);
    switch (cheb_tile_pairs()) {
        case 8: return launch_cheb_p<Q, NB, 8>(fwd, a, stream);
        case 2: return launch_cheb_p<Q, NB, 2>(fwd, a, stream);
        default: return launch_cheb_p<Q, NB, 4>(fwd, a, stream);
    }
}

static int regs_num_sms()
{
#ifdef DB_EMU
    return 1;                    // tests: several tiles per persistent CTA
#else
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
#endif
}

template <int Q, int NB>
int launch_regs(bool fwd, const RegArgs& a, int64_t outer, void* stream)
{
    constexpr int NA = 3 * Q, THREADS = RegGeom<Q, NB>::THREADS;
    const size_t bytes = (size_t)2 * NA * NB * RR_P * sizeof(double2);
    TileWalk w;
    w.tiles_per_outer = a.inner / (2 * RR_P);
    w.total = w.tiles_per_outer * outer;
    // persistent grid: as many CTAs as fit (2 per SM at n = 384: 96 KB each), never more than there are tiles
    int per_sm = (int)((size_t)(226 * 1024) / (bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t nblk = (int64_t)regs_num_sms() * per_sm;
    if (nblk > w.total) nblk = w.total;
    dim3 grid((unsigned)nblk);
#ifndef DB_EMU
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_rbwd_regs<Q, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_MAX_SMEM);
        cudaFuncSetAttribute(k_rbwd_regs<Q, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_MAX_SMEM);
        cudaFuncSetAttribute(k_rfwd_regs<Q, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_MAX_SMEM);
        cudaFuncSetAttribute(k_rbwd_regs<Q, NB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_rbwd_regs<Q, NB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(k_rfwd_regs<Q, NB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr = true;
    }
#endif
    if (fwd) DB_LAUNCH((k_rfwd_regs<Q, NB>), grid, dim3(THREADS), bytes, stream, a, w);
    else if (a.deriv > 0) DB_LAUNCH((k_rbwd_regs<Q, NB, true>), grid, dim3(THREADS), bytes, stream, a, w);
    else DB_LAUNCH((k_rbwd_regs<Q, NB, false>), grid, dim3(THREADS), bytes, stream, a, w);
    return db_check_launch(fwd ? "rfft_forward(regs)" : "rfft_backward(regs)");
}

static long long g_regs_launches = 0;
}  // namespace

extern "C" long long db_rfft_regs_launches(void) { return g_regs_launches; }

// Returns -1 if the register kernels do not cover this case (the caller falls back to the generic shared-memory
// kernel), otherwise the launch status.
int db_rfft_regs_try(bool fwd, const db_fft_plan* plan, const double* in, double* out, int64_t outer, int32_t n_coeff,
                     int64_t inner, int32_t deriv, double kscale, void* stream,
                     int32_t in_rpb, int64_t in_blk_stride, int32_t out_rpb, int64_t out_blk_stride,
                     int32_t out_peers, double* const* out_blk_ptr)
{
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("DB_FFT_REGS"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    if (!enabled) return -1;
    const int n = plan->n;
    if (!plan->half || plan->twn == nullptr) return -1;
    if (inner < 2 * RR_P || inner % (2 * RR_P) != 0 || deriv > 4) return -1;
    if (n_coeff % 2 != 0 || n_coeff < 2 || (int64_t)3 * n_coeff > (int64_t)2 * n) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return -1;
    RegArgs a;
    a.in = in; a.out = out; a.twn = plan->twn; a.inner = inner; a.M = n_coeff; a.deriv = deriv; a.kscale = kscale;
    a.in_rpb = in_rpb; a.in_blk_stride = in_blk_stride; a.out_rpb = out_rpb; a.out_blk_stride = out_blk_stride;
    a.out_peers = 0;
    for (int i = 0; i < 8; ++i) a.out_blk_ptr[i] = nullptr;
    if (out_peers > 0) {
        if (out_peers > 8 || out_rpb <= 0 || out_blk_ptr == nullptr) return -1;
        a.out_peers = out_peers;
        for (int i = 0; i < out_peers; ++i) {
            if (reinterpret_cast<uintptr_t>(out_blk_ptr[i]) & 15) return -1;
            a.out_blk_ptr[i] = out_blk_ptr[i];
        }
    }
    if (((in_rpb ? in_blk_stride : 0) | (out_rpb ? out_blk_stride : 0)) & 1) return -1;          // 16-byte alignment of every block
    ++g_regs_launches;
    switch (n) {
        case 384: return launch_regs<8, 16>(fwd, a, outer, stream);
        case 192: return launch_regs<4, 16>(fwd, a, outer, stream);
        case 96:  return launch_regs<4, 8>(fwd, a, outer, stream);
        case 48:  return launch_regs<2, 8>(fwd, a, outer, stream);
        case 24:  return launch_regs<2, 4>(fwd, a, outer, stream);
        case 768: return launch_regs<8, 32>(fwd, a, outer, stream);
        default: --g_regs_launches; return -1;
    }
}

// Chebyshev on contiguous lines (inner == 1): returns -1 if not covered (generic kernel in fft.cu takes over).
// backward: plain DCT-III, optionally preceded (in shared memory) by the derivative pre-apply + the back-substitution
// of a conversion with only the main and the second super-diagonal (`sol2`, compact storage as db_band_lines stride 2);
// forward: conversion diagonals fused into the store.
int db_cheb_regs_try(bool fwd, const db_fft_plan* plan, const double* in, double* out, int64_t lines, int32_t n_coeff,
                     const double* diags, int32_t nd, const double* pre, int32_t npre, const double* sol2, void* stream)
{
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("DB_CHEB_REGS"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    if (!enabled) return -1;
    const int n = plan->n;
    if (plan->twn == nullptr || plan->twq == nullptr) return -1;
    if (n_coeff % 2 != 0 || n_coeff < 2 || n_coeff > n || lines > 2147483647LL * 4) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return -1;
    ChebArgs a;
    a.in = in; a.out = out; a.twn = plan->twn; a.twq = plan->twq; a.diags = diags; a.lines = lines; a.M = n_coeff; a.nd = nd;
    a.pre = pre; a.npre = npre; a.sol2 = sol2;
    if (npre > 3 || (npre > 0 && sol2 == nullptr)) return -1;
    ++g_regs_launches;
    switch (n) {
        case 384: return launch_cheb<8, 16>(fwd, a, stream);
        case 192: return launch_cheb<4, 16>(fwd, a, stream);
        case 96:  return launch_cheb<4, 8>(fwd, a, stream);
        case 48:  return launch_cheb<2, 8>(fwd, a, stream);
        case 24:  return launch_cheb<2, 4>(fwd, a, stream);
        default: --g_regs_launches; return -1;
    }
}
