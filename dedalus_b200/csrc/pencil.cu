// Batched pencil-system kernels, structure-of-arrays, one thread per system.
//
// Reference behaviour replaced: the per-pencil Python loops of core/timesteppers.py:588-643 (gather, CSR
// mat-vecs, RHS axpy, SuperLU solve, scatter) and core/subsystems.py:340-371.
// B200 design: every batch is S structurally identical systems; thread s owns system s, arrays are
// [row or entry][system] so a warp's 32 loads of one entry are one 256-byte coalesced transaction and the
// integer program (shared by all systems) is read uniformly.  The LU values are stored in the exact order
// the triangular solves consume them, so the solve is a pure stream over the factors.
#include "db_common.cuh"

#define PB_THREADS 128

// ---------------------------------------------------------------------------------------------------------
// gather / scatter: tiled transposes between z-contiguous field lines and system-contiguous pencil vectors
// ---------------------------------------------------------------------------------------------------------
template <bool GATHER>
__global__ void k_pencil_move(const double* __restrict__ src, double* __restrict__ dst, int S, int ld,
                              const int64_t* __restrict__ line_base, const int32_t* __restrict__ line_kind,
                              const int32_t* __restrict__ line_ptr, const int32_t* __restrict__ line_pos,
                              const int64_t* __restrict__ sys_off, int ld_sys)
{
    // block (32, 8): tile of 32 modes x 32 systems of line blockIdx.z
    DB_SMEM(double, tile);                      // [32][33]
    const int q = blockIdx.z;
    const int len = line_ptr[q + 1] - line_ptr[q];
    const int m0 = blockIdx.y * 32;
    if (m0 >= len) return;
    const int s0 = blockIdx.x * 32;
    const int64_t base = line_base[q];
    const int64_t* so = sys_off + (int64_t)line_kind[q] * ld_sys;
    const int32_t* pos = line_pos + line_ptr[q];
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (GATHER) {
        for (int r = ty; r < 32; r += 8) {       // read arena: tx -> mode (contiguous)
            int s = s0 + r, m = m0 + tx;
            if (s < S && m < len) tile[r * 33 + tx] = src[base + so[s] + m];
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {       // write vec: tx -> system (contiguous)
            int m = m0 + r, s = s0 + tx;
            if (s < S && m < len) dst[(int64_t)pos[m] * ld + s] = tile[tx * 33 + r];
        }
    } else {
        for (int r = ty; r < 32; r += 8) {
            int m = m0 + r, s = s0 + tx;
            if (s < S && m < len) tile[tx * 33 + r] = src[(int64_t)pos[m] * ld + s];
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            int s = s0 + r, m = m0 + tx;
            if (s < S && m < len) dst[base + so[s] + m] = tile[r * 33 + tx];
        }
    }
}

extern "C" int db_pencil_gather(const double* arena, double* vec, int32_t S, int32_t ld, int32_t nlines, int32_t max_len,
                                const int64_t* line_base, const int32_t* line_kind, const int32_t* line_ptr, const int32_t* line_pos,
                                const int64_t* sys_off, int32_t ld_sys, void* stream)
{
    if (S <= 0 || nlines <= 0 || max_len <= 0) return 0;
    dim3 grid((S + 31) / 32, (max_len + 31) / 32, nlines), block(32, 8);
    DB_LAUNCH(k_pencil_move<true>, grid, block, 32 * 33 * sizeof(double), stream, arena, vec, S, ld, line_base, line_kind, line_ptr, line_pos, sys_off, ld_sys);
    return db_check_launch("pencil_gather");
}

extern "C" int db_pencil_scatter(const double* vec, double* arena, int32_t S, int32_t ld, int32_t nlines, int32_t max_len,
                                 const int64_t* line_base, const int32_t* line_kind, const int32_t* line_ptr, const int32_t* line_pos,
                                 const int64_t* sys_off, int32_t ld_sys, void* stream)
{
    if (S <= 0 || nlines <= 0 || max_len <= 0) return 0;
    dim3 grid((S + 31) / 32, (max_len + 31) / 32, nlines), block(32, 8);
    DB_LAUNCH(k_pencil_move<false>, grid, block, 32 * 33 * sizeof(double), stream, vec, arena, S, ld, line_base, line_kind, line_ptr, line_pos, sys_off, ld_sys);
    return db_check_launch("pencil_scatter");
}

// ---------------------------------------------------------------------------------------------------------
// template mat-vecs  y = (sum_mono mono_vals[mono][s] T_mono) x   for M and L in one pass over rows
// grid: (ceil(S/128), row chunks)
// ---------------------------------------------------------------------------------------------------------
#define MV_ROWS_PER_BLOCK 16
__global__ void k_pencil_matvec(int n, int S, int ld, const double* __restrict__ mono_vals, const double* __restrict__ x,
                                const int32_t* __restrict__ m_ptr, const int32_t* __restrict__ m_col, const int32_t* __restrict__ m_mono, const double* __restrict__ m_val, double* __restrict__ y_m,
                                const int32_t* __restrict__ l_ptr, const int32_t* __restrict__ l_col, const int32_t* __restrict__ l_mono, const double* __restrict__ l_val, double* __restrict__ y_l)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int r0 = blockIdx.y * MV_ROWS_PER_BLOCK;
    const int r1 = (r0 + MV_ROWS_PER_BLOCK < n) ? r0 + MV_ROWS_PER_BLOCK : n;
    for (int i = r0; i < r1; ++i) {
        if (y_m) {
            double acc = 0.0;
            for (int t = m_ptr[i]; t < m_ptr[i + 1]; ++t)
                acc = fma(m_val[t] * mono_vals[(int64_t)m_mono[t] * ld + s], x[(int64_t)m_col[t] * ld + s], acc);
            y_m[(int64_t)i * ld + s] = acc;
        }
        if (y_l) {
            double acc = 0.0;
            for (int t = l_ptr[i]; t < l_ptr[i + 1]; ++t)
                acc = fma(l_val[t] * mono_vals[(int64_t)l_mono[t] * ld + s], x[(int64_t)l_col[t] * ld + s], acc);
            y_l[(int64_t)i * ld + s] = acc;
        }
    }
}

extern "C" int db_pencil_matvec(int32_t n, int32_t S, int32_t ld, const double* mono_vals, const double* x,
                                const int32_t* m_ptr, const int32_t* m_col, const int32_t* m_mono, const double* m_val, double* y_m,
                                const int32_t* l_ptr, const int32_t* l_col, const int32_t* l_mono, const double* l_val, double* y_l,
                                void* stream)
{
    if (S <= 0 || n <= 0) return 0;
    dim3 grid((S + PB_THREADS - 1) / PB_THREADS, (n + MV_ROWS_PER_BLOCK - 1) / MV_ROWS_PER_BLOCK), block(PB_THREADS);
    DB_LAUNCH(k_pencil_matvec, grid, block, 0, stream, n, S, ld, mono_vals, x, m_ptr, m_col, m_mono, m_val, y_m, l_ptr, l_col, l_mono, l_val, y_l);
    return db_check_launch("pencil_matvec");
}

// ---------------------------------------------------------------------------------------------------------
// LHS assembly on the static fill pattern
// ---------------------------------------------------------------------------------------------------------
#define ASM_ENTRIES_PER_BLOCK 64
__global__ void k_pencil_assemble(double* __restrict__ lu, int n_entries, int S, int ld, const double* __restrict__ mono_vals,
                                  const int32_t* __restrict__ asm_ptr, const int32_t* __restrict__ asm_mono, const double* __restrict__ asm_val)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int e0 = blockIdx.y * ASM_ENTRIES_PER_BLOCK;
    const int e1 = (e0 + ASM_ENTRIES_PER_BLOCK < n_entries) ? e0 + ASM_ENTRIES_PER_BLOCK : n_entries;
    for (int e = e0; e < e1; ++e) {
        double v = 0.0;
        for (int t = asm_ptr[e]; t < asm_ptr[e + 1]; ++t)
            v = fma(asm_val[t], mono_vals[(int64_t)asm_mono[t] * ld + s], v);
        lu[(int64_t)e * ld + s] = v;
    }
}

extern "C" int db_pencil_assemble(double* lu, int32_t n_entries, int32_t S, int32_t ld, const double* mono_vals,
                                  const int32_t* asm_ptr, const int32_t* asm_mono, const double* asm_val, void* stream)
{
    if (S <= 0 || n_entries <= 0) return 0;
    dim3 grid((S + PB_THREADS - 1) / PB_THREADS, (n_entries + ASM_ENTRIES_PER_BLOCK - 1) / ASM_ENTRIES_PER_BLOCK), block(PB_THREADS);
    DB_LAUNCH(k_pencil_assemble, grid, block, 0, stream, lu, n_entries, S, ld, mono_vals, asm_ptr, asm_mono, asm_val);
    return db_check_launch("pencil_assemble");
}

// ---------------------------------------------------------------------------------------------------------
// numeric LU on the static schedule (no pivot search: the order was fixed by the host's joint threshold
// pivoting); reciprocal pivots are stored on the diagonal
// ---------------------------------------------------------------------------------------------------------
__global__ void k_pencil_factor(double* __restrict__ lu, int n, int S, int ld, const int32_t* __restrict__ diag_eid,
                                const int32_t* __restrict__ fl_ptr, const int32_t* __restrict__ fl_eid,
                                const int32_t* __restrict__ fu_ptr, const int32_t* __restrict__ fu_eid,
                                const int32_t* __restrict__ fd_eid, int32_t* __restrict__ info)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int64_t dp = 0;
    bool bad = false;
    for (int k = 0; k < n; ++k) {
        const int64_t d = (int64_t)diag_eid[k] * ld + s;
        const double piv = lu[d];
        if (!(fabs(piv) > 0.0) || !(fabs(piv) < 1e300)) bad = true;
        const double inv = 1.0 / piv;
        lu[d] = inv;
        const int u0 = fu_ptr[k], u1 = fu_ptr[k + 1];
        for (int a = fl_ptr[k]; a < fl_ptr[k + 1]; ++a) {
            const int64_t le = (int64_t)fl_eid[a] * ld + s;
            const double l = lu[le] * inv;
            lu[le] = l;
            for (int b = u0; b < u1; ++b, ++dp) {
                const int64_t de = (int64_t)fd_eid[dp] * ld + s;
                lu[de] = fma(-l, lu[(int64_t)fu_eid[b] * ld + s], lu[de]);
            }
        }
    }
    if (bad) atomicAdd(info, 1);
}

extern "C" int db_pencil_factor(double* lu, int32_t n, int32_t S, int32_t ld, const int32_t* diag_eid,
                                const int32_t* fl_ptr, const int32_t* fl_eid, const int32_t* fu_ptr, const int32_t* fu_eid,
                                const int32_t* fd_eid, int32_t* info, void* stream)
{
    if (S <= 0 || n <= 0) return 0;
    dim3 grid((S + 63) / 64), block(64);
    DB_LAUNCH(k_pencil_factor, grid, block, 0, stream, lu, n, S, ld, diag_eid, fl_ptr, fl_eid, fu_ptr, fu_eid, fd_eid, info);
    return db_check_launch("pencil_factor");
}

// ---------------------------------------------------------------------------------------------------------
// triangular solves streaming the factors; RHS linear combination fused into the load
// ---------------------------------------------------------------------------------------------------------
__global__ void k_pencil_solve(const double* __restrict__ lu, int n, int S, int ld,
                               const int32_t* __restrict__ fwd_ptr, const int32_t* __restrict__ fwd_col,
                               const int32_t* __restrict__ bwd_ptr, const int32_t* __restrict__ bwd_col,
                               db_lincomb rhs, double* __restrict__ x)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double* __restrict__ f = lu + s;      // stream pointer (entry-major)
    int64_t e = 0;
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j < rhs.nvec; ++j) acc = fma(rhs.coef[j], rhs.vec[j][(int64_t)i * ld + s], acc);
        const int t1 = fwd_ptr[i + 1];
        for (int t = fwd_ptr[i]; t < t1; ++t, ++e)
            acc = fma(-f[e * ld], x[(int64_t)fwd_col[t] * ld + s], acc);
        x[(int64_t)i * ld + s] = acc;
    }
    for (int p = 0; p < n; ++p) {
        const int i = n - 1 - p;
        const double inv = f[e * ld]; ++e;
        double acc = x[(int64_t)i * ld + s];
        const int t1 = bwd_ptr[p + 1];
        for (int t = bwd_ptr[p]; t < t1; ++t, ++e)
            acc = fma(-f[e * ld], x[(int64_t)bwd_col[t] * ld + s], acc);
        x[(int64_t)i * ld + s] = acc * inv;
    }
}

extern "C" int db_pencil_solve(const double* lu, int32_t n, int32_t S, int32_t ld,
                               const int32_t* fwd_ptr, const int32_t* fwd_col, const int32_t* bwd_ptr, const int32_t* bwd_col,
                               const db_lincomb* rhs, double* x, void* stream)
{
    if (S <= 0 || n <= 0) return 0;
    if (rhs->nvec < 0 || rhs->nvec > 16) { db_set_error("pencil_solve: nvec out of range"); return 1; }
    dim3 grid((S + 63) / 64), block(64);
    DB_LAUNCH(k_pencil_solve, grid, block, 0, stream, lu, n, S, ld, fwd_ptr, fwd_col, bwd_ptr, bwd_col, *rhs, x);
    return db_check_launch("pencil_solve");
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_lincomb(db_lincomb terms, double* __restrict__ out, int64_t count)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j < terms.nvec; ++j) acc = fma(terms.coef[j], terms.vec[j][i], acc);
        out[i] = acc;
    }
}

extern "C" int db_lincomb_apply(const db_lincomb* terms, double* out, int64_t count, void* stream)
{
    if (count <= 0) return 0;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    DB_LAUNCH(k_lincomb, dim3((unsigned)blocks), dim3(256), 0, stream, *terms, out, count);
    return db_check_launch("lincomb");
}
