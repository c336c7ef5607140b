// Batches of DENSE systems with many right-hand sides: the per-degree pencil systems of spherical-shell problems (one matrix
// per l, shared by every azimuthal wavenumber m <= l: reference core/subsystems.py:272-274 Subproblem.shape = (n, n_subsystems);
// libraries/matsolvers.py:126-183 SuperLU-factors each).  The tau method leaves them as banded blocks bordered by dense boundary
// rows and tau columns; with n = 5 Nr + O(10) and ~l columns per system the solves are a small-matrix, many-column problem, so
// the matrices are kept dense, row-major, all systems padded to the same n (identity rows for the padding):
//     factor :  P A = L U in place, partial pivoting, one CTA per system (right-looking, rank-1 updates across the CTA)
//     solve  :  x = A^{-1} (sum_k coef_k v_k): one THREAD per right-hand-side column, the factor entries are broadcast loads
//               shared by the 32 columns of a warp, the columns' running values stay coalesced in x
//     matvec :  y = A x for one or two matrices (M.X, L.X)
// Vectors of system s: element (i, r) at vec_off + i * ncols + r  (ncols = right-hand sides of that system).
#include "db_common.cuh"
#include <cstdlib>

#define DN_THREADS 256

// ---------------------------------------------------------------------------------------------------------
// out = a0 * M + b0 * L  (all systems, contiguous [nsys][n][n])
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dense_combine(int64_t total, double a0, const double* __restrict__ m, double b0, const double* __restrict__ l, double* __restrict__ out)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = a0 * m[e] + b0 * l[e];
}

extern "C" int db_dense_combine(int32_t nsys, int32_t n, double a0, const double* m, double b0, const double* l, double* out, void* stream)
{
    const int64_t total = (int64_t)nsys * n * n;
    if (total <= 0) return 0;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    DB_LAUNCH(k_dense_combine, dim3((unsigned)blocks), dim3(256), 0, stream, total, a0, m, b0, l, out);
    return db_check_launch("dense_combine");
}

// ---------------------------------------------------------------------------------------------------------
// LU with partial pivoting, one CTA per system.  ipiv[s][k] = row interchanged with row k; info[s] = zero / non-finite pivots.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DN_THREADS)
k_dense_factor(int n, double* __restrict__ a_all, int32_t* __restrict__ ipiv_all, int32_t* __restrict__ info)
{
    DB_SMEM(double, smem);                              // [DN_THREADS] candidate values, then the pivot value, then the rows
    double* red_v = smem;
    double& s_pivot = smem[DN_THREADS];
    int* red_i = reinterpret_cast<int*>(smem + DN_THREADS + 1);
    int& s_prow = red_i[DN_THREADS];
    double* __restrict__ A = a_all + (int64_t)blockIdx.x * n * n;
    int32_t* __restrict__ ipiv = ipiv_all + (int64_t)blockIdx.x * n;
    const int tid = threadIdx.x;
    int bad = 0;
    for (int k = 0; k < n; ++k) {
        // pivot search down column k
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < n; i += DN_THREADS) {
            const double v = fabs(A[(int64_t)i * n + k]);
            if (v > best) { best = v; bi = i; }
        }
        red_v[tid] = best; red_i[tid] = bi;
        __syncthreads();
        for (int s = DN_THREADS / 2; s > 0; s >>= 1) {
            if (tid < s) {
                const double ov = red_v[tid + s]; const int oi = red_i[tid + s];
                if (ov > red_v[tid] || (ov == red_v[tid] && oi < red_i[tid])) { red_v[tid] = ov; red_i[tid] = oi; }
            }
            __syncthreads();
        }
        if (tid == 0) { s_prow = red_i[0]; s_pivot = red_v[0]; ipiv[k] = red_i[0]; }
        __syncthreads();
        const int p = s_prow;
        const double pv = s_pivot;
        if (!(pv > 0.0) || !(pv - pv == 0.0)) { ++bad; __syncthreads(); continue; }
        if (p != k)
            for (int j = tid; j < n; j += DN_THREADS) {
                const double t = A[(int64_t)k * n + j]; A[(int64_t)k * n + j] = A[(int64_t)p * n + j]; A[(int64_t)p * n + j] = t;
            }
        __syncthreads();
        const double rp = 1.0 / A[(int64_t)k * n + k];
        __syncthreads();
        for (int i = k + 1 + tid; i < n; i += DN_THREADS) A[(int64_t)i * n + k] *= rp;
        __syncthreads();
        // trailing update: A[i][j] -= A[i][k] * A[k][j], i, j > k; threads sweep rows of the trailing block (j fastest: coalesced)
        const int m = n - k - 1;
        for (int64_t e = tid; e < (int64_t)m * m; e += DN_THREADS) {
            const int i = k + 1 + (int)(e / m), j = k + 1 + (int)(e % m);
            A[(int64_t)i * n + j] = fma(-A[(int64_t)i * n + k], A[(int64_t)k * n + j], A[(int64_t)i * n + j]);
        }
        __syncthreads();
    }
    if (tid == 0) info[blockIdx.x] = bad;
}

extern "C" int db_dense_factor(int32_t nsys, int32_t n, double* a, int32_t* ipiv, int32_t* info, void* stream)
{
    if (nsys <= 0 || n <= 0) return 0;
    const size_t smem = (DN_THREADS + 1) * sizeof(double) + (DN_THREADS + 1) * sizeof(int);
    DB_LAUNCH(k_dense_factor, dim3((unsigned)nsys), dim3(DN_THREADS), smem, stream, n, a, ipiv, info);
    return db_check_launch("dense_factor");
}

// ---------------------------------------------------------------------------------------------------------
// Solve for all columns: thread = one column of one system (CTA = 32 columns).  b = P (sum_k coef_k v_k) is built in x, then
// forward (unit lower) and backward sweeps run in place; rows of L / U are the same for every lane (broadcast loads).
// ---------------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(32)
k_dense_solve(const db_dense_sys* __restrict__ sys, int n, const double* __restrict__ lu_all, const int32_t* __restrict__ ipiv_all,
              db_veccomb rhs, double* __restrict__ x_all)
{
    // SMEM: the 32 columns of the CTA live in shared memory ([n][32], conflict-free: lane = column) for both sweeps, so the only
    // global traffic of the sweeps is the factor rows, read once per CTA as broadcast loads; otherwise the columns stay in x.
    DB_SMEM(double, xs);
    const db_dense_sys S = sys[blockIdx.x];
    const int lane = threadIdx.x;
    const int r = blockIdx.y * 32 + lane;
    if (blockIdx.y * 32 >= S.ncols) return;
    const bool live = r < S.ncols;
    const double* __restrict__ LU = lu_all + (int64_t)blockIdx.x * n * n;
    const int32_t* __restrict__ ipiv = ipiv_all + (int64_t)blockIdx.x * n;
    double* __restrict__ xg = x_all + S.vec_off + (live ? r : 0);
    const int64_t ldg = S.ncols;
    double* __restrict__ x = SMEM ? xs + lane : xg;
    const int64_t ld = SMEM ? 32 : ldg;
    if (!SMEM && !live) return;
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        if (live)
            for (int k = 0; k < rhs.nvec; ++k) acc = fma(rhs.coef[k], rhs.vec[k][S.vec_off + (int64_t)i * ldg + r], acc);
        x[(int64_t)i * ld] = acc;
    }
    for (int k = 0; k < n; ++k) {
        const int p = ipiv[k];
        if (p != k) { const double t = x[(int64_t)k * ld]; x[(int64_t)k * ld] = x[(int64_t)p * ld]; x[(int64_t)p * ld] = t; }
    }
    for (int i = 1; i < n; ++i) {
        const double* __restrict__ row = LU + (int64_t)i * n;
        double a0 = x[(int64_t)i * ld], a1 = 0.0, a2 = 0.0, a3 = 0.0;           // four chains: the FMA latency is not the bound
        int j = 0;
        for (; j + 4 <= i; j += 4) {
            a0 = fma(-row[j], x[(int64_t)j * ld], a0);
            a1 = fma(-row[j + 1], x[(int64_t)(j + 1) * ld], a1);
            a2 = fma(-row[j + 2], x[(int64_t)(j + 2) * ld], a2);
            a3 = fma(-row[j + 3], x[(int64_t)(j + 3) * ld], a3);
        }
        for (; j < i; ++j) a0 = fma(-row[j], x[(int64_t)j * ld], a0);
        x[(int64_t)i * ld] = (a0 + a1) + (a2 + a3);
    }
    for (int i = n - 1; i >= 0; --i) {
        const double* __restrict__ row = LU + (int64_t)i * n;
        double a0 = x[(int64_t)i * ld], a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int j = i + 1;
        for (; j + 4 <= n; j += 4) {
            a0 = fma(-row[j], x[(int64_t)j * ld], a0);
            a1 = fma(-row[j + 1], x[(int64_t)(j + 1) * ld], a1);
            a2 = fma(-row[j + 2], x[(int64_t)(j + 2) * ld], a2);
            a3 = fma(-row[j + 3], x[(int64_t)(j + 3) * ld], a3);
        }
        for (; j < n; ++j) a0 = fma(-row[j], x[(int64_t)j * ld], a0);
        x[(int64_t)i * ld] = ((a0 + a1) + (a2 + a3)) / row[i];
    }
    if (SMEM && live)
        for (int i = 0; i < n; ++i) xg[(int64_t)i * ldg] = x[(int64_t)i * ld];
}

extern "C" int db_dense_solve(const db_dense_sys* sys, int32_t nsys, int32_t n, int32_t max_ncols, const double* lu, const int32_t* ipiv,
                              const db_veccomb* rhs, double* x, void* stream)
{
    if (nsys <= 0 || n <= 0 || max_ncols <= 0) return 0;
    if (rhs->nvec < 1 || rhs->nvec > 16) { db_set_error("dense_solve: 1..16 right-hand-side vectors"); return 1; }
    const dim3 grid((unsigned)nsys, (unsigned)((max_ncols + 31) / 32));
    const size_t smem = (size_t)n * 32 * sizeof(double);
    const char* env = getenv("DB_DENSE_SOLVE_SMEM");          // "0": keep the columns in global memory (the path taken for n > 800)
    if (smem <= 200 * 1024 && !(env && env[0] == '0')) {
        static bool attr_set = false;
        if (!attr_set) { DB_SET_SMEM_ATTR(k_dense_solve<true>); attr_set = true; }
        DB_LAUNCH(k_dense_solve<true>, grid, dim3(32), smem, stream, sys, n, lu, ipiv, *rhs, x);
    } else {
        DB_LAUNCH(k_dense_solve<false>, grid, dim3(32), 0, stream, sys, n, lu, ipiv, *rhs, x);
    }
    return db_check_launch("dense_solve");
}

// ---------------------------------------------------------------------------------------------------------
// ya = A x and / or yb = B x; thread = (row, column) of one system
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dense_matvec(const db_dense_sys* __restrict__ sys, int n, const double* __restrict__ a_all, const double* __restrict__ b_all,
               const double* __restrict__ x, double* __restrict__ ya, double* __restrict__ yb)
{
    const db_dense_sys S = sys[blockIdx.x];
    const int64_t total = (int64_t)n * S.ncols;
    const double* __restrict__ A = a_all + (int64_t)blockIdx.x * n * n;
    const double* __restrict__ B = b_all + (int64_t)blockIdx.x * n * n;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.y * blockDim.x) {
        const int i = (int)(e / S.ncols); const int r = (int)(e - (int64_t)i * S.ncols);
        double sa = 0.0, sb = 0.0;
        for (int j = 0; j < n; ++j) {
            const double xv = x[S.vec_off + (int64_t)j * S.ncols + r];
            if (ya) sa = fma(A[(int64_t)i * n + j], xv, sa);
            if (yb) sb = fma(B[(int64_t)i * n + j], xv, sb);
        }
        if (ya) ya[S.vec_off + e] = sa;
        if (yb) yb[S.vec_off + e] = sb;
    }
}

extern "C" int db_dense_matvec(const db_dense_sys* sys, int32_t nsys, int32_t n, const double* a, const double* b, const double* x,
                               double* ya, double* yb, void* stream)
{
    if (nsys <= 0 || n <= 0) return 0;
    DB_LAUNCH(k_dense_matvec, dim3((unsigned)nsys, 16), dim3(256), 0, stream, sys, n, a, b, x, ya, yb);
    return db_check_launch("dense_matvec");
}

// ---------------------------------------------------------------------------------------------------------
// Sparse form of the same products: the pencil OPERATORS M, L of a shell problem are 3-4 % dense (banded blocks + boundary rows
// + tau columns; only their LU factors fill in), so M.X and L.X use CSR: ptr [nsys][n + 1] with offsets into col / val that are
// global over the batch.  Thread = (row, column) of one system; either matrix may be absent (ptr == NULL).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_csr_matvec(const db_dense_sys* __restrict__ sys, int n, const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col,
             const double* __restrict__ a_val, const int64_t* __restrict__ b_ptr, const int32_t* __restrict__ b_col,
             const double* __restrict__ b_val, const double* __restrict__ x, double* __restrict__ ya, double* __restrict__ yb)
{
    const db_dense_sys S = sys[blockIdx.x];
    const int64_t total = (int64_t)n * S.ncols;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.y * blockDim.x) {
        const int i = (int)(e / S.ncols); const int r = (int)(e - (int64_t)i * S.ncols);
        const double* __restrict__ xr = x + S.vec_off + r;
        if (ya) {
            double s = 0.0;
            const int64_t p0 = a_ptr[(int64_t)blockIdx.x * (n + 1) + i], p1 = a_ptr[(int64_t)blockIdx.x * (n + 1) + i + 1];
            for (int64_t p = p0; p < p1; ++p) s = fma(a_val[p], xr[(int64_t)a_col[p] * S.ncols], s);
            ya[S.vec_off + e] = s;
        }
        if (yb) {
            double s = 0.0;
            const int64_t p0 = b_ptr[(int64_t)blockIdx.x * (n + 1) + i], p1 = b_ptr[(int64_t)blockIdx.x * (n + 1) + i + 1];
            for (int64_t p = p0; p < p1; ++p) s = fma(b_val[p], xr[(int64_t)b_col[p] * S.ncols], s);
            yb[S.vec_off + e] = s;
        }
    }
}

extern "C" int db_csr_matvec(const db_dense_sys* sys, int32_t nsys, int32_t n, const int64_t* a_ptr, const int32_t* a_col, const double* a_val,
                             const int64_t* b_ptr, const int32_t* b_col, const double* b_val, const double* x, double* ya, double* yb,
                             void* stream)
{
    if (nsys <= 0 || n <= 0) return 0;
    DB_LAUNCH(k_csr_matvec, dim3((unsigned)nsys, 16), dim3(256), 0, stream, sys, n, a_ptr, a_col, a_val, b_ptr, b_col, b_val, x, ya, yb);
    return db_check_launch("csr_matvec");
}
