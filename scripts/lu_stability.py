"""Host-side diagnostic: element growth and solve error of the STATIC pivot order of every batch (dense numpy LU without
pivot search in the batch's ordering, exactly what k_batches_factor does) against a partially pivoted dense solve.
usage: python scripts/lu_stability.py [Nh] [Nz] [dt] [threshold]"""
import sys, time
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, ".")
from dedalus_b200 import examples
from dedalus_b200.pencils import PencilSystemBuilder, build_batches

GAMMA = (2 - np.sqrt(2)) / 2


def static_lu_growth(A):
    """LU without pivoting; returns (growth = max|LU|/max|A|, LU)."""
    LU = A.copy()
    n = LU.shape[0]
    amax = np.abs(A).max()
    g = amax
    for k in range(n - 1):
        piv = LU[k, k]
        rows = k + 1 + np.nonzero(LU[k + 1:, k])[0]
        if rows.size:
            l = LU[rows, k] / piv
            LU[rows, k] = l
            cols = k + 1 + np.nonzero(LU[k, k + 1:])[0]
            if cols.size:
                LU[np.ix_(rows, cols)] -= np.outer(l, LU[k, cols])
            g = max(g, np.abs(l).max())
    g = max(g, np.abs(LU).max())
    return g / amax, LU


def solve_lu(LU, b):
    n = LU.shape[0]
    L = np.tril(LU, -1) + np.eye(n)
    U = np.triu(LU)
    y = sla.solve_triangular(L, b, lower=True, unit_diagonal=True)
    return sla.solve_triangular(U, y)


def main():
    Nh = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    Nz = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0025
    thr = float(sys.argv[4]) if len(sys.argv) > 4 else None
    pb = examples.rayleigh_benard(dim=3, Nh=Nh, Nz=Nz)
    builder = PencilSystemBuilder(pb['problem'])
    batches = build_batches(builder)
    a0, b0 = 1.0, dt * GAMMA
    rng = np.random.default_rng(0)
    for ib, batch in enumerate(batches):
        t0 = time.time()
        if thr is None:
            batch.compute_ordering(a0, b0)
        else:
            batch.compute_ordering(a0, b0, threshold=thr)
        t1 = time.time()
        F = batch.symbolic_lu()
        worst_g, worst_e = 0, 0
        for g in batch.representative_groups(4):
            A = batch.matrix((a0, b0), g).toarray()
            growth, LU = static_lu_growth(A)
            b = rng.standard_normal(A.shape[0])
            x = solve_lu(LU, b)
            xr = np.linalg.solve(A, b)
            err = np.abs(x - xr).max() / np.abs(xr).max()
            worst_g, worst_e = max(worst_g, growth), max(worst_e, err)
        print(f"batch {ib}: zero_axes={batch.cls.zero_axes} n={batch.n} S={batch.S} nnz(LU)/n={F.sum() / batch.n:.1f} "
              f"growth={worst_g:.2e} err={worst_e:.2e} order {t1 - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
