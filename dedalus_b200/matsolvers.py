"""Matrix solvers behind the reference's Matsolver plugin contract (libraries/matsolvers.py:10-13, 126-194):

    matsolvers[name.lower()]             registry filled by @add_solver
    solver = cls(matrix, solver=None)    factorises a scipy.sparse matrix
    x = solver.solve(vector)             vector of shape (n,) or (n, k); RETURNS A NEW ARRAY (core/timesteppers.py:183, 642)

so that `problem.build_solver(..., matsolver='B200Dense')` of the reference -- or its `[linear algebra] MATRIX_FACTORIZER` setting --
can route each pencil's factorisation and solves to the device kernels of this library one matrix at a time.  This is the drop-in
form of row S3; the fast path of this repo factorises and solves ALL pencils of a problem in one launch each (`pencils.py`,
`sphere.SphereSystems`, `shell_ivp.ShellSystems`) through the same C-ABI entry points.

  B200Dense   LU with partial pivoting of the dense matrix (csrc/dense.cu: db_dense_factor / db_dense_solve, one column per thread)
  B200Banded  band LU with partial pivoting (csrc/banded.cu: db_banded_factor / db_banded_solve), for matrices whose bandwidth is
              small against n (the reference's ScipyBanded, matsolvers.py:186-194)

Inputs and outputs are numpy arrays (the reference's callers hand over host arrays); the factors live on the device."""
import ctypes as C
import numpy as np

matsolvers = {}


def add_solver(solver):
    matsolvers[solver.__name__.lower()] = solver
    return solver


class _DeviceSolver:
    def __init__(self, matrix, solver=None):
        import torch
        from scipy import sparse
        from .lib import get_lib, compute_device, current_stream, DedalusB200Error
        if np.iscomplexobj(matrix.data if sparse.issparse(matrix) else matrix):
            raise NotImplementedError("complex matrices: embed as real 2 x 2 blocks (dedalus_b200/pencils.py)")
        self.lib, self.device, self.stream, self.Error = get_lib(), compute_device(), current_stream, DedalusB200Error
        self.n = int(matrix.shape[0])
        if matrix.shape[0] != matrix.shape[1]:
            raise ValueError("matrix must be square")
        self.torch = torch
        self._factor(sparse.csr_matrix(matrix))

    def _to_dev(self, a, dtype=None):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def _check(self, info):
        if int(info.sum().item()) != 0:
            raise self.Error("matrix is singular to working precision (zero or non-finite pivot)")

    def solve(self, vector):
        vector = np.asarray(vector, dtype=np.float64)
        b = vector.reshape(self.n, -1)
        x = self._solve(self._to_dev(b))
        return x.cpu().numpy().reshape(vector.shape)


@add_solver
class B200Dense(_DeviceSolver):
    """Dense LU with partial pivoting on the device."""

    def _factor(self, A):
        torch = self.torch
        n = self.n
        self.lu = self._to_dev(A.toarray())
        self.ipiv = torch.zeros(n, dtype=torch.int32, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.lib.call("db_dense_factor", 1, n, self.lu.data_ptr(), self.ipiv.data_ptr(), info.data_ptr(), self.stream())
        self._check(info)

    def _solve(self, b):
        from .lib import DenseSys, VecComb
        torch = self.torch
        k = b.shape[1]
        sysd = DenseSys(); sysd.ncols, sysd.pad, sysd.vec_off = k, 0, 0
        desc = torch.from_numpy(np.frombuffer(bytes(sysd), dtype=np.uint8).copy()).to(self.device)
        x = torch.empty_like(b)
        vc = VecComb(); vc.nvec = 1; vc.vec[0] = b.data_ptr(); vc.coef[0] = 1.0
        self.lib.call("db_dense_solve", desc.data_ptr(), 1, self.n, k, self.lu.data_ptr(), self.ipiv.data_ptr(), C.byref(vc), x.data_ptr(),
                      self.stream())
        return x


@add_solver
class B200Banded(_DeviceSolver):
    """Band LU with partial pivoting on the device (LAPACK band storage with kl extra rows for the fill-in)."""

    def _factor(self, A):
        from .lib import BandedSys
        torch = self.torch
        n = self.n
        coo = A.tocoo()
        kl = int(max(0, (coo.row - coo.col).max())) if coo.nnz else 0
        ku = int(max(0, (coo.col - coo.row).max())) if coo.nnz else 0
        self.kl, self.ku = kl, ku = max(kl, 1), max(ku, 1)
        ld0, ldf = kl + ku + 1, 2 * kl + ku + 1
        ab = np.zeros(n * ld0)
        ab[coo.col.astype(np.int64) * ld0 + (ku + coo.row - coo.col)] = coo.data          # column-major bands: (ku + i - j, j)
        sysd = BandedSys(); sysd.n, sysd.nrhs, sysd.op_off, sysd.lu_off, sysd.piv_off, sysd.vec_off = n, 1, 0, 0, 0, 0
        self._sys = sysd
        self.desc = torch.from_numpy(np.frombuffer(bytes(sysd), dtype=np.uint8).copy()).to(self.device)
        op = self._to_dev(ab)
        self.lu = torch.zeros(n * ldf, dtype=torch.float64, device=self.device)
        self.ipiv = torch.zeros(n, dtype=torch.int32, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        st = self.stream()
        self.lib.call("db_banded_combine", self.desc.data_ptr(), 1, kl, ku, 0.0, op.data_ptr(), 1.0, op.data_ptr(), self.lu.data_ptr(), st)
        self.lib.call("db_banded_factor", self.desc.data_ptr(), 1, kl, ku, self.lu.data_ptr(), self.ipiv.data_ptr(), info.data_ptr(), st)
        self._check(info)

    def _solve(self, b):
        from .lib import BandedSys, VecComb
        torch = self.torch
        k = b.shape[1]
        sysd = BandedSys(); sysd.n, sysd.nrhs, sysd.op_off, sysd.lu_off, sysd.piv_off, sysd.vec_off = self.n, k, 0, 0, 0, 0
        desc = torch.from_numpy(np.frombuffer(bytes(sysd), dtype=np.uint8).copy()).to(self.device)
        x = torch.empty_like(b)
        vc = VecComb(); vc.nvec = 1; vc.vec[0] = b.data_ptr(); vc.coef[0] = 1.0
        self.lib.call("db_banded_solve", desc.data_ptr(), 1, self.kl, self.ku, self.n, k, self.lu.data_ptr(), self.ipiv.data_ptr(),
                      C.byref(vc), x.data_ptr(), self.stream())
        return x
