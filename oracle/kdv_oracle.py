"""CPU restatement of the KdV-Burgers example (examples/ivp_1d_kdv_burgers/kdv_burgers.py:22-54):
dt(u) - a*dx(dx(u)) - b*dx(dx(dx(u))) = -u*dx(u), RealFourier, SBDF2 or Runge-Kutta.  TEST INFRASTRUCTURE.
The fully separable problem is solved per 2x2 wavenumber block (the reference couples the last axis into a
single block-diagonal subproblem, core/solvers.py:70-74; the solutions are identical)."""
import numpy as np
from . import transforms_oracle as T
from . import imex

J2 = np.array([[0., -1.], [1., 0.]])


class KdVOracle:
    def __init__(self, N, Lx=10.0, a=1e-4, b=2e-4, dealias=1.5):
        self.N, self.G = N, int(dealias * N)
        self.k = np.arange(N // 2) * 2 * np.pi / Lx
        self.a, self.b = a, b

    def L_blocks(self):
        out = []
        for k in self.k:
            D = k * J2
            out.append(-self.a * D @ D - self.b * D @ D @ D)
        return out

    def rhs(self, u):
        du = np.zeros_like(u)
        du[0::2] = -self.k * u[1::2]; du[1::2] = self.k * u[0::2]
        ug = T.rf_backward_fft(u, self.G, 0); dug = T.rf_backward_fft(du, self.G, 0)
        return T.rf_forward_fft(-ug * dug, self.N, 0)

    def apply(self, blocks, u):
        out = np.zeros_like(u)
        for i, B in enumerate(blocks):
            out[2 * i:2 * i + 2] = B @ u[2 * i:2 * i + 2]
        out[1] = 0
        return out

    def solve(self, a0, b0, rhs):
        out = np.zeros_like(rhs)
        for i, B in enumerate(self.L_blocks()):
            A = a0 * np.eye(2) + b0 * B
            if i == 0:
                out[0] = rhs[0] / A[0, 0]
            else:
                out[2 * i:2 * i + 2] = np.linalg.solve(A, rhs[2 * i:2 * i + 2])
        return out

    def run(self, u0, steps, dt, scheme="SBDF2"):
        u = np.array(u0, dtype=float, copy=True)
        Lb = self.L_blocks()
        if scheme == "SBDF2":
            hist = dict(MX=[], LX=[], F=[], dts=[])
            for it in range(steps):
                hist['dts'] = [dt] + hist['dts'][:1]
                a, b, c = imex.sbdf2(hist['dts'][0], hist['dts'][1] if len(hist['dts']) > 1 else dt, it)
                hist['MX'] = [u.copy()] + hist['MX'][:1]
                hist['LX'] = [self.apply(Lb, u)] + hist['LX'][:1]
                hist['F'] = [self.rhs(u)] + hist['F'][:1]
                rhs = np.zeros_like(u)
                for j in range(1, 3):
                    if c[j] != 0: rhs += c[j] * hist['F'][j - 1]
                    if a[j] != 0: rhs -= a[j] * hist['MX'][j - 1]
                    if b[j] != 0: rhs -= b[j] * hist['LX'][j - 1]
                u = self.solve(a[0], b[0], rhs)
            return u
        tab = imex.RK[scheme]; A, H = tab['A'], tab['H']; s = len(tab['c']) - 1
        for it in range(steps):
            MX0 = u.copy(); LX = [self.apply(Lb, u)]; F = []
            for i in range(1, s + 1):
                if i > 1:
                    LX.append(self.apply(Lb, u))
                F.append(self.rhs(u))
                rhs = MX0.copy()
                for j in range(i):
                    rhs += dt * A[i, j] * F[j] - dt * H[i, j] * LX[j]
                u = self.solve(1.0, dt * H[i, i], rhs)
        return u
