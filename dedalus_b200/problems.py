"""Initial value problems  M.dt(X) + L.X = F(X, t)   (reference core/problems.py:40-362, IVP 270-362) and the linear
boundary value problems  L.X = F  the stock scripts use to balance their initial conditions (LBVP, problems.py:155-203).

`add_equation("LHS = RHS")` evaluates both sides in the user namespace (exactly like the reference,
problems.py:66-100), splits the LHS into the dt-terms (M) and the rest (L) and keeps the RHS as the explicit
term F.  Linear-map extraction and the per-wavenumber matrix templates are built lazily by the solver.
"""
import numbers
import numpy as np
from . import operators as ops
from .operators import Operand
from .field import Field


class IVP:
    def __init__(self, variables, time='t', namespace=None):
        self.variables = list(variables)
        self.LHS_variables = self.variables
        self.dist = self.variables[0].dist
        self.dtype = self.variables[0].dtype
        self.equations = []
        self.time = Field(self.dist, name=time) if isinstance(time, str) else time
        self.namespace = {}
        # default namespace: public operators, then variables, then user entries (user wins, as in the reference)
        import dedalus_b200 as pkg
        for k in dir(pkg):
            if not k.startswith('_'):
                self.namespace[k] = getattr(pkg, k)
        self.namespace['dt'] = ops.dt
        # numpy ufuncs by name, as the reference's parsing namespace has them (core/operators.py:558-560: 'sin', 'tanh', 'abs', ...)
        for name in ('absolute', 'sign', 'exp', 'exp2', 'log', 'log2', 'log10', 'sqrt', 'square', 'sin', 'cos', 'tan', 'arcsin',
                     'arccos', 'arctan', 'sinh', 'cosh', 'tanh', 'arcsinh', 'arccosh', 'arctanh'):
            self.namespace[name] = getattr(np, name)
        self.namespace['abs'] = np.absolute
        self.namespace[self.time.name] = self.time
        for v in self.variables:
            if v.name:
                self.namespace[v.name] = v
        if namespace:
            self.namespace.update({k: v for k, v in namespace.items() if not k.startswith('__')})
        self.namespace['dt'] = self.namespace.get('dt') if callable(self.namespace.get('dt')) else ops.dt

    @staticmethod
    def _split_equation(eq):
        depth = 0
        idx = []
        for i, ch in enumerate(eq):
            if ch in '([{':
                depth += 1
            elif ch in ')]}':
                depth -= 1
            elif ch == '=' and depth == 0:
                idx.append(i)
        if len(idx) != 1:
            raise ValueError("Equation string must contain exactly one top-level '='.")
        return eq[:idx[0]].strip(), eq[idx[0] + 1:].strip()

    def add_equation(self, equation, condition=None):
        if isinstance(equation, str):
            lhs_s, rhs_s = self._split_equation(equation)
            LHS = eval(lhs_s, dict(self.namespace))
            RHS = eval(rhs_s, dict(self.namespace))
        else:
            LHS, RHS = equation
        if not isinstance(LHS, Operand):
            raise ValueError("LHS must be an operand expression of the problem variables.")
        eq = {'LHS': LHS, 'RHS': RHS, 'condition': condition,
              'tensorsig': LHS.tensorsig, 'bases': LHS.bases, 'dtype': self.dtype}
        if isinstance(RHS, Operand):
            if RHS.tensorsig != LHS.tensorsig:
                raise ValueError("LHS and RHS tensor signatures differ.")
        elif isinstance(RHS, numbers.Number):
            if RHS != 0 and LHS.tensorsig:
                raise ValueError("Nonzero numeric RHS requires a scalar equation.")
        else:
            raise ValueError("RHS must be an operand or a number.")
        self.equations.append(eq)
        return eq

    def build_solver(self, timestepper, **kw):
        from .solvers import InitialValueSolver
        return InitialValueSolver(self, timestepper, **kw)


class LBVP(IVP):
    """L.X = F with F independent of X (reference core/problems.py:155-203).  Same namespace and equation parsing as the IVP;
    time derivatives are not allowed."""

    def __init__(self, variables, namespace=None):
        super().__init__(variables, time='t', namespace=namespace)

    def add_equation(self, equation, condition=None):
        eq = super().add_equation(equation, condition)
        from .sphere import _has
        if _has(eq['LHS'], (ops.TimeDerivative,)):
            raise ValueError("LBVP LHS must not contain time derivatives.")
        return eq

    def build_solver(self, **kw):
        from .solvers import LinearBoundaryValueSolver
        return LinearBoundaryValueSolver(self, **kw)
