"""CFL timestep controller and global flow properties evaluated on the device
(reference extras/flow_tools.py:64-233; operators.AdvectiveCFL core/operators.py:4342-4400).

The advective frequency  sum_i |u_i| / dx_i  is reduced to its grid maximum by one kernel (csrc/core.cu
k_cfl_max) on the dealiased grid values of the velocity, every `cadence` iterations at the start of the step
(the reference evaluates its dictionary handler inside stage 1, core/timesteppers.py:607-608); only the scalar
maximum crosses to the host (and, on several GPUs, an all-reduce of one double)."""
import ctypes as C
import numpy as np


class CFL:
    def __init__(self, solver, initial_dt, cadence=1, safety=1., max_dt=np.inf, min_dt=0., max_change=np.inf,
                 min_change=0., threshold=0.):
        self.solver = solver
        self.stored_dt = initial_dt
        self.cadence = cadence
        self.safety = safety
        self.max_dt, self.min_dt = max_dt, min_dt
        self.max_change, self.min_change = max_change, min_change
        self.threshold = threshold
        self.velocities = []
        self.max_freq = None
        solver.step_hooks.append(self._on_step)

    def add_velocity(self, velocity):
        if len(velocity.tensorsig) != 1:
            raise ValueError("Velocity must be a vector")
        self.velocities.append(velocity)

    # called by the solver at the start of every step, state in coefficient space on the device
    def _on_step(self, solver):
        if solver.iteration % self.cadence != 0 or not self.velocities:
            return
        import torch
        from ..lib import get_lib, current_stream
        from ..transforms import cached_plan
        lib = get_lib()
        out = torch.zeros(1, dtype=torch.float64, device=solver.device)
        for u in self.velocities:
            dist = u.dist
            scales = u.dealias
            # grid values of every component at dealias scales (own buffers: the state stays in coefficient space)
            g = self._to_grid(u)
            if getattr(u.tensorsig[0], 'curvilinear', False):
                self._spherical_frequency(u, g, out)
                continue
            dim = dist.dim
            comps = [g[i] for i in range(g.shape[0])]
            inv = []
            for i, coord in enumerate(u.tensorsig[0].coords):
                ax = dist.get_axis(coord)
                b = u.bases[ax]
                dx = self._cfl_spacing(b, scales[ax])
                sl = dist.grid_local_slice(ax, b, scales[ax])
                inv.append(torch.from_numpy(np.ascontiguousarray(1.0 / np.abs(dx[sl]))).to(solver.device))
            shape = [1, 1, 1]
            gs = list(comps[0].shape)
            shape[3 - len(gs):] = gs
            up = (C.c_void_p * len(comps))(*[c.data_ptr() for c in comps])
            ip = (C.c_void_p * len(inv))(*[t.data_ptr() for t in inv])
            lib.call("db_cfl_max", up, ip, len(comps), shape[0], shape[1], shape[2], out.data_ptr(), current_stream())
        if solver.dist.size > 1:
            import torch.distributed as td
            td.all_reduce(out, op=td.ReduceOp.MAX)
        self.max_freq = out       # read lazily in compute_timestep (one host sync every `cadence` steps)

    def _spherical_frequency(self, u, g, out):
        """Velocities on a sphere or in a spherical shell: sqrt(u_phi^2 + u_theta^2) sqrt(Lmax (Lmax + 1)) / r + |u_r| / dr_eff
        (reference S2AdvectiveCFL / Spherical3DAdvectiveCFL, core/basis.py:6156-6212), one reduction kernel."""
        import torch
        from ..lib import get_lib, current_stream
        from ..sphere import sphere_basis_of
        from ..shell import shell_basis_of
        dev = g.device
        shell = shell_basis_of(u)
        if shell is not None:
            scale = shell.dealias[2]
            r = shell.global_grid_radius(scale)
            L = shell.Lmax
            inv_h = (np.sqrt(L * (L + 1)) / r) if L > 0 else np.zeros_like(r)
            inv_dr = 1.0 / np.abs(np.gradient(r, edge_order=2) * scale)
            n_r = r.size
            ur = g[2].contiguous()
        else:
            sb = sphere_basis_of(u)
            L = sb.Lmax
            inv_h = np.array([np.sqrt(L * (L + 1)) / sb.radius if L > 0 else 0.0])
            inv_dr, n_r, ur = np.zeros(1), 1, None
        up, ut = g[0].contiguous(), g[1].contiguous()
        ih = torch.from_numpy(np.ascontiguousarray(inv_h)).to(dev)
        idr = torch.from_numpy(np.ascontiguousarray(inv_dr)).to(dev)
        get_lib().call("db_cfl_max_spherical", up.data_ptr(), ut.data_ptr(), ur.data_ptr() if ur is not None else None, ih.data_ptr(),
                       idr.data_ptr(), up.numel() // n_r, n_r, out.data_ptr(), current_stream())

    @staticmethod
    def _cfl_spacing(basis, dealias):
        """Effective grid spacing used by the reference's CartesianAdvectiveCFL.cfl_spacing (core/basis.py:6083-6104)."""
        from ..basis import Jacobi, RealFourier, ComplexFourier
        N = basis.grid_size(dealias)
        if isinstance(basis, Jacobi) and basis.a == -0.5 and basis.b == -0.5:
            theta = np.pi * (np.arange(N) + 0.5) / N
            return dealias * basis.COV.stretch * np.sin(theta) * np.pi / N
        if isinstance(basis, (RealFourier, ComplexFourier)):
            return np.full(N, dealias * (2 * np.pi / N) * basis.COV.stretch)
        grid = basis.global_grid(dealias)
        return np.gradient(grid, edge_order=2) * dealias

    def _to_grid(self, u):
        import torch
        tmp = u.copy_device_to_grid()
        return tmp

    def compute_timestep(self):
        """Same update rule as the reference (flow_tools.py:191-214)."""
        iteration = self.solver.iteration
        if (iteration - 1) % self.cadence == 0:
            if (iteration - 1) <= self.solver.initial_iteration or self.max_freq is None:
                return self.stored_dt
            max_global_freq = float(self.max_freq.item())
            dt = np.inf if max_global_freq == 0. else 1 / max_global_freq
            dt *= self.safety
            dt = min(dt, self.max_dt, self.max_change * self.stored_dt)
            dt = max(dt, self.min_dt, self.min_change * self.stored_dt)
            if abs(dt - self.stored_dt) > self.threshold * self.stored_dt:
                self.stored_dt = dt
        return self.stored_dt

    compute_dt = compute_timestep


class GlobalFlowProperty:
    """Scalar reductions of grid-space quantities (reference flow_tools.py:64-150): max / min / mean of a field or of an operator
    expression (`np.sqrt(u@u)/nu` in the stock scripts).  Expressions are evaluated on the device every `cadence` iterations at
    the start of the step, like the reference's dictionary handler (iter=cadence), and reduced at scale 1; the reductions
    return the value of the last scheduled evaluation.  A field property is read when the reduction is called, as in the
    reference, where the handler's output IS the field."""

    def __init__(self, solver, cadence=1):
        self.solver = solver
        self.cadence = cadence
        self.properties = {}
        self._values = {}
        self._fields = {}
        solver.step_hooks.append(self._on_step)

    def add_property(self, property, name, precompute_integral=False):
        from ..operators import Operand
        if not isinstance(property, Operand):
            raise ValueError("flow properties must be fields or operator expressions")
        self.properties[name] = property

    def _on_step(self, solver):
        if solver.iteration % self.cadence != 0:
            return
        for name, p in self.properties.items():
            if not hasattr(p, 'copy_device_to_grid'):
                self._fields[name] = p.evaluate()
                self._values[name] = self._fields[name].copy_device_to_grid(scales=1)

    def _grid(self, name):
        p = self.properties[name]
        if hasattr(p, 'copy_device_to_grid'):
            return p.copy_device_to_grid()
        if name not in self._values:              # asked before the first scheduled evaluation
            self._fields[name] = p.evaluate()
            self._values[name] = self._fields[name].copy_device_to_grid(scales=1)
        return self._values[name]

    def _reduce(self, v, op):
        if self.solver.dist.size > 1:
            import torch.distributed as td
            td.all_reduce(v, op=getattr(td.ReduceOp, op))
        return float(v.item())

    def max(self, name):
        return self._reduce(self._grid(name).max(), 'MAX')

    def min(self, name):
        return self._reduce(self._grid(name).min(), 'MIN')

    def grid_average(self, name):
        """Mean over all grid points (reference flow_tools.py:112-115)."""
        import torch
        g = self._grid(name)
        total = self._reduce(g.sum(), 'SUM')
        count = self._reduce(torch.tensor(float(g.numel()), dtype=torch.float64, device=g.device), 'SUM')
        return total / count

    def volume_integral(self, name):
        """Volume integral of a property (reference flow_tools.py:117-130: Integrate of the handler's output field), Cartesian
        domains."""
        from ..operators import Integrate
        p = self.properties[name]
        if not hasattr(p, 'copy_device_to_grid'):
            self._grid(name)
            p = self._fields[name]
        if getattr(p, 'copy_device_to_grid', None) is None or getattr(p.dist.coordsys, 'curvilinear', False):
            raise NotImplementedError("volume integrals of flow properties on curvilinear domains")
        return float(np.asarray(Integrate(p).evaluate()['g']).ravel()[0])

    def volume_average(self, name):
        raise NotImplementedError("missing definition of hypervolume")      # as in the reference (flow_tools.py:132-137)


class AdvectiveCFL:
    """The advective CFL frequency as a field on the dealiased grid (reference operators.AdvectiveCFL, core/operators.py:4342-4400;
    spacings core/basis.py:6078-6212).  `evaluate()` returns the frequency field, `cfl_spacing()` the per-direction spacings it
    divides by.  The CFL controller above reduces the same expression to its maximum in one kernel without forming this field."""

    def __init__(self, operand, coords=None):
        if len(operand.tensorsig) != 1:
            raise ValueError("Velocity must be a vector")
        self.operand = operand
        self.coords = coords if coords is not None else operand.tensorsig[0]
        self.dist = operand.dist

    def _shaped(self, values, axis):
        shape = [1] * self.dist.dim
        shape[axis] = len(values)
        return np.asarray(values, dtype=float).reshape(shape)

    def cfl_spacing(self):
        from ..sphere import sphere_basis_of
        from ..shell import shell_basis_of
        u = self.operand
        dist = self.dist
        cs = u.tensorsig[0]
        if getattr(cs, 'curvilinear', False):
            shell = shell_basis_of(u)
            basis = shell if shell is not None else sphere_basis_of(u)
            L = basis.Lmax
            ax = dist.get_basis_axis(basis)
            if shell is not None:
                scale = shell.dealias[2]
                r = shell.global_grid_radius(scale)
                s2 = self._shaped(r / np.sqrt(L * (L + 1)) if L > 0 else np.full(r.size, np.inf), ax + 2)
                return [s2, self._shaped(np.abs(np.gradient(r, edge_order=2) * scale), ax + 2)]
            return [np.array(basis.radius / np.sqrt(L * (L + 1)) if L > 0 else np.inf).reshape([1] * dist.dim)]
        out = []
        for coord in cs.coords:
            ax = dist.get_axis(coord)
            b = u.bases[ax]
            if b is None:
                out.append(np.inf)
                continue
            scale = b.dealias[0]
            dx = CFL._cfl_spacing(b, scale)
            out.append(self._shaped(np.abs(dx[dist.grid_local_slice(ax, b, scale)]), ax))
        return out

    def evaluate(self):
        import torch
        from ..field import Field
        u = self.operand
        g = u.copy_device_to_grid()                       # (ncomp, grid) at dealias scales
        sp = [torch.from_numpy(np.array(np.broadcast_to(s, g.shape[1:]), dtype=float)).to(g.device) for s in self.cfl_spacing()]
        u.change_scales(u.dealias)      # as in the reference, the operand is left on the dealiased grid scales
        if getattr(u.tensorsig[0], 'curvilinear', False):
            freq = torch.sqrt(g[0] ** 2 + g[1] ** 2) / sp[0]
            if len(sp) > 1:
                freq = freq + g[2].abs() / sp[1]
        else:
            freq = sum(g[i].abs() / sp[i] for i in range(len(sp)))
        out = Field(self.dist, bases=u.unique_bases(), dtype=u.dtype)
        self.dist._fields.pop()
        out.set_device_data(freq.contiguous(), 'g', scales=u.dealias)
        return out
