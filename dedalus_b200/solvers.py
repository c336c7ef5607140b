"""InitialValueSolver with a device-resident IMEX stage loop.

Reference: core/solvers.py:503-806 (InitialValueSolver; `step` 683-711) driving core/timesteppers.py:95-187
(MultistepIMEX.step) and 552-644 (RungeKuttaIMEX.step), whose bodies are Python loops over pencils.
Here one step is a fixed sequence of kernel launches per *batch* of structurally identical pencil systems
(dedalus_b200/pencils.py): gather -> template mat-vecs (M.X, L.X) -> RHS evaluation (evaluator.RHSPlan)
-> fused RHS-combination + triangular solves -> scatter.  Factorisations are rebuilt only when the LHS
coefficients (a0, b0) change, exactly as the reference drops its LHS_solvers (timesteppers.py:135-140, 577-583).
"""
import time
import ctypes as C
from collections import deque
import numpy as np
from . import timesteppers as ts
from .pencils import (PencilSystemBuilder, build_batches, compile_batch, assembly_program, Arena, line_maps)
from .lib import get_lib, LinComb, DedalusB200Error


def _i32(t, a, dev):
    return t.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


class Timed:
    """Optional CUDA-event bracket around one kernel launch (bench.py roofline accounting)."""

    def __init__(self, prof, name, nbytes):
        self.prof, self.name, self.nbytes = prof, name, nbytes

    def __enter__(self):
        if self.prof is not None:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.prof is not None:
            self.e1.record()
            self.prof.append((self.name, self.e0, self.e1, self.nbytes))
        return False


class DeviceBatch:
    """Device copies of one batch's programs and work vectors."""

    def __init__(self, solver, batch):
        import torch
        self.batch = batch
        self.solver = solver
        self.prog = None
        self.lu = {}          # b0/a0 ratio key -> LU tensor

    def upload(self, a0, b0):
        import torch
        dev = self.solver.device
        batch = self.batch
        prog = compile_batch(batch, a0, b0)
        self.prog = prog
        n, S = prog.n, prog.S
        self.n, self.S = n, S
        self.ld = ld = ((S + 31) // 32) * 32
        mono = np.zeros((len(prog.monos), ld)); mono[:, :S] = prog.mono_vals
        self.mono = torch.from_numpy(mono).to(dev)
        f = lambda a: _i32(torch, a, dev)
        self.diag_eid, self.fl_ptr, self.fl_eid = f(prog.diag_eid), f(prog.fl_ptr), f(prog.fl_eid)
        self.fu_ptr, self.fu_eid, self.fd_eid = f(prog.fu_ptr), f(prog.fu_eid), f(prog.fd_eid)
        self.fwd_ptr, self.fwd_col, self.bwd_ptr, self.bwd_col = f(prog.fwd_ptr), f(prog.fwd_col), f(prog.bwd_ptr), f(prog.bwd_col)
        self.mv = {}
        for name in ('M', 'L'):
            ptr, col, mono_i, val = prog.mv[name]
            self.mv[name] = (f(ptr), f(col), f(mono_i), torch.from_numpy(np.ascontiguousarray(val)).to(dev))
        self.maps = {}
        for side, arena in (('cols', self.solver.var_arena), ('rows', self.solver.eq_arena)):
            m = line_maps(batch, arena, side)
            so = np.zeros((m.sys_off.shape[0], ld), dtype=np.int64); so[:, :S] = m.sys_off
            self.maps[side] = dict(nlines=len(m.line_base), max_len=int(m.line_len.max()) if len(m.line_len) else 0,
                                   base=torch.from_numpy(m.line_base).to(dev), kind=f(m.line_kind), ptr=f(m.line_ptr),
                                   pos=f(m.line_pos), sys_off=torch.from_numpy(so).to(dev))
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)

    def vec(self):
        import torch
        return torch.zeros((self.n, self.ld), dtype=torch.float64, device=self.solver.device)

    # ---- kernels -----------------------------------------------------------------------------------
    def gather(self, side, arena_t, vec):
        m = self.maps[side]
        with Timed(self.solver.prof, "pencil_gather", 16 * self.n * self.S):
          self.solver.lib.call("db_pencil_gather", arena_t.data_ptr(), vec.data_ptr(), self.S, self.ld, m['nlines'], m['max_len'],
                             m['base'].data_ptr(), m['kind'].data_ptr(), m['ptr'].data_ptr(), m['pos'].data_ptr(),
                             m['sys_off'].data_ptr(), self.ld, self.solver.stream())

    def scatter(self, vec, arena_t):
        m = self.maps['cols']
        with Timed(self.solver.prof, "pencil_scatter", 16 * self.n * self.S):
          self.solver.lib.call("db_pencil_scatter", vec.data_ptr(), arena_t.data_ptr(), self.S, self.ld, m['nlines'], m['max_len'],
                             m['base'].data_ptr(), m['kind'].data_ptr(), m['ptr'].data_ptr(), m['pos'].data_ptr(),
                             m['sys_off'].data_ptr(), self.ld, self.solver.stream())

    def matvec(self, x, y_m=None, y_l=None):
        M, L = self.mv['M'], self.mv['L']
        p = lambda t: t.data_ptr() if t is not None else None
        nout = (y_m is not None) + (y_l is not None)
        with Timed(self.solver.prof, "pencil_matvec", 8 * self.n * self.S * (1 + nout)):
          self.solver.lib.call("db_pencil_matvec", self.n, self.S, self.ld, self.mono.data_ptr(), x.data_ptr(),
                             M[0].data_ptr(), M[1].data_ptr(), M[2].data_ptr(), M[3].data_ptr(), p(y_m),
                             L[0].data_ptr(), L[1].data_ptr(), L[2].data_ptr(), L[3].data_ptr(), p(y_l), self.solver.stream())

    def factor(self, key, a0, b0):
        import torch
        dev = self.solver.device
        ptr, mono_i, val = assembly_program(self.batch, self.prog, a0, b0)
        lu = self.lu.get(key)
        if lu is None:
            lu = torch.empty((self.prog.nE, self.ld), dtype=torch.float64, device=dev)
            self.lu[key] = lu
        t_ptr, t_mono, t_val = _i32(torch, ptr, dev), _i32(torch, mono_i, dev), torch.from_numpy(val).to(dev)
        lib = self.solver.lib
        lib.call("db_pencil_assemble", lu.data_ptr(), self.prog.nE, self.S, self.ld, self.mono.data_ptr(),
                 t_ptr.data_ptr(), t_mono.data_ptr(), t_val.data_ptr(), self.solver.stream())
        self.info.zero_()
        lib.call("db_pencil_factor", lu.data_ptr(), self.n, self.S, self.ld, self.diag_eid.data_ptr(), self.fl_ptr.data_ptr(),
                 self.fl_eid.data_ptr(), self.fu_ptr.data_ptr(), self.fu_eid.data_ptr(), self.fd_eid.data_ptr(),
                 self.info.data_ptr(), self.solver.stream())
        return lu

    def solve(self, lu, terms, x):
        lc = LinComb()
        lc.nvec = len(terms)
        for j, (v, c) in enumerate(terms):
            lc.vec[j] = v.data_ptr(); lc.coef[j] = c
        # algorithmic bytes: every stored LU entry once + each RHS vector once + the solution written once
        nbytes = 8 * self.S * (self.prog.nE + self.n * (len(terms) + 1))
        with Timed(self.solver.prof, "pencil_solve", nbytes):
          self.solver.lib.call("db_pencil_solve", lu.data_ptr(), self.n, self.S, self.ld, self.fwd_ptr.data_ptr(), self.fwd_col.data_ptr(),
                             self.bwd_ptr.data_ptr(), self.bwd_col.data_ptr(), C.byref(lc), x.data_ptr(), self.solver.stream())


class InitialValueSolver:
    """Drop-in for the reference InitialValueSolver on the Cartesian IVP hot path."""

    def __init__(self, problem, timestepper, enforce_real_cadence=100, warmup_iterations=10, entry_cutoff=1e-12, **kw):
        if kw:
            unknown = set(kw) - {"ncc_cutoff", "max_ncc_terms", "matsolver", "bc_top", "tau_left", "interleave_components",
                                 "store_expanded_matrices", "profile", "parallel_profile"}
            if unknown:
                raise ValueError(f"Unknown solver options: {sorted(unknown)}")
        self.problem = problem
        self.dist = problem.dist
        self.dtype = problem.dtype
        self.state = problem.variables
        self.enforce_real_cadence = enforce_real_cadence
        self.warmup_iterations = warmup_iterations
        if isinstance(timestepper, str):
            timestepper = ts.schemes[timestepper]
        self.timestepper_class = timestepper
        self.sim_time = self.initial_sim_time = 0.0
        self.iteration = self.initial_iteration = 0
        self.dt = None
        self.stop_sim_time = np.inf
        self.stop_wall_time = np.inf
        self.stop_iteration = np.inf
        self.start_time = time.time()
        # ---- host setup: templates and batches (replaces Subproblem.build_matrices loops, subsystems.py:72-81)
        t0 = time.time()
        self.builder = PencilSystemBuilder(problem, entry_cutoff=entry_cutoff)
        self.batches = build_batches(self.builder)
        self.var_arena = Arena(self.dist, [(v.tshape, v.bases) for v in problem.variables])
        self.eq_arena = Arena(self.dist, [(tuple(cs.dim for cs in eq['tensorsig']), eq['bases']) for eq in problem.equations])
        self.total_modes = sum(int(c.valid_cols.sum()) * len(c.groups) for c in self.builder.classes.values())
        self.setup_time = time.time() - t0
        self._device_ready = False
        self.prof = None            # set to a list to collect (name, start_event, end_event, bytes) per launch
        self._lhs_key = None
        self._ts_iteration = 0
        self._dt_hist = deque([0.0] * getattr(timestepper, 'steps', 1))
        self.warmup_time = None
        self.run_time_start = None

    # ------------------------------------------------------------------------------------------------
    def stream(self):
        from .lib import current_stream
        return current_stream()

    def _init_device(self):
        import torch
        from .lib import compute_device
        self.device = compute_device()        # raises without a CUDA device: no CPU fallback
        self.lib = get_lib()
        self.dist.device = self.device
        if np.issubdtype(self.dtype, np.complexfloating):
            raise NotImplementedError("complex-dtype IVPs run through the complex pencil path (not in this build)")
        self.state_t = torch.zeros(self.var_arena.size, dtype=torch.float64, device=self.device)
        self.eq_t = torch.zeros(self.eq_arena.size, dtype=torch.float64, device=self.device)
        self.state_views = []
        for v, off, (tsh, shp) in zip(self.state, self.var_arena.offsets, self.var_arena.shapes):
            n = int(np.prod(tsh, dtype=int)) * int(np.prod(shp, dtype=int))
            self.state_views.append(self.state_t[off:off + n].view(tuple(tsh) + tuple(shp)))
        from .evaluator import RHSPlan
        self.rhs_plan = RHSPlan(self)
        self.rhs_plan.set_static(self.eq_t)
        self.dbatches = [DeviceBatch(self, b) for b in self.batches]
        self._device_ready = True

    def _prepare_batches(self, a0, b0):
        cls = self.timestepper_class
        for db in self.dbatches:
            if db.prog is None:
                db.upload(a0, b0)
                db.X = db.vec()
                if cls.kind == "rk":
                    db.MX0 = db.vec()
                    db.LX = [db.vec() for _ in range(cls.stages)]
                    db.F = [db.vec() for _ in range(cls.stages)]
                else:
                    db.MX = deque(db.vec() for _ in range(cls.amax))
                    db.LX = deque(db.vec() for _ in range(cls.bmax))
                    db.F = deque(db.vec() for _ in range(cls.cmax))

    def _sync_state_to_device(self):
        """Make the state arena hold the current coefficient data of every variable (uploads host edits)."""
        for v, view in zip(self.state, self.state_views):
            v.change_layout('c')
            dev = v.device_data()
            if dev.data_ptr() != view.data_ptr():
                view.copy_(dev.reshape(view.shape))
                v.set_device_data(view, 'c')

    def _mark_state_on_device(self):
        # like the reference, state fields are left in coefficient space at their dealias scales after the
        # RHS evaluation (core/evaluator.py:116-133), so e.g. u['g'] returns the dealiased grid unless the user
        # calls change_scales(1) first (as the stock scripts do)
        for v, view in zip(self.state, self.state_views):
            v.set_device_data(view, 'c', scales=v.dealias)

    def _check_factor_info(self):
        bad = sum(int(db.info.item()) for db in self.dbatches)
        if bad:
            raise DedalusB200Error(f"{bad} pencil systems hit a zero / non-finite pivot during factorisation.")

    # ------------------------------------------------------------------------------------------------
    @property
    def proceed(self):
        if self.sim_time >= self.stop_sim_time:
            return False
        if (time.time() - self.start_time) >= self.stop_wall_time:
            return False
        if self.iteration >= self.stop_iteration:
            return False
        return True

    def step(self, dt):
        """Advance the system by one timestep (reference solvers.py:683-711)."""
        if not np.isfinite(dt):
            raise ValueError("Invalid timestep")
        if not self._device_ready:
            self._init_device()
        if self.iteration == self.initial_iteration + self.warmup_iterations:
            self.run_time_start = self._sync_clock()
            self.warmup_time = self.run_time_start - self.start_time
        self.dt = dt
        if self.timestepper_class.kind == "rk":
            self._step_rk(dt)
        else:
            self._step_multistep(dt)
        self.iteration += 1

    def _sync_clock(self):
        import torch
        if self.device.type == 'cuda':
            torch.cuda.synchronize()
        return time.time()

    # ---- Runge-Kutta IMEX (reference timesteppers.py:552-644) -------------------------------------------
    def _step_rk(self, dt):
        cls = self.timestepper_class
        A, H, c = cls.A, cls.H, cls.c
        k = dt
        update = (k != self._lhs_key)
        self._lhs_key = k
        self._prepare_batches(1.0, k * H[1, 1])
        self._sync_state_to_device()
        sim_time_0 = self.sim_time
        if update:
            # one factorisation per distinct H_ii (RK222 and RK443 share a single one across stages)
            self._stage_keys = []
            done = {}
            for i in range(1, cls.stages + 1):
                hii = float(H[i, i])
                if hii not in done:
                    done[hii] = True
                    for db in self.dbatches:
                        db.factor(hii, 1.0, k * hii)
                self._stage_keys.append(hii)
            self._check_factor_info()
        for db in self.dbatches:
            db.gather('cols', self.state_t, db.X)
            db.matvec(db.X, y_m=db.MX0, y_l=db.LX[0])
        for i in range(1, cls.stages + 1):
            if i > 1:
                for db in self.dbatches:
                    db.matvec(db.X, y_l=db.LX[i - 1])
            self.rhs_plan.evaluate(self.eq_t)
            for db in self.dbatches:
                db.gather('rows', self.eq_t, db.F[i - 1])
                terms = [(db.MX0, 1.0)]
                for j in range(i):
                    if A[i, j] != 0:
                        terms.append((db.F[j], k * float(A[i, j])))
                    if H[i, j] != 0:
                        terms.append((db.LX[j], -k * float(H[i, j])))
                db.solve(db.lu[self._stage_keys[i - 1]], terms, db.X)
                db.scatter(db.X, self.state_t)
            self._mark_state_on_device()
            self.sim_time = sim_time_0 + k * c[i]

    # ---- multistep IMEX (reference timesteppers.py:95-187) ----------------------------------------------
    def _step_multistep(self, dt):
        cls = self.timestepper_class
        self._dt_hist.rotate()
        self._dt_hist[0] = dt
        a, b, c = cls.compute_coefficients(list(self._dt_hist), self._ts_iteration)
        self._ts_iteration += 1
        self._prepare_batches(a[0], b[0])
        self._sync_state_to_device()
        key = (float(a[0]), float(b[0]))
        update = (key != self._lhs_key)
        self._lhs_key = key
        if update:
            for db in self.dbatches:
                db.lu_cur = db.factor('ms', a[0], b[0])
            self._check_factor_info()
        for db in self.dbatches:
            db.MX.rotate(); db.LX.rotate(); db.F.rotate()
            db.gather('cols', self.state_t, db.X)
            db.matvec(db.X, y_m=db.MX[0], y_l=db.LX[0])
        self.rhs_plan.evaluate(self.eq_t)
        for db in self.dbatches:
            db.gather('rows', self.eq_t, db.F[0])
            terms = []
            for j in range(1, len(c)):
                if c[j] != 0:
                    terms.append((db.F[j - 1], float(c[j])))
            for j in range(1, len(a)):
                if a[j] != 0:
                    terms.append((db.MX[j - 1], -float(a[j])))
            for j in range(1, len(b)):
                if b[j] != 0:
                    terms.append((db.LX[j - 1], -float(b[j])))
            db.solve(db.lu_cur, terms, db.X)
            db.scatter(db.X, self.state_t)
        self._mark_state_on_device()
        self.sim_time += dt

    # ------------------------------------------------------------------------------------------------
    def evolve(self, timestep_function, log_cadence=100):
        try:
            while self.proceed:
                self.step(timestep_function())
        finally:
            self.log_stats()

    def log_stats(self, format=".4g"):
        """Timing summary (reference solvers.py:755-778): mode-stages per second of run time."""
        end = self._sync_clock() if self._device_ready else time.time()
        stats = {"setup_time": self.setup_time, "iterations": self.iteration - self.initial_iteration}
        if self.run_time_start is not None:
            run_time = end - self.run_time_start
            its = self.iteration - self.initial_iteration - self.warmup_iterations
            stages = getattr(self.timestepper_class, 'stages', 1)
            stats.update(warmup_time=self.warmup_time, run_time=run_time,
                         speed_mode_stages_per_sec=self.total_modes * stages * max(its, 0) / max(run_time, 1e-30))
        self.stats = stats
        return stats
