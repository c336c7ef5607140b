"""InitialValueSolver with a device-resident IMEX stage loop.

Reference: core/solvers.py:503-806 (InitialValueSolver; `step` 683-711) driving core/timesteppers.py:95-187
(MultistepIMEX.step) and 552-644 (RungeKuttaIMEX.step), whose bodies are Python loops over pencils.
Here one step is a fixed sequence of kernel launches per *batch* of structurally identical pencil systems
(dedalus_b200/pencils.py): gather -> template mat-vecs (M.X, L.X) -> RHS evaluation (evaluator.RHSPlan)
-> fused RHS-combination + triangular solves -> scatter.  Factorisations are rebuilt only when the LHS
coefficients (a0, b0) change, exactly as the reference drops its LHS_solvers (timesteppers.py:135-140, 577-583).
"""
import os
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


def matvec_windows(prog, R, wmax):
    """Per block of R rows: [first column, count] of the x window staged in shared memory by k_batches_matvec -- the
    densest run of at most `wmax` columns among the terms of the block's rows (both operators); columns of the dense
    boundary rows mostly fall outside and are read from global memory."""
    n = prog.n
    out = np.zeros((-(-n // R), 2), dtype=np.int32)
    for rb in range(out.shape[0]):
        r0, r1 = rb * R, min(n, (rb + 1) * R)
        cols = np.concatenate([np.asarray(prog.mv[k][1][prog.mv[k][0][r0]:prog.mv[k][0][r1]], dtype=np.int64) for k in ('M', 'L')])
        if cols.size == 0:
            continue
        cols.sort()
        # window start maximising the number of terms covered
        hi = np.searchsorted(cols, cols + wmax, side='left')
        best = int(np.argmax(hi - np.arange(cols.size)))
        w0 = int(cols[best]); w1 = min(n, w0 + wmax)
        w1 = min(w1, int(cols[-1]) + 1)
        out[rb] = (w0, w1 - w0)
    return out


class DeviceBatch:
    """Device copies of one batch's programs and work vectors (kept alive for the fused descriptors)."""

    def __init__(self, solver, batch, a0, b0, nslots, vecs=None):
        import torch
        self.batch, self.solver = batch, solver
        dev = solver.device
        prog = compile_batch(batch, a0, b0)
        self.prog = prog
        self.n, self.S = prog.n, prog.S
        self.R = batch.R                # sign-equivalent members sharing each pencil's factorisation (pencils.merge_sign_equivalent)
        self.ld = ld = prog.ld          # pencils padded to whole tiles of 64 (tile-major storage, see include/dedalus_b200.h)
        mono = np.zeros((len(prog.monos), ld)); mono[:, :prog.S] = prog.mono_vals
        f = lambda a: _i32(torch, a, dev)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        self.t = dict(mono=d(mono), prog=f(prog.prog), ctrl=f(prog.ctrl.ravel()), diag_eid=f(prog.diag_eid), fl_ptr=f(prog.fl_ptr), fl_eid=f(prog.fl_eid),
                      fu_ptr=f(prog.fu_ptr), fu_eid=f(prog.fu_eid), fd_eid=f(prog.fd_eid))
        win = matvec_windows(prog, 32, 80)
        for name in ('M', 'L'):
            ptr, col, mono_i, val = prog.mv[name]
            k = name.lower()
            self.t[k + '_ptr'], self.t[k + '_col'], self.t[k + '_mono'], self.t[k + '_val'] = f(ptr), f(col), f(mono_i), d(val)
            # fused kernel: 16-byte records, each row's terms reordered so that those inside the row block's shared-memory
            # window come first (col_off = window row) and the others (col_off = column * tile, read from global memory) last
            ptr_a, col_a, mono_a, val_a = (np.asarray(a) for a in (ptr, col, mono_i, val))
            rec = np.zeros(len(col_a), dtype=np.dtype([('val', '<f8'), ('col_off', '<i4'), ('mono', '<i4')]))     # db_term
            split = np.zeros(prog.n, dtype=np.int32)
            for i in range(prog.n):
                t0, t1 = int(ptr_a[i]), int(ptr_a[i + 1])
                w0, wl = win[i // 32]
                cols = col_a[t0:t1].astype(np.int64)
                inside = (cols >= w0) & (cols < w0 + wl)
                order = np.concatenate([np.nonzero(inside)[0], np.nonzero(~inside)[0]])
                n_in = int(inside.sum())
                rec['val'][t0:t1] = val_a[t0:t1][order]
                rec['mono'][t0:t1] = mono_a[t0:t1][order]
                rec['col_off'][t0:t0 + n_in] = cols[order[:n_in]] - w0
                rec['col_off'][t0 + n_in:t1] = cols[order[n_in:]] * prog.tile
                split[i] = t0 + n_in
            self.t[k + '_rec'] = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy()).to(dev)
            self.t[k + '_split'] = f(split)
        self.t['mv_win'] = f(win.ravel())
        self.maps = []
        for side, arena in (('cols', solver.var_arena), ('rows', solver.eq_arena)):
            m = line_maps(batch, arena, side)
            so = np.zeros((m.sys_off.shape[0], ld), dtype=np.int64); so[:, :prog.S] = m.sys_off
            self.maps.append(dict(nlines=m.line_base.shape[1], max_len=int(m.line_len.max()) if len(m.line_len) else 0,
                                  base=torch.from_numpy(np.ascontiguousarray(m.line_base)).to(dev), sign=d(m.line_sign),
                                  kind=f(m.line_kind), ptr=f(m.line_ptr), pos=f(m.line_pos), sys_off=torch.from_numpy(so).to(dev)))
        ptr, mono_i, val = assembly_program(batch, prog, a0, b0)
        self.t['asm_ptr'], self.t['asm_mono'], self.t['asm_val'] = f(ptr), f(mono_i), d(val)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        # work vectors survive a re-ordering of the batch (row-space vectors are unaffected by a new pivot column order)
        # allocated for a whole number of member groups of 4 (SOLVE_MAX_RHS): the members-in-registers solve kernel carries 2 or
        # 4 members per thread without per-member predicates; missing members are all-zero columns nobody reads
        self.R_alloc = -(-self.R // 4) * 4
        self.vecs = vecs if vecs is not None else [torch.zeros(self.n * ld * self.R_alloc, dtype=torch.float64, device=dev) for _ in range(nslots)]
        self.lu = {}

    def lu_tensor(self, slot):
        import torch
        if slot not in self.lu:
            self.lu[slot] = torch.zeros(self.prog.nE * self.ld, dtype=torch.float64, device=self.solver.device)
        return self.lu[slot]

    def set_lhs(self, a0, b0):
        import torch
        ptr, mono_i, val = assembly_program(self.batch, self.prog, a0, b0)
        self.t['asm_val'].copy_(torch.from_numpy(np.ascontiguousarray(val)))


class BatchSet:
    """All batches of a solver behind ONE device descriptor array (include/dedalus_b200.h: db_batch)."""

    # relative backward error above which a factorisation is rejected (a healthy static order gives 1e-16 .. 1e-13; the
    # unstable orders of round 1 gave 1e-4 .. 1)
    VERIFY_TOL = 1e-10

    def __init__(self, solver, a0, b0, nslots, nlu):
        from .lib import DB_MAX_VECS, DB_MAX_LU
        if nslots > DB_MAX_VECS or nlu > DB_MAX_LU:
            raise NotImplementedError("too many work vectors / factor sets for the fused batch descriptor")
        self.solver = solver
        self.nslots, self.nlu = nslots, nlu
        self.items = [DeviceBatch(solver, b, a0, b0, nslots) for b in solver.batches]
        self.nb = len(self.items)
        self.reorders = 0            # number of batches whose pivot order had to be recomputed after a failed verification
        self.last_verify = None
        self._build_descriptors()

    def _build_descriptors(self):
        import torch
        from .lib import Batch as CBatch
        solver, nlu = self.solver, self.nlu
        arr = (CBatch * max(self.nb, 1))()
        blk = dict(solve=0, matvec=0, move0=0, move1=0, asm=0)
        for i, db in enumerate(self.items):
            c = arr[i]
            t_ = db.t
            c.n, c.S, c.ld, c.n_entries = db.n, db.S, db.ld, db.prog.nE
            c.nrhs = db.R
            c.n_fwd, c.n_bwd = db.prog.n_fwd, db.prog.nE - db.prog.n_fwd
            c.blk_solve = blk['solve']; blk['solve'] += (db.S + 63) // 64
            # mat-vec CTAs: one 64-system tile x 32 consecutive rows, x window in shared memory (csrc/pencil.cu k_batches_matvec)
            c.mv_rows = 32
            c.mv_win = t_['mv_win'].data_ptr()
            c.blk_matvec = blk['matvec']; blk['matvec'] += ((db.S + 63) // 64) * db.R * (-(-db.n // 32))
            c.blk_assemble = blk['asm']; blk['asm'] += ((db.S + 127) // 128) * ((db.prog.nE + 63) // 64)
            for side in (0, 1):
                m = db.maps[side]
                c.nlines[side], c.max_len[side] = m['nlines'], m['max_len']
                c.blk_move[side] = blk[f'move{side}']
                blk[f'move{side}'] += m['nlines'] * ((db.S + 63) // 64) * db.R     # one CTA per (line, member, 64 pencils): MOVE_T in csrc/pencil.cu
                c.line_base[side], c.line_kind[side] = m['base'].data_ptr(), m['kind'].data_ptr()
                c.line_sign[side] = m['sign'].data_ptr()
                c.line_ptr[side], c.line_pos[side], c.sys_off[side] = m['ptr'].data_ptr(), m['pos'].data_ptr(), m['sys_off'].data_ptr()
            t = db.t
            c.prog, c.mono, c.ctrl = t['prog'].data_ptr(), t['mono'].data_ptr(), t['ctrl'].data_ptr()
            for j, v in enumerate(db.vecs):
                c.vec[j] = v.data_ptr()
            for j in range(nlu):
                c.lu[j] = db.lu_tensor(j).data_ptr()
            c.m_ptr, c.m_col, c.m_mono, c.m_val = (t[k].data_ptr() for k in ('m_ptr', 'm_col', 'm_mono', 'm_val'))
            c.l_ptr, c.l_col, c.l_mono, c.l_val = (t[k].data_ptr() for k in ('l_ptr', 'l_col', 'l_mono', 'l_val'))
            c.m_rec, c.l_rec, c.n_mono = t['m_rec'].data_ptr(), t['l_rec'].data_ptr(), len(db.prog.monos)
            c.m_split, c.l_split = t['m_split'].data_ptr(), t['l_split'].data_ptr()
            if len(db.prog.monos) > 16:
                raise NotImplementedError("more than 16 wavenumber monomials in one class (MV_MAX_MONO in csrc/pencil.cu)")
            c.diag_eid, c.fl_ptr, c.fl_eid = t['diag_eid'].data_ptr(), t['fl_ptr'].data_ptr(), t['fl_eid'].data_ptr()
            c.fu_ptr, c.fu_eid, c.fd_eid = t['fu_ptr'].data_ptr(), t['fu_eid'].data_ptr(), t['fd_eid'].data_ptr()
            c.asm_ptr, c.asm_mono, c.asm_val = t['asm_ptr'].data_ptr(), t['asm_mono'].data_ptr(), t['asm_val'].data_ptr()
            c.info = db.info.data_ptr()
        self.blocks = blk
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.desc = torch.from_numpy(raw).to(solver.device)
        self.host_desc = arr
        # byte counts for the roofline accounting
        self.sum_nS = sum(db.n * db.S * db.R for db in self.items)      # unknowns of all systems
        self.sum_ES = sum(db.prog.nE * db.S for db in self.items)        # stored factor entries (one set per pencil)
        self.max_nrhs = max((db.R for db in self.items), default=1)

    def _call(self, name, *args):
        self.solver.lib.call(name, self.desc.data_ptr(), self.nb, *args, self.solver.stream())

    def move(self, side, gather, slot, arena_t):
        with Timed(self.solver.prof, "pencil_gather" if gather else "pencil_scatter", 16 * self.sum_nS):
            self._call("db_batches_move", self.blocks[f'move{side}'], side, 1 if gather else 0, slot, arena_t.data_ptr())

    def matvec(self, x_slot, ym_slot=-1, yl_slot=-1):
        nout = (ym_slot >= 0) + (yl_slot >= 0)
        with Timed(self.solver.prof, "pencil_matvec", 8 * self.sum_nS * (1 + nout)):
            self._call("db_batches_matvec", self.blocks['matvec'], x_slot, ym_slot, yl_slot)

    def solve(self, lu_slot, x_slot, terms):
        from .lib import SlotComb
        sc = SlotComb()
        sc.nvec = len(terms)
        for j, (slot, coef) in enumerate(terms):
            sc.slot[j] = slot; sc.coef[j] = coef
        # algorithmic bytes: every stored LU entry once + each RHS vector once + the solution written once
        with Timed(self.solver.prof, "pencil_solve", 8 * (self.sum_ES + self.sum_nS * (len(terms) + 1))):
            self._call("db_batches_solve", self.blocks['solve'], self.max_nrhs, lu_slot, x_slot, C.byref(sc))

    def factor(self, lu_slot, a0, b0):
        for db in self.items:
            db.set_lhs(a0, b0)
            db.info.zero_()
        self._call("db_batches_assemble", self.blocks['asm'], lu_slot)
        self._call("db_batches_factor", self.blocks['solve'], lu_slot)

    def check_info(self):
        bad = sum(int(db.info.item()) for db in self.items)
        if bad:
            raise DedalusB200Error(f"{bad} pencil systems hit a zero / non-finite pivot during factorisation.")

    # ---- verification of the shared pivot order (every system, after every factorisation) -------------------
    def probe(self, lu_slot, a0, b0, slots, seed=1234):
        """Relative backward error of  x = LU^{-1} b  for a random b, per system: list over batches of numpy arrays.
        slots = (b, x, Mx, Lx) work-vector slots that may be overwritten."""
        import torch
        s_b, s_x, s_m, s_l = slots
        dev = self.solver.device
        gen = torch.Generator(device=dev); gen.manual_seed(seed)
        for db in self.items:
            db.vecs[s_b].normal_(generator=gen)
        self.solve(lu_slot, s_x, [(s_b, 1.0)])
        self.matvec(s_x, s_m, s_l)
        nblk = self.blocks['solve']
        out = torch.zeros(nblk * 64, dtype=torch.float64, device=dev)
        self._call("db_batches_residual", nblk, s_b, s_m, s_l, float(a0), float(b0), out.data_ptr())
        host = out.cpu().numpy()
        res, start = [], 0
        for db in self.items:
            res.append(host[start:start + db.S].copy())
            start += ((db.S + 63) // 64) * 64
        return res

    def factor_verified(self, lhs, slots):
        """lhs: list of (lu_slot, a0, b0).  Factorise every a0 M + b0 L into its slot and verify every system of every batch
        (probe()).  Batches with a member above VERIFY_TOL get a new pivot order -- true partial pivoting on the
        representatives plus the offending members, jointly at every LHS seen so far -- and all sets are factorised again;
        raises if that does not help.  Returns the worst backward error (also kept in `last_verify`)."""
        lhs_seen = self.__dict__.setdefault('_lhs_seen', [])
        for _, a0, b0 in lhs:
            if (float(a0), float(b0)) not in lhs_seen:
                lhs_seen.append((float(a0), float(b0)))
        del lhs_seen[:-6]
        for attempt in range(3):
            worst, failing = 0.0, {}
            for lu_slot, a0, b0 in lhs:
                self.factor(lu_slot, a0, b0)
                self.check_info()
                for i, r in enumerate(self.probe(lu_slot, a0, b0, slots)):
                    if not r.size:
                        continue
                    r = np.nan_to_num(r, nan=np.inf)
                    worst = max(worst, float(r.max()))
                    if not r.max() <= self.VERIFY_TOL:
                        bad = [int(j) for j in np.argsort(-r)[:4] if r[int(j)] > self.VERIFY_TOL]
                        failing.setdefault(i, []).extend(bad)
            self.last_verify = worst
            if not failing:
                return worst
            if attempt == 2:
                break
            a0, b0 = lhs[0][1], lhs[0][2]
            for i, members in failing.items():
                db = self.items[i]
                extra = getattr(db.batch, 'extra_groups', []) + [db.batch.groups[j] for j in sorted(set(members))]
                db.batch.extra_groups = extra[-8:]
                db.batch.compute_ordering(a0, b0, threshold=1.0, extra_groups=db.batch.extra_groups,
                                          extra_lhs=[ab for ab in lhs_seen if ab != (float(a0), float(b0))])
                self.items[i] = DeviceBatch(self.solver, db.batch, a0, b0, self.nslots, vecs=db.vecs)
                self.reorders += 1
            self._build_descriptors()
        raise DedalusB200Error(f"pencil factorisation failed verification: backward error {worst:.2e} > {self.VERIFY_TOL:.0e} "
                               f"in batches {sorted(failing)} after re-ordering.")


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
        from .sphere import sphere_basis_of
        self.entry_cutoff = entry_cutoff
        # curvilinear problems (S2 sphere): per-m banded systems, dedalus_b200/sphere.py SphereSystems, built on the device side
        from .shell import shell_basis_of
        self.shell = any(shell_basis_of(v) is not None for v in problem.variables)
        self.curvilinear = self.shell or any(sphere_basis_of(v) is not None for v in problem.variables)
        if self.curvilinear:
            self.builder, self.batches = None, []
        else:
            self.builder = PencilSystemBuilder(problem, entry_cutoff=entry_cutoff)
            self.batches = build_batches(self.builder)
        self.complex = bool(np.issubdtype(self.dtype, np.complexfloating))
        lead = (2,) if self.complex else ()      # complex fields live in the arenas as a real and an imaginary plane (pencils.py)
        self.var_arena = Arena(self.dist, [(lead + tuple(v.tshape), v.bases) for v in problem.variables])
        self.eq_arena = Arena(self.dist, [(lead + tuple(cs.dim for cs in eq['tensorsig']), eq['bases']) for eq in problem.equations])
        self.total_modes = 0 if self.curvilinear else sum(int(c.valid_cols.sum()) * len(c.groups) for c in self.builder.classes.values())
        self.setup_time = time.time() - t0
        self._device_ready = False
        self.step_hooks = []        # callables(solver) run at the start of each step (CFL, flow properties, output handlers)
        from .handlers import Evaluator
        self.evaluator = Evaluator(self)      # solver.evaluator.add_file_handler(...).add_task(...) as in the reference
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
        if self.complex and (self.curvilinear or self.dist.size > 1):
            raise NotImplementedError("complex-dtype problems are single-GPU Cartesian in this build")
        self.state_t = torch.zeros(self.var_arena.size, dtype=torch.float64, device=self.device)
        self.eq_t = torch.zeros(self.eq_arena.size, dtype=torch.float64, device=self.device)
        self.state_views = []
        for v, off, (tsh, shp) in zip(self.state, self.var_arena.offsets, self.var_arena.shapes):
            n = int(np.prod(tsh, dtype=int)) * int(np.prod(shp, dtype=int))
            self.state_views.append(self.state_t[off:off + n].view(tuple(tsh) + tuple(shp)))
        if self.curvilinear:
            # the fused plans cover products of state fields and their separable derivatives; anything else (grid functions,
            # forcings, radial profiles, operators of products) goes through the general evaluator, equation by equation
            try:
                if self.shell:
                    from .shell_ivp import ShellRHSPlan
                    self.rhs_plan = ShellRHSPlan(self)
                else:
                    from .sphere import SphereRHSPlan
                    self.rhs_plan = SphereRHSPlan(self)
            except NotImplementedError:
                from .analysis import GenericCurvilinearRHS
                self.rhs_plan = GenericCurvilinearRHS(self)
        elif self.complex:
            from .complex_path import ComplexRHSPlan
            self.rhs_plan = ComplexRHSPlan(self)
        else:
            from .evaluator import RHSPlan, StagedRHSPlan, NonPolynomialError
            try:
                self.rhs_plan = RHSPlan(self)
            except NonPolynomialError:
                # grid functions or derivatives of products on the right-hand side: evaluated in stages (evaluator.Stager)
                self.rhs_plan = StagedRHSPlan(self)
        self.rhs_plan.set_static(self.eq_t)
        self.bset = None
        self._device_ready = True

    def _prepare_batches(self, a0, b0):
        if getattr(self, 'bset', None) is not None:
            return
        cls = self.timestepper_class
        if cls.kind == "rk":
            # slots: 0 X, 1 MX0, 2.. LX[i], 2+stages.. F[i]
            s = cls.stages
            self.slot_X, self.slot_MX0 = 0, 1
            self.slot_LX = [2 + i for i in range(s)]
            self.slot_F = [2 + s + i for i in range(s)]
            nslots = 2 + 2 * s
            nlu = len({float(cls.H[i, i]) for i in range(1, s + 1)})
        else:
            self.slot_X = 0
            self.slot_MX = deque(1 + j for j in range(cls.amax))
            self.slot_LX = deque(1 + cls.amax + j for j in range(cls.bmax))
            self.slot_F = deque(1 + cls.amax + cls.bmax + j for j in range(cls.cmax))
            nslots = 1 + cls.amax + cls.bmax + cls.cmax
            nlu = 1
        if self.shell:
            from .shell_ivp import ShellSystems
            self.bset = ShellSystems(self, nslots, nlu)
            self.total_modes = self.bset.total_modes
        elif self.curvilinear:
            from .sphere import SphereSystems
            self.bset = SphereSystems(self, nslots, nlu)
            self.total_modes = self.bset.total_modes
        else:
            self.bset = BatchSet(self, a0, b0, nslots, nlu)

    def _sync_state_to_device(self):
        """Make the state arena hold the current coefficient data of every variable (uploads host edits)."""
        import torch
        for v, view in zip(self.state, self.state_views):
            v.change_layout('c')
            dev = v.device_data()
            if self.complex:
                view.copy_(torch.view_as_real(dev).movedim(-1, 0))      # interleaved complex field -> re / im planes
                continue
            if dev.data_ptr() != view.data_ptr():
                view.copy_(dev.reshape(view.shape))
                v.set_device_data(view, 'c')

    def _mark_state_on_device(self):
        # like the reference, state fields are left in coefficient space at their dealias scales after the
        # RHS evaluation (core/evaluator.py:116-133), so e.g. u['g'] returns the dealiased grid unless the user
        # calls change_scales(1) first (as the stock scripts do)
        for v, view in zip(self.state, self.state_views):
            if self.complex:
                import torch
                v.set_device_data(torch.complex(view[0], view[1]), 'c', scales=v.dealias)
            else:
                v.set_device_data(view, 'c', scales=v.dealias)

    def _check_factor_info(self):
        self.bset.check_info()

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
        if self.step_hooks:
            self._prepare_hooks()
        if self.timestepper_class.kind == "rk":
            self._step_rk(dt)
        else:
            self._step_multistep(dt)
        # Hermitian-symmetry enforcement for real variables, same cadence rule as the reference (solvers.py:704-708)
        if self.enforce_real_cadence and np.issubdtype(self.dtype, np.floating):
            if self.iteration % self.enforce_real_cadence < getattr(self.timestepper_class, 'steps', 1):
                self.enforce_hermitian_symmetry(self.state)
        self.iteration += 1

    def _prepare_hooks(self):
        """Scheduled evaluations at the start of the step (the reference fires its handlers in stage 1 of the step,
        core/timesteppers.py:150-151, 607-608): state is brought to coefficient space first."""
        for v in self.state:
            v.change_layout('c')
        for hook in self.step_hooks:
            hook(self)

    def enforce_hermitian_symmetry(self, fields):
        """Transform fields to the dealiased grid and back (reference solvers.py:675-681)."""
        for f in fields:
            if all(b is None for b in f.bases):
                continue
            try:
                f.change_scales(f.dealias)
                f.change_layout('g')
                f.change_layout('c')
            except NotImplementedError as exc:
                # partially-based fields (e.g. tau fields on (x, y)) on a distributed mesh have no grid round trip yet: the
                # reference enforces them too (core/solvers.py:675-681), so say so once instead of skipping silently
                if not getattr(self, "_warned_hermitian", False):
                    import warnings
                    warnings.warn(f"Hermitian-symmetry enforcement skipped for field {f.name!r} ({exc}); "
                                  "the reference applies it to every state field")
                    self._warned_hermitian = True

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
        bs = self.bset
        if update:
            # one factorisation per distinct H_ii (RK222 and RK443 share a single one across stages), each verified on the
            # device for every system; the probe overwrites X, MX0, LX[0], F[0], all of which are rebuilt below
            self._stage_lu = []
            done = {}
            for i in range(1, cls.stages + 1):
                hii = float(H[i, i])
                if hii not in done:
                    done[hii] = len(done)
                self._stage_lu.append(done[hii])
            bs.factor_verified([(slot, 1.0, k * hii) for hii, slot in done.items()],
                               (self.slot_F[0], self.slot_X, self.slot_MX0, self.slot_LX[0]))
        bs.move(0, True, self.slot_X, self.state_t)
        bs.matvec(self.slot_X, self.slot_MX0, self.slot_LX[0])
        for i in range(1, cls.stages + 1):
            if i > 1:
                bs.matvec(self.slot_X, -1, self.slot_LX[i - 1])
            self.rhs_plan.evaluate(self.eq_t)
            bs.move(1, True, self.slot_F[i - 1], self.eq_t)
            terms = [(self.slot_MX0, 1.0)]
            for j in range(i):
                if A[i, j] != 0:
                    terms.append((self.slot_F[j], k * float(A[i, j])))
                if H[i, j] != 0:
                    terms.append((self.slot_LX[j], -k * float(H[i, j])))
            bs.solve(self._stage_lu[i - 1], self.slot_X, terms)
            bs.move(0, False, self.slot_X, self.state_t)
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
        bs = self.bset
        self.slot_MX.rotate(); self.slot_LX.rotate(); self.slot_F.rotate()
        if update:
            # the probe overwrites X and the oldest history vectors, which this step overwrites anyway
            bs.factor_verified([(0, float(a[0]), float(b[0]))], (self.slot_F[0], self.slot_X, self.slot_MX[0], self.slot_LX[0]))
        bs.move(0, True, self.slot_X, self.state_t)
        bs.matvec(self.slot_X, self.slot_MX[0], self.slot_LX[0])
        self.rhs_plan.evaluate(self.eq_t)
        bs.move(1, True, self.slot_F[0], self.eq_t)
        terms = []
        for j in range(1, len(c)):
            if c[j] != 0:
                terms.append((self.slot_F[j - 1], float(c[j])))
        for j in range(1, len(a)):
            if a[j] != 0:
                terms.append((self.slot_MX[j - 1], -float(a[j])))
        for j in range(1, len(b)):
            if b[j] != 0:
                terms.append((self.slot_LX[j - 1], -float(b[j])))
        bs.solve(0, self.slot_X, terms)
        bs.move(0, False, self.slot_X, self.state_t)
        self._mark_state_on_device()
        self.sim_time += dt

    # ------------------------------------------------------------------------------------------------
    def load_state(self, path, index=-1, allow_missing=False):
        """Restore iteration, sim_time and the state from a FileHandler set (reference core/solvers.py:632-673)."""
        from .handlers import load_state
        return load_state(self, path, index=index, allow_missing=allow_missing)

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


class _DirectSolve:
    """Placeholder 'scheme' of the boundary value solver: one factorisation of L, no history."""
    kind, steps, stages = "lbvp", 1, 0


class LinearBoundaryValueSolver(InitialValueSolver):
    """L.X = F on the device (reference LinearBoundaryValueSolver, core/solvers.py:286-375): `solve()` evaluates F, solves every
    pencil system with the factorised L and scatters the solution into the problem variables.  Real Cartesian problems (the stock
    Poisson script) and sphere problems (the balanced-height problem of the stock shallow-water script); F may read any field but
    not the unknowns."""

    rhs_reads_state = False

    def __init__(self, problem, **kw):
        super().__init__(problem, _DirectSolve, **kw)
        if self.complex:
            raise NotImplementedError("LBVPs are built for real-dtype problems in this build")

    def step(self, dt):
        raise TypeError("boundary value solvers have no time step; call solve()")

    def solve(self, rebuild_matrices=False):
        if not self._device_ready:
            self._init_device()
        if self.bset is None or rebuild_matrices:
            if self.shell:
                from .shell_ivp import ShellSystems
                self.bset = ShellSystems(self, 4, 1)          # the per-l dense systems with M = 0
                self.total_modes = self.bset.total_modes
            elif self.curvilinear:
                from .sphere import SphereSystems
                self.bset = SphereSystems(self, 4, 1)         # slots: 0 F, 1 X, 2 / 3 probe products
                self.total_modes = self.bset.total_modes
            else:
                self.bset = BatchSet(self, 0.0, 1.0, 4, 1)    # the Cartesian pencil batches with M = 0: ordering chosen for L alone
                # L alone is not diagonally dominated the way M + dt L is (curl-grad systems, pure Neumann rows): the backward error
                # a static ordering reaches is a few 1e-10 there; still far inside the accuracy the solution is used at
                self.bset.VERIFY_TOL = 1e-8
            self.bset.factor_verified([(0, 0.0, 1.0)], (0, 1, 2, 3))
        bs = self.bset
        self.rhs_plan.evaluate(self.eq_t)
        bs.move(1, True, 0, self.eq_t)
        bs.solve(0, 1, [(0, 1.0)])
        bs.move(0, False, 1, self.state_t)
        for v, view in zip(self.state, self.state_views):      # the unknowns keep their scales (reference solvers.py:397-398: preset_layout only)
            v.set_device_data(view, 'c')
        self.iteration += 1
