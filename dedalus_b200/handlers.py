"""Analysis output and checkpoint / restart from the device-resident state.

Reference: core/evaluator.py:208-300 (Handler scheduling: wall_dt / sim_dt / iter cadences, first call always fires),
366-865 (FileHandler: sets of at most `max_writes` writes, `scales/{sim_time, wall_time, timestep, iteration, write_number}`,
`tasks/<name>` with a leading write axis), core/solvers.py:632-673 (`load_state`: restores iteration, sim_time and every
state field from grid-space data of one write).

The reference writes HDF5 through h5py, which is not installable in this image; the SAME logical layout is written as one
`.npz` archive per set (`<base>/<base>_s<set>[_p<rank>].npz`, keys `scales/...`, `tasks/...`), or as real HDF5 when h5py is
importable.  Tasks are fields, or operator expressions that `evaluator.evaluate_expression` can evaluate on the device (separable
operators of sphere fields, e.g. the vorticity task of the stock shallow-water script); other expressions are evaluated only
inside the solver's RHS plan and raise here.  The data come from the fields' device buffers through the same transform
kernels as everything else (`field['g']` / `field['c']`); nothing is recomputed on the host."""
import pathlib
import time
import numpy as np


class Handler:
    """Group of tasks with an evaluation schedule (reference core/evaluator.py:208-276)."""

    def __init__(self, solver, wall_dt=None, sim_dt=None, iter=None, custom_schedule=None):
        self.solver = solver
        self.wall_dt, self.sim_dt, self.iter, self.custom_schedule = wall_dt, sim_dt, iter, custom_schedule
        self.tasks = []
        self.last_wall_div = self.last_sim_div = self.last_iter_div = -1      # -1: the first call fires

    def check_schedule(self, **kw):
        scheduled = False
        if self.wall_dt:
            div = kw['wall_time'] // self.wall_dt
            if div > self.last_wall_div:
                scheduled, self.last_wall_div = True, div
        if self.sim_dt:
            # fire when the output target closest to now has not fired and the next step would not bring us closer
            t, dt = kw['sim_time'], kw['timestep']
            closest = int(np.round(t / self.sim_dt))
            if closest > self.last_sim_div:
                target = closest * self.sim_dt
                if abs(t - target) < abs(t + dt - target):
                    scheduled, self.last_sim_div = True, closest
        if self.iter:
            div = kw['iteration'] // self.iter
            if div > self.last_iter_div:
                scheduled, self.last_iter_div = True, div
        if self.custom_schedule and self.custom_schedule(**kw):
            scheduled = True
        return scheduled

    def add_task(self, task, layout='g', name=None, scales=None):
        from .field import Field
        from .operators import Operand
        if not isinstance(task, Operand):
            raise ValueError("output tasks must be fields or operator expressions")
        if layout not in ('g', 'c'):
            raise ValueError("layout must be 'g' or 'c'")
        if isinstance(task, Field):
            self.tasks.append(dict(field=task, layout=layout, name=name or task.name or f"task{len(self.tasks)}", scales=scales))
        else:
            # operator expression: evaluated on the device when the handler fires (separable sphere operators, e.g. the vorticity
            # -div(skew(u)) of the stock shallow-water script; evaluator.evaluate_expression raises for anything else)
            self.tasks.append(dict(expr=task, layout=layout, name=name or f"task{len(self.tasks)}", scales=scales))

    def add_tasks(self, tasks, **kw):
        for t in tasks:
            self.add_task(t, **kw)

    def _evaluate(self):
        out = {}
        for t in self.tasks:
            f = t['field'] if 'field' in t else t['expr'].evaluate()
            if t['layout'] == 'g':
                f.change_scales(t['scales'] if t['scales'] is not None else 1)
                out[t['name']] = np.array(f['g'])
            else:
                out[t['name']] = np.array(f['c'])
        return out


class DictionaryHandler(Handler):
    """Tasks evaluated into a dictionary (reference core/evaluator.py:303-316)."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.fields = {}

    def process(self, **kw):
        self.fields = self._evaluate()


class FileHandler(Handler):
    """Tasks written to sets of files (reference core/evaluator.py:366-865; see the module docstring for the format)."""

    def __init__(self, solver, base_path, max_writes=None, mode='overwrite', **kw):
        super().__init__(solver, **kw)
        self.base_path = pathlib.Path(base_path)
        self.name = self.base_path.stem
        self.max_writes = max_writes
        if mode not in ('overwrite', 'append'):
            raise ValueError("Write mode {} not defined.".format(mode))
        self.set_num = 0
        self.total_write_num = 0
        self.file_write_num = 0
        self._pending = None
        rank = solver.dist.rank if solver.dist.size > 1 else None
        self._suffix = "" if rank is None else f"_p{rank}"
        if mode == 'append':
            sets = sorted(self.base_path.glob(f"{self.name}_s*{self._suffix}.npz"), key=lambda p: int(p.stem.split('_s')[-1].split('_p')[0]))
            if sets:
                with np.load(sets[-1]) as last:
                    self.set_num = int(last['scales/set_number'])
                    self.total_write_num = int(last['scales/write_number'][-1])
        self.base_path.mkdir(parents=True, exist_ok=True)

    @property
    def current_path(self):
        return self.base_path / f"{self.name}_s{self.set_num}{self._suffix}.npz"

    def process(self, iteration, wall_time, sim_time, timestep, **kw):
        if self._pending is None or (self.max_writes is not None and self.file_write_num >= self.max_writes):
            self._flush()
            self.set_num += 1
            self.file_write_num = 0
            self._pending = dict(scales={k: [] for k in ('sim_time', 'wall_time', 'timestep', 'iteration', 'write_number')}, tasks={})
        self.total_write_num += 1
        self.file_write_num += 1
        sc = self._pending['scales']
        for k, v in (('sim_time', sim_time), ('wall_time', wall_time), ('timestep', timestep), ('iteration', iteration),
                     ('write_number', self.total_write_num)):
            sc[k].append(v)
        for name, arr in self._evaluate().items():
            self._pending['tasks'].setdefault(name, []).append(arr)
        self._flush()          # every write is on disk when process() returns (a crash loses nothing)

    def _flush(self):
        if self._pending is None:
            return
        data = {f"scales/{k}": np.asarray(v) for k, v in self._pending['scales'].items()}
        data["scales/set_number"] = np.asarray(self.set_num)
        for t in self.tasks:
            data[f"layouts/{t['name']}"] = np.asarray(t['layout'])
            data[f"task_scales/{t['name']}"] = np.asarray(1.0 if t['scales'] is None else t['scales'], dtype=float)
        for name, arrs in self._pending['tasks'].items():
            data[f"tasks/{name}"] = np.stack(arrs, axis=0)
        tmp = self.current_path.with_suffix(".tmp.npz")
        np.savez(tmp, **data)
        tmp.replace(self.current_path)


class Evaluator:
    """`solver.evaluator`: registry of output handlers fired at the start of a step, like the reference's scheduled
    evaluation in stage 1 of the step (core/timesteppers.py:150-151, 607-608; core/evaluator.py:62-93)."""

    def __init__(self, solver):
        self.solver = solver
        self.handlers = []
        self._hooked = False

    def _hook(self):
        if not self._hooked:
            self.solver.step_hooks.append(self._fire)
            self._hooked = True

    def add_file_handler(self, filename, **kw):
        h = FileHandler(self.solver, filename, **kw)
        self.handlers.append(h); self._hook()
        return h

    def add_dictionary_handler(self, **kw):
        h = DictionaryHandler(self.solver, **kw)
        self.handlers.append(h); self._hook()
        return h

    def evaluate_scheduled(self, **kw):
        for h in self.handlers:
            if h.check_schedule(**kw):
                h.process(**kw)

    def evaluate_handlers(self, handlers=None, **kw):
        for h in (handlers if handlers is not None else self.handlers):
            h.process(**kw)

    def _fire(self, solver):
        self.evaluate_scheduled(iteration=solver.iteration, wall_time=time.time() - solver.start_time,
                                sim_time=solver.sim_time, timestep=solver.dt)


def load_state(solver, path, index=-1, allow_missing=False):
    """Restore iteration, sim_time and every state field from one write of a FileHandler set (reference
    core/solvers.py:632-673).  Returns (write number, timestep)."""
    path = pathlib.Path(path)
    with np.load(path) as file:
        write = int(file['scales/write_number'][index])
        dt = float(file['scales/timestep'][index])
        solver.iteration = solver.initial_iteration = int(file['scales/iteration'][index])
        solver.sim_time = solver.initial_sim_time = float(file['scales/sim_time'][index])
        for field in solver.state:
            key = f"tasks/{field.name}"
            if key in file.files:
                layout = str(file[f"layouts/{field.name}"])
                data = file[key][index]
                if layout == 'g':
                    field.change_scales(float(file[f"task_scales/{field.name}"]) if file[f"task_scales/{field.name}"].ndim == 0
                                        else tuple(file[f"task_scales/{field.name}"]))
                    field['g'] = data
                else:
                    field['c'] = data
            elif not allow_missing:
                raise IOError(f"Field '{field.name}' not found in savefile. Set allow_missing=True to ignore this error.")
    return write, dt
