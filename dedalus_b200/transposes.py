"""Distributed pencil transposes between the two layouts of a 1-D process mesh (X1).

Reference: Transpose (core/distributor.py:696-924) driving FFTWTranspose (core/transposes.pyx:22-246): every rank
copies its slab into a (N1, n2_local, N0, N3) buffer, FFTW-MPI exchanges blocks (MPI alltoall inside FFTW) and the
result is copied out again; with GROUP_TRANSPOSES all fields sharing a shape travel together (842-871).
Three realisations, fastest first:
  * PeerExchange (this file) + db_rfft_*_peer: the Fourier pass in front of the hop stores every peer's rows straight into that
    peer's receive buffer over NVLink (no pack / unpack, no communication kernel; evaluator._backward_blocked_levels);
  * blocked row addressing + NCCL all-to-all (no pack / unpack; the all-to-all overlaps the neighbouring passes);
  * general fallback, any sizes: pack kernel (per-destination contiguous blocks) -> one NCCL all-to-all for the whole stack of
    fields -> unpack kernel (csrc/pointwise.cu: db_transpose_*).  The planner keeps the reference's method names:
    localize_columns (towards grid space) / localize_rows (towards coefficient space).
"""
import ctypes as C
import numpy as np
from .lib import get_lib, current_stream


class PeerExchange:
    """Receive buffers of a pencil transpose that every GPU of the box can STORE into (NVLink peer memory): the Fourier pass in
    front of a hop writes each peer's rows straight into that peer's buffer (db_rfft_*_peer), so the all-to-all is the
    transform's own stores and no communication kernel runs; `barrier()` (signal pads, device side, stream ordered) then orders
    the consumers after all writers.  Buffers come in pairs used alternately: a rank reaches the barrier of use k + 1 only
    after its own reads of use k (same stream), so nobody overwrites a buffer that is still being read.
    Built on torch's symmetric-memory allocator (plumbing: allocation + handle exchange + signal pads); unavailable -> None
    and the NCCL all-to-all path is used.  Replaces the FFTW-MPI exchange of core/transposes.pyx:173-192."""

    def __init__(self, dist):
        import torch
        import torch.distributed as td
        import torch.distributed._symmetric_memory as symm
        self.symm, self.td, self.torch = symm, td, torch
        self.P, self.rank = dist.size, td.get_rank()
        self.group = td.group.WORLD
        self._bufs = {}
        self._use = {}

    @staticmethod
    def create(dist):
        import os
        mode = os.environ.get("DB_PEER_TRANSPOSE", "1")
        # measured on 2 and 4 GPUs (profiles/README.md); on 8 the NCCL all-to-all path is the measured one, so the peer stores are
        # opt-in there (DB_PEER_TRANSPOSE=force) until they have run on that many GPUs
        if mode == "0" or dist.size > 8 or (dist.size > 4 and mode != "force"):
            return None
        try:
            import torch.distributed as td
            if td.get_backend() != "nccl":
                return None
            px = PeerExchange(dist)
            px.buffers('probe', 256)               # allocation + rendezvous must work on this box
            ok = px.torch.ones(1, device='cuda')
            td.all_reduce(ok, op=td.ReduceOp.MIN)
            return px if float(ok.item()) == 1.0 else None
        except Exception as exc:                   # e.g. no P2P / handle exchange not permitted in this container
            import warnings
            warnings.warn(f"peer-memory transposes unavailable ({type(exc).__name__}: {exc}); using the NCCL all-to-all")
            try:       # keep the ranks' collectives matched: everybody learns that somebody failed
                t = __import__('torch').zeros(1, device='cuda'); td.all_reduce(t, op=td.ReduceOp.MIN)
            except Exception:
                pass
            return None

    def buffers(self, key, numel):
        """(local tensor, [base pointer on every rank]) of the next buffer of pair `key` (symmetric: same size on all ranks)."""
        pair = self._bufs.get(key)
        if pair is None or pair[0][0].numel() < numel:
            pair = []
            for _ in range(2):
                t = self.symm.empty(int(numel), dtype=self.torch.float64, device=self.torch.device('cuda', self.torch.cuda.current_device()))
                h = self.symm.rendezvous(t, self.group)
                pair.append((t, h, [int(p) for p in h.buffer_ptrs]))
            self._bufs[key] = pair
            self._use[key] = 0
        k = self._use[key] & 1
        self._use[key] += 1
        t, h, ptrs = pair[k]
        self._last = h
        return t, ptrs, h

    def barrier(self, handle):
        handle.barrier(channel=0)


class TransposePlanner:
    def __init__(self, dist):
        import torch.distributed as td
        self.dist = dist
        self.P = dist.size
        self.td = td
        self._bufs = {}

    def _buf(self, key, numel, like):
        import torch
        t = self._bufs.get(key)
        if t is None or t.numel() < numel or t.device != like.device:
            t = torch.empty(numel, dtype=like.dtype, device=like.device)
            self._bufs[key] = t
        return t[:numel]

    def _alltoall(self, recv, send):
        td = self.td
        if td.get_backend() == "gloo":
            # gloo has no all_to_all for CPU tensors in every build: exchange blocks with paired send/recv
            P, rank = self.P, td.get_rank()
            sb, rb = send.view(P, -1), recv.view(P, -1)
            rb[rank].copy_(sb[rank])
            reqs = []
            for p in range(P):
                if p != rank:
                    reqs.append(td.isend(sb[p].contiguous(), p))
                    reqs.append(td.irecv(rb[p], p))
            for r in reqs:
                r.wait()
        else:
            td.all_to_all_single(recv, send)

    def alltoall_async(self, recv, send):
        """Start the exchange and return a handle whose wait() orders the current stream after it (NCCL: the collective
        runs on the communicator's stream, so kernels launched meanwhile overlap it); gloo: synchronous, returns None."""
        td = self.td
        if td.get_backend() == "gloo":
            self._alltoall(recv, send)
            return None
        return td.all_to_all_single(recv, send, async_op=True)

    def localize_columns(self, a, out):
        """(B, n1_local, n2, n3) distributed along axis 1  ->  (B, n1, n2_local, n3) distributed along axis 2."""
        B, n1loc, n2, n3 = a.shape
        P = self.P
        send = self._buf('send', a.numel(), a); recv = self._buf('recv', out.numel(), a)
        lib = get_lib()
        lib.call("db_transpose_pack", a.data_ptr(), send.data_ptr(), B, n1loc, n2, n3, P, current_stream())
        self._alltoall(recv, send)
        lib.call("db_transpose_unpack", recv.data_ptr(), out.data_ptr(), B, n1loc * P, n2 // P, n3, P, current_stream())

    def localize_rows(self, a, out):
        """(B, n1, n2_local, n3) distributed along axis 2  ->  (B, n1_local, n2, n3) distributed along axis 1."""
        B, n1, n2loc, n3 = a.shape
        P = self.P
        send = self._buf('send', a.numel(), a); recv = self._buf('recv', out.numel(), a)
        lib = get_lib()
        lib.call("db_transpose_pack_rev", a.data_ptr(), send.data_ptr(), B, n1, n2loc, n3, P, current_stream())
        self._alltoall(recv, send)
        lib.call("db_transpose_unpack_rev", recv.data_ptr(), out.data_ptr(), B, n1 // P, n2loc * P, n3, P, current_stream())


class B200Transpose:
    """The reference's transpose plugin contract (core/transposes.pyx:22-246 FFTWTranspose, selected as `TransposePlanner` in
    core/distributor.py:22-30):

        plan = B200Transpose(global_shape, chunk_shape, dtype, axis, comm_sub)
        plan.localize_columns(RL, CL)      # row-local (distributed along `axis`) -> column-local (distributed along axis + 1)
        plan.localize_rows(CL, RL)         # and back

    over the pack -> NCCL all-to-all -> unpack kernels of this file.  `comm_sub` is a torch.distributed process group (None: the
    default group); arrays are numpy arrays (staged through the device) or device tensors; the block distribution is the
    reference's (whole chunks per rank, transposes.pyx:73-93) and must be even here."""

    def __init__(self, global_shape, chunk_shape, dtype, axis, comm_sub=None):
        import torch.distributed as td
        from types import SimpleNamespace
        self.global_shape = tuple(int(v) for v in global_shape)
        self.axis = axis
        self.datasize = {np.float64: 1, np.complex128: 2}[np.dtype(dtype).type]
        self.dtype = np.dtype(dtype)
        self.group = comm_sub
        self.P = td.get_world_size(comm_sub) if td.is_initialized() else 1
        self.rank = td.get_rank(comm_sub) if td.is_initialized() else 0
        if td.is_initialized() and self.P != td.get_world_size():
            raise NotImplementedError("sub-communicators (meshes with more than one distributed axis) are not supported: 1-D meshes only")
        gs = self.global_shape
        self.N0 = int(np.prod(gs[:axis], dtype=np.int64)); self.N1 = gs[axis]; self.N2 = gs[axis + 1]
        self.N3 = int(np.prod(gs[axis + 2:], dtype=np.int64)) * self.datasize
        C1, C2 = int(chunk_shape[axis]), int(chunk_shape[axis + 1])
        for N, Cc in ((self.N1, C1), (self.N2, C2)):
            if N % Cc or (N // Cc) % self.P:
                raise ValueError(f"B200Transpose needs an even block distribution: {N} entries in chunks of {Cc} over {self.P} ranks")
        self.RL_reduced_shape = (self.N0, self.N1 // self.P, self.N2, self.N3)
        self.CL_reduced_shape = (self.N0, self.N1, self.N2 // self.P, self.N3)
        self._planner = TransposePlanner(SimpleNamespace(size=self.P)) if self.P > 1 else None

    def _dev(self, a, shape):
        import torch
        from .lib import compute_device
        if torch.is_tensor(a):
            t = a
            if t.is_complex():
                t = torch.view_as_real(t)
            return t.reshape(shape), None
        host = np.ascontiguousarray(a).view(np.float64).reshape(shape)
        return torch.from_numpy(host).to(compute_device()), a

    @staticmethod
    def _store(dev, host):
        if host is not None:
            np.copyto(host, dev.cpu().numpy().view(host.dtype).reshape(host.shape))

    def localize_columns(self, RL, CL):
        rl, _ = self._dev(RL, self.RL_reduced_shape)
        cl, host = self._dev(CL, self.CL_reduced_shape)
        if self.P == 1:
            cl.copy_(rl)
        else:
            self._planner.localize_columns(rl.contiguous(), cl)
        self._store(cl, host)

    def localize_rows(self, CL, RL):
        cl, _ = self._dev(CL, self.CL_reduced_shape)
        rl, host = self._dev(RL, self.RL_reduced_shape)
        if self.P == 1:
            rl.copy_(cl)
        else:
            self._planner.localize_rows(cl.contiguous(), rl)
        self._store(rl, host)


def get_planner(dist):
    """One planner per Distributor, stored ON the distributor (a cache keyed by id(dist) could hand a stale planner to a new
    distributor that reuses the id of a garbage-collected one)."""
    planner = getattr(dist, "_transpose_planner", None)
    if planner is None:
        planner = dist._transpose_planner = TransposePlanner(dist)
    return planner


def check_divisible(dist, bases, scales):
    P = dist.size
    b0, b1 = bases[0], bases[1] if len(bases) > 1 else None
    if b0 is None or b1 is None:
        raise NotImplementedError("Distributed fields need bases along the first two axes.")
    if (b0.size // b0.group_size) % P:
        raise ValueError(f"Axis-0 groups ({b0.size // b0.group_size}) must divide evenly over {P} ranks.")
    if b1.grid_size(scales[1]) % P:
        raise ValueError(f"Axis-1 grid size ({b1.grid_size(scales[1])}) must divide evenly over {P} ranks.")


def transform_field_distributed(field, layout):
    """Coefficient <-> grid layout change of one field on a 1-D mesh:
    coeff (x local blocks) -> T(last..1) -> transpose(0 <-> 1) -> T(0) -> grid (axis 1 local blocks)
    (reference Distributor._build_layouts, distributor.py:131-175 with R=1)."""
    import torch
    from .transforms import cached_plan
    dist = field.dist
    dim = dist.dim
    if any(b is None for b in field.bases[:2]):
        if all(b is None for b in field.bases):
            field.layout = field._layout_name(layout)
            return
        raise NotImplementedError("Distributed layout changes need bases along the first two axes.")
    check_divisible(dist, field.bases, field.scales)
    planner = get_planner(dist)
    nt = len(field.tensorsig)
    scales = field.scales
    cur = field.device_data()
    B = int(np.prod(cur.shape[:nt], dtype=int))
    if layout == 'g':
        for ax in range(dim - 1, 0, -1):
            b = field.bases[ax]
            if b is None:
                continue
            plan = cached_plan(b, scales[ax])
            shp = list(cur.shape); shp[nt + ax] = plan.N
            out = torch.empty(shp, dtype=cur.dtype, device=cur.device)
            plan.backward(cur.contiguous(), out, nt + ax)
            cur = out
        n1loc, n2 = cur.shape[nt], cur.shape[nt + 1]
        n3 = int(np.prod(cur.shape[nt + 2:], dtype=int))
        out = torch.empty(cur.shape[:nt] + (n1loc * dist.size, n2 // dist.size) + cur.shape[nt + 2:], dtype=cur.dtype, device=cur.device)
        planner.localize_columns(cur.contiguous().view(B, n1loc, n2, n3), out.view(B, n1loc * dist.size, n2 // dist.size, n3))
        cur = out
        plan = cached_plan(field.bases[0], scales[0])
        shp = list(cur.shape); shp[nt] = plan.N
        out = torch.empty(shp, dtype=cur.dtype, device=cur.device)
        plan.backward(cur, out, nt)
        field.set_device_data(out, 'g')
    else:
        plan = cached_plan(field.bases[0], scales[0])
        shp = list(cur.shape); shp[nt] = plan.M
        out = torch.empty(shp, dtype=cur.dtype, device=cur.device)
        plan.forward(cur.contiguous(), out, nt)
        cur = out
        n1, n2loc = cur.shape[nt], cur.shape[nt + 1]
        n3 = int(np.prod(cur.shape[nt + 2:], dtype=int))
        out = torch.empty(cur.shape[:nt] + (n1 // dist.size, n2loc * dist.size) + cur.shape[nt + 2:], dtype=cur.dtype, device=cur.device)
        planner.localize_rows(cur.view(B, n1, n2loc, n3), out.view(B, n1 // dist.size, n2loc * dist.size, n3))
        cur = out
        for ax in range(1, dim):
            b = field.bases[ax]
            if b is None:
                continue
            plan = cached_plan(b, scales[ax])
            shp = list(cur.shape); shp[nt + ax] = plan.M
            out = torch.empty(shp, dtype=cur.dtype, device=cur.device)
            plan.forward(cur.contiguous(), out, nt + ax)
            cur = out
        field.set_device_data(cur, 'c')
