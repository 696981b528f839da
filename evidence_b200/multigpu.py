"""
Multi-GPU sharding of the likelihood batch: one process per GPU (``torch.distributed``), rows of
theta split into contiguous blocks, per-rank lnL vectors gathered with one all-gather
(NCCL over NVLink on the GPU box; gloo on CPU for the host-logic tests).

The path has no other exchange step: every parameter vector's lnL depends only on its own row and
on the small, replicated epoch data.  Rank r owns rows
``[r*ceil(B/R), min(B, (r+1)*ceil(B/R)))``; the tail is padded so that the collective has equal
counts on every rank.

Two ways to do the gather:

* ``ShardedLikelihood`` -- the likelihood kernel, then ``all_gather_into_tensor`` (NCCL).
* ``SharedHostGather`` -- for HOST consumers (a sampler running one process per GPU): every rank's
  kernel stores its lnL block straight into a host memory segment shared by the ranks' processes
  (POSIX shared memory, mapped and CUDA-registered by each of them); no device-side gathered vector,
  no D2H copy of world x rows values per rank (``rvl_loglike_scatter_host`` / ``rvl_wait_host_flags``).
* ``FusedGatherLikelihood`` -- the all-gather is fused into the producing kernel: the work item
  that finishes a point's lnL stores it straight into every rank's gathered vector through
  NVLink peer-mapped (symmetric) memory, and the launch's last block then stores a sequence
  number into a completion slot of every peer buffer; each rank waits (one warp, stream-ordered)
  for the slots of its own buffer (``rvl_loglike_dev_gather``).  No collective, no barrier, no
  staging copy on these 4 KB - 10 MB messages.
"""
import math

import numpy as np


def shard_bounds(B, world_size, rank):
    """Row block of ``rank``: (lo, hi, per) with per = ceil(B / world_size)."""
    per = int(math.ceil(B / world_size)) if B > 0 else 0
    lo = min(B, rank * per)
    hi = min(B, lo + per)
    return lo, hi, per


class ShardedLikelihood:
    """
    Wraps a per-rank evaluator ``local_eval(theta_block) -> lnL_block`` (on the GPU box:
    ``RVModel.log_likelihood_device`` on this rank's device) and presents the whole-batch call
    a sampler makes: every rank passes the same ``theta[B, ndim]`` and receives the full
    ``lnL[B]``.
    """

    def __init__(self, local_eval, ndim, device=None, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.local_eval = local_eval
        self.ndim = ndim
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device

    def __call__(self, theta):
        """theta: torch tensor [B, ndim] (float64, on ``device``), identical on every rank."""
        import torch
        B = theta.shape[0]
        lo, hi, per = shard_bounds(B, self.world, self.rank)
        local = torch.full((per,), float("nan"), dtype=torch.float64, device=theta.device)
        if hi > lo:
            local[: hi - lo] = self.local_eval(theta[lo:hi].contiguous())
        if self.world == 1:
            return local[:B]
        out = torch.empty(per * self.world, dtype=torch.float64, device=theta.device)
        self.dist.all_gather_into_tensor(out, local, group=self.group)
        return out[:B]

    def evaluate_local(self, theta_block):
        """Weak-scaling use (bench, parameter sweeps): this rank's own block, then the gather."""
        import torch
        local = self.local_eval(theta_block)
        if self.world == 1:
            return local
        out = torch.empty(local.numel() * self.world, dtype=torch.float64, device=local.device)
        self.dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out


def split_rows(theta, world_size):
    """Host-side helper: the list of row blocks the ranks own (for tests / CPU tools)."""
    theta = np.asarray(theta)
    B = theta.shape[0]
    return [theta[slice(*shard_bounds(B, world_size, r)[:2])] for r in range(world_size)]


class FusedGatherLikelihood:
    """
    Weak-scaling evaluator whose all-gather is fused into the likelihood kernel (see the module
    docstring).  Every rank calls ``evaluate_local(theta_block)`` with its own block of the same
    row count and receives the gathered ``lnL[world * rows]`` (a view into symmetric memory that
    stays valid until this rank's NEXT call: two buffers alternate, and a peer can only be two
    exchanges ahead once this rank has entered the next one).

    ``signal="flags"`` (default): the launch itself tells the peers when it is done -- its last
    block stores a sequence number into a completion slot of every peer buffer, and a one-warp
    kernel behind it waits for all ranks' slots (``rvl_loglike_dev_gather``).
    ``signal="barrier"``: peer stores, then torch's symmetric-memory barrier.
    """

    def __init__(self, model, max_rows, group=None, signal="flags"):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if signal not in ("flags", "barrier"):
            raise ValueError("signal must be 'flags' or 'barrier'")
        self.model = model
        self.signal = signal
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.max_rows = int(max_rows)
        grp = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bufs, self.hdls, self.ptrs, self.seq = [], [], [], []
        self.flag_off = self.world * self.max_rows  # completion slots follow the gathered vector
        for _ in range(2):
            buf = symm.empty(self.flag_off + self.world, dtype=torch.float64, device=dev)
            hdl = symm.rendezvous(buf, grp)
            buf.zero_()
            self.bufs.append(buf)
            self.hdls.append(hdl)
            self.ptrs.append([int(p) for p in hdl.buffer_ptrs])
            self.seq.append(0)
        torch.cuda.synchronize()
        for hdl in self.hdls:
            hdl.barrier()  # every rank's slots are zero before anybody signals
        self.local = torch.empty(self.max_rows, dtype=torch.float64, device=dev)
        self.turn = 0

    def evaluate_local(self, theta_block):
        rows = theta_block.shape[0]
        if rows > self.max_rows:
            raise ValueError("block larger than the symmetric buffer")
        k = self.turn
        self.turn ^= 1
        # every lnL value is written into all ranks' buffer k at [rank*rows + i] by the kernel
        if self.signal == "flags":
            self.seq[k] += 1
            self.model.log_likelihood_device_gather(theta_block, self.local[:rows], self.ptrs[k],
                                                    self.rank, self.rank * rows, self.flag_off,
                                                    self.seq[k])
        else:
            self.model.log_likelihood_device_scatter(theta_block, self.local[:rows], self.ptrs[k],
                                                     self.rank * rows)
            self.hdls[k].barrier()  # all peers' stores have landed
        return self.bufs[k][: self.world * rows]

    def evaluate_local_host(self, theta_host, out_all_host):
        """
        The same exchange for HOST buffers in one library call (``rvl_loglike_gather``):
        ``theta_host[rows, ndim]`` (numpy, page-locked for the in-place read) ->
        ``out_all_host[world * rows]`` (numpy) holding every rank's lnL.  Synchronous.
        """
        rows = theta_host.shape[0]
        if self.signal != "flags":
            raise ValueError("the host-buffer form uses the completion flags")
        if rows != self.max_rows:
            # the completion slots sit behind world * max_rows values: a smaller block would leave a
            # gap between the ranks' blocks, which the one D2H copy of the call does not skip
            raise ValueError("evaluate_local_host needs rows == max_rows")
        k = self.turn
        self.turn ^= 1
        self.seq[k] += 1
        return self.model.log_likelihood_gather_host(theta_host, out_all_host, self.ptrs[k],
                                                     self.rank, self.flag_off, self.seq[k])


class SharedHostGather:
    """
    Weak-scaling evaluator for HOST buffers whose gather goes through a host memory segment shared
    by the ranks' processes.  Every rank calls ``evaluate_local_host(theta_host)`` with its own block
    of ``rows`` rows (page-locked for the in-place read) and receives a numpy view ``lnL[world * rows]``
    of the shared segment holding every rank's values (valid until this rank's NEXT call: two segments
    alternate, and a peer can only be two exchanges ahead once this rank has entered the next one).  Per rank and step the PCIe link carries ``rows`` doubles out (plus theta in) instead
    of ``world * rows``; the transfer overlaps the arithmetic (the kernel stores as it goes).
    """

    def __init__(self, model, rows, group=None, timeout_ms=10000):
        import ctypes
        from multiprocessing import shared_memory
        import torch.distributed as dist
        from . import _abi
        self.model, self.rows, self.timeout_ms = model, int(rows), int(timeout_ms)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.lib = _abi.load()
        nbytes = (self.world * self.rows + self.world) * 8
        names = [None, None]
        self._owned = []
        if self.rank == 0:
            try:
                for k in range(2):
                    shm = shared_memory.SharedMemory(create=True, size=nbytes)
                    self._owned.append(shm)
                    names[k] = shm.name
            except OSError:
                names = [None, None]  # every rank raises together, below
        dist.broadcast_object_list(names, src=0, group=group)
        self.shm, self.data, self.flags, self.dev, self._addr, self.seq = [], [], [], [], [], [0, 0]
        err = None
        try:
            if names[0] is None:
                raise RuntimeError("rank 0 could not create the shared memory segments")
            for k in range(2):
                if self.rank == 0:
                    shm = self._owned[k]
                else:
                    shm = shared_memory.SharedMemory(name=names[k])
                    try:  # the creator unlinks; an attaching process must not track the segment
                        from multiprocessing import resource_tracker
                        resource_tracker.unregister(shm._name, "shared_memory")
                    except Exception:  # noqa: BLE001
                        pass
                self.shm.append(shm)
                arr = np.ndarray((self.world * self.rows + self.world,), dtype=np.float64, buffer=shm.buf)
                if self.rank == 0:
                    arr[:] = 0.0
                self.data.append(arr[: self.world * self.rows])
                self.flags.append(arr[self.world * self.rows:].view(np.uint64))
                addr = arr.ctypes.data
                dev = ctypes.c_uint64()
                rc = self.lib.rvl_host_register(ctypes.c_void_p(addr), nbytes, ctypes.byref(dev))
                if rc != 0:
                    raise RuntimeError("rvl_host_register: " + self.lib.rvl_last_error(None).decode())
                self._addr.append(addr)
                self.dev.append(dev.value)
        except Exception as exc:  # noqa: BLE001 -- reported after the ranks have agreed
            err = exc
        # the ranks succeed or fail TOGETHER (a rank that raised alone would leave the others waiting)
        import torch
        ok = torch.tensor([0.0 if err is not None else 1.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok) < 1.0:
            self.close()
            raise RuntimeError(f"SharedHostGather unavailable on some rank ({err!r})")
        dist.barrier(group=group)  # every rank has mapped both segments (and rank 0 zeroed them)
        self.turn = 0

    def evaluate_local_host(self, theta_host):
        import ctypes
        if theta_host.shape[0] != self.rows:
            raise ValueError("SharedHostGather needs blocks of exactly `rows` rows")
        k = self.turn
        self.turn ^= 1
        self.seq[k] += 1
        m = self.model
        rc = self.lib.rvl_loglike_scatter_host(m._h, theta_host.ctypes.data, self.rows, self.dev[k],
                                               self.rank * self.rows, self.world * self.rows,
                                               self.rank, self.seq[k])
        if rc != 0:
            m._check(rc)
        rc = self.lib.rvl_wait_host_flags(ctypes.c_void_p(self.flags[k].ctypes.data), self.world,
                                          self.seq[k], self.timeout_ms)
        if rc != 0:
            raise RuntimeError("rvl_wait_host_flags: " + self.lib.rvl_last_error(None).decode())
        return self.data[k]

    def close(self):
        for addr in self._addr:
            self.lib.rvl_host_unregister(__import__("ctypes").c_void_p(addr))
        self._addr = []
        self.data, self.flags = [], []
        for shm in self.shm:
            try:
                shm.close()
            except BufferError:
                pass
        for shm in self._owned:
            try:
                shm.unlink()
            except FileNotFoundError:
                pass
        self.shm, self._owned = [], []
