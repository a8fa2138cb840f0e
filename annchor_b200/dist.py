"""Multi-GPU plumbing for the sharded fit() (the reference is single-host
and has no counterpart).  One process per GPU, ``torch.distributed`` for the rendezvous and the
collectives (NCCL over NVLink on the GPUs; gloo in the CPU tests).

What is sharded: the Theta(N^2) tile sweeps.  An index created with (rank, world) sweeps only the
tiles ``t % world == rank``; X, the anchor-distance matrix and the known-pair store are replicated
(X is N*d*4 B -- 5 GB at N=10M, d=128 -- so replication costs nothing on a 180 GB part).
Exchange steps, all through this module:

  * per-row thresholds / guarantee_nmin lists : sum all-reduce of device slices (every row block is
    computed by exactly one rank, the others hold zeros)                       -> ``Comm.reducer``
  * probability-level histograms, tie-break digit histograms, branch flags : small host-buffer
    sum all-reduces                                                            -> ``Comm.reducer``
  * exactly evaluated pairs (i, j, d) of a refine round and tightened bounds (i, j, lb, ub) of an
    update_anchor_points round : variable-length all-gather, then inserted into every other rank's
    store                                                                      -> ``exchange``

libannb calls back into ``Comm.reducer`` through the C ABI hook annb_index_set_reducer
(include/annb.h); bulk data never leaves the devices.
"""
import ctypes as C

import numpy as np

RED_U64, RED_F32, RED_I32 = 0, 1, 2
RED_DEVICE = 0x100  # flag: the buffer is a device pointer on the index's GPU

REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int)

_NP = {RED_U64: np.uint64, RED_F32: np.float32, RED_I32: np.int32}


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias a raw device pointer."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


class Comm:
    """The collectives the sharded index needs, over a torch.distributed process group."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun and call "
                               "init_process_group first)")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device  # torch.device of this rank's GPU (None: CPU tensors, gloo)
        self._cb = REDUCE_FN(self._reduce_cb)  # keep the callback object alive
        self.n_reductions = 0
        self.bytes_gathered = 0

    # -- sum all-reduce used by libannb -----------------------------------------------------
    def _reduce_cb(self, user, buf, count, dtype):
        try:
            self.reduce_raw(buf, int(count), int(dtype))
            return 0
        except Exception as e:  # no exception may cross the C boundary
            import sys
            sys.stderr.write("annchor_b200.dist: reduce callback failed: %r\n" % (e,))
            return 1

    def reduce_raw(self, buf, count, dtype):
        """In-place sum all-reduce of `count` elements at address `buf` (host, or device when the
        RED_DEVICE bit is set)."""
        import torch
        on_dev = bool(dtype & RED_DEVICE)
        base = dtype & 0xff
        npdt = _NP[base]
        if count == 0:
            return
        if on_dev:
            typestr = {RED_U64: "<u8", RED_F32: "<f4", RED_I32: "<i4"}[base]
            t = torch.as_tensor(_DevArray(buf, count, typestr), device=self.device)
            if base == RED_U64:
                t = t.view(torch.int64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            torch.cuda.synchronize(self.device)
        else:
            a = np.ctypeslib.as_array(C.cast(buf, C.POINTER(np.ctypeslib.as_ctypes_type(npdt))),
                                      shape=(count,))
            # uint64 counters stay far below 2^63: reduce them as int64 (gloo/NCCL have no uint64 sum)
            t = torch.from_numpy(a.view(np.int64) if base == RED_U64 else a)
            if self.backend == "nccl":
                g = t.to(self.device)
                self.dist.all_reduce(g, op=self.dist.ReduceOp.SUM, group=self.group)
                t.copy_(g.cpu())
            else:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        self.n_reductions += 1

    @property
    def reducer(self):
        return self._cb

    def all_reduce_sum(self, value):
        """Sum of a Python int over the ranks."""
        a = np.array([int(value)], dtype=np.int64)
        self.reduce_raw(a.ctypes.data, 1, RED_U64)
        return int(a[0])

    def barrier(self):
        self.dist.barrier(group=self.group)

    def sync(self):
        """Wait for the collectives enqueued on torch's streams: libannb works on its own stream."""
        if self.device is not None:
            import torch
            torch.cuda.synchronize(self.device)

    # -- variable-length all-gather ----------------------------------------------------------
    def all_gather_var(self, columns):
        """columns: list of 1-D tensors of equal length n_r on this rank (different per rank).
        Returns, per column, the list [tensor from rank 0, ..., tensor from rank world-1]."""
        import torch
        n = int(columns[0].shape[0]) if columns else 0
        counts = torch.zeros(self.world, dtype=torch.int64, device=columns[0].device if columns else None)
        counts[self.rank] = n
        self.dist.all_reduce(counts, op=self.dist.ReduceOp.SUM, group=self.group)
        counts = counts.cpu().tolist()
        cap = max(max(counts), 1)
        out = []
        for col in columns:
            pad = torch.zeros(cap, dtype=col.dtype, device=col.device)
            pad[:n] = col
            parts = [torch.empty(cap, dtype=col.dtype, device=col.device) for _ in range(self.world)]
            self.dist.all_gather(parts, pad, group=self.group)
            out.append([p[:c] for p, c in zip(parts, counts)])
            self.bytes_gathered += cap * col.element_size() * (self.world - 1)
        return out


def exchange(comm, local_columns, import_fn, chunk=1 << 24):
    """All-gather this rank's result columns and hand every OTHER rank's rows to import_fn(cols).
    Works through the rows in chunks so that the staging buffers stay small (at N=1M a refine round
    moves ~10^8 rows per rank).  Returns the number of rows imported."""
    import torch
    n = int(local_columns[0].shape[0]) if local_columns else 0
    counts = torch.zeros(comm.world, dtype=torch.int64, device=local_columns[0].device)
    counts[comm.rank] = n
    comm.dist.all_reduce(counts, op=comm.dist.ReduceOp.SUM, group=comm.group)
    n_max = int(counts.max().item())
    imported = 0
    for lo in range(0, n_max, chunk):
        part = [c[min(lo, n):min(lo + chunk, n)] for c in local_columns]
        gathered = comm.all_gather_var(part)
        comm.sync()  # the gathered columns are read next by kernels on the index's own stream
        for r in range(comm.world):
            if r == comm.rank:
                continue
            cols = [g[r] for g in gathered]
            if cols and cols[0].shape[0]:
                import_fn(cols)
                imported += int(cols[0].shape[0])
        del gathered
    return imported


def tiles_of_rank(n_tiles, rank, world):
    """The tile shard of a rank: t with t % world == rank (what the sweeps in libannb use)."""
    return range(rank, n_tiles, world)


class IndexExchange:
    """Device-to-device export / all-gather / import of an index's per-rank results."""

    def __init__(self, index, comm):
        import torch
        self.ix = index
        self.comm = comm
        self.torch = torch
        self.dev = comm.device

    def _bufs(self, n, dtypes):
        return [self.torch.empty(max(n, 1), dtype=dt, device=self.dev) for dt in dtypes]

    def refined(self):
        """After refine_selected(): share (i, j, d) of the pairs this rank evaluated."""
        t = self.torch
        n = self.ix.export_refined(None, None, None, 0)
        bi, bj, bd = self._bufs(n, [t.int32, t.int32, t.float32])
        if n:
            self.ix.export_refined(bi.data_ptr(), bj.data_ptr(), bd.data_ptr(), n)
        t.cuda.synchronize(self.dev)

        def imp(cols):
            i, j, d = [c.contiguous() for c in cols]
            self.ix.import_dev(1, i.data_ptr(), j.data_ptr(), d.data_ptr(), None, i.shape[0])

        return exchange(self.comm, [bi[:n], bj[:n], bd[:n]], imp)

    def tightened(self):
        """After update_bounds(): share (i, j, lb, ub) of the pairs this rank tightened."""
        t = self.torch
        n = self.ix.export_tightened(None, None, None, None, 0)
        bi, bj, ba, bb = self._bufs(n, [t.int32, t.int32, t.float32, t.float32])
        if n:
            self.ix.export_tightened(bi.data_ptr(), bj.data_ptr(), ba.data_ptr(), bb.data_ptr(), n)
        t.cuda.synchronize(self.dev)

        def imp(cols):
            i, j, a, b = [c.contiguous() for c in cols]
            self.ix.import_dev(2, i.data_ptr(), j.data_ptr(), a.data_ptr(), b.data_ptr(), i.shape[0])

        return exchange(self.comm, [bi[:n], bj[:n], ba[:n], bb[:n]], imp)
