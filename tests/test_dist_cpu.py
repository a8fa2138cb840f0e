"""Host-side logic of the sharded fit() (annchor_b200/dist.py) on CPU: world_size 2, gloo.

The device kernels cannot run here; what is covered is everything that runs on the host when the
index is sharded -- the sum all-reduce callback libannb calls through annb_index_set_reducer
(invoked through ctypes exactly as the C side does), the variable-length all-gather, the
export / import exchange protocol (against a fake index holding numpy columns) and the tile
partition."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from annchor_b200 import dist as D
        comm = D.Comm()
        assert (comm.rank, comm.world) == (rank, world)
        # ---- reducer callback, called the way libannb calls it (C function pointer, host buffers)
        fn = C.cast(comm.reducer, C.c_void_p).value
        cb = D.REDUCE_FN(fn)
        u = np.arange(5, dtype=np.uint64) * (rank + 1)
        f = np.zeros(8, dtype=np.float32)
        f[rank * 4:(rank + 1) * 4] = np.arange(4) + 10 * rank + 0.5  # each rank owns one slice
        i = np.full(3, rank + 1, dtype=np.int32)
        assert cb(None, u.ctypes.data, 5, D.RED_U64) == 0
        assert cb(None, f.ctypes.data, 8, D.RED_F32) == 0
        assert cb(None, i.ctypes.data, 3, D.RED_I32) == 0
        assert u.tolist() == [0, 3, 6, 9, 12]
        assert f.tolist() == [0.5, 1.5, 2.5, 3.5, 10.5, 11.5, 12.5, 13.5]
        assert i.tolist() == [3, 3, 3]
        assert cb(None, None, 0, D.RED_U64) == 0  # empty reductions are legal
        assert comm.all_reduce_sum(7 + rank) == 15
        # ---- variable-length all-gather (rank r contributes r + 2 rows)
        n = rank + 2
        ci = torch.arange(n, dtype=torch.int32) + 100 * rank
        cd = torch.arange(n, dtype=torch.float32) * 0.5 + rank
        gi, gd = comm.all_gather_var([ci, cd])
        assert [g.tolist() for g in gi] == [[0, 1], [100, 101, 102]]
        assert [g.tolist() for g in gd] == [[0.0, 0.5], [1.0, 1.5, 2.0]]
        # a rank with nothing to share
        e = comm.all_gather_var([torch.zeros(0 if rank == 0 else 3, dtype=torch.int32)])[0]
        assert [x.shape[0] for x in e] == [0, 3]

        # ---- exchange protocol: after it every rank's store holds the union
        class FakeIndex:
            def __init__(self):
                self.store = {}

            def add(self, i, j, d):
                for a, b, c in zip(i.tolist(), j.tolist(), d.tolist()):
                    self.store[(a, b)] = c

        ix = FakeIndex()
        # rank r "evaluated" the pairs of its tiles: pair p belongs to tile p % world
        pairs = [(p, p + 1, float(p) * 0.25) for p in range(11) if p % world == rank]
        li = torch.tensor([p[0] for p in pairs], dtype=torch.int32)
        lj = torch.tensor([p[1] for p in pairs], dtype=torch.int32)
        ld = torch.tensor([p[2] for p in pairs], dtype=torch.float32)
        ix.add(li, lj, ld)
        got = D.exchange(comm, [li, lj, ld], lambda cols: ix.add(*cols))
        assert got == 11 - len(pairs)
        assert ix.store == {(p, p + 1): p * 0.25 for p in range(11)}
        # chunked staging (what bounds the buffers at N=1M): same union, ragged last chunk, and a
        # rank that runs out of rows before the other keeps taking part in the collectives
        ix2 = FakeIndex()
        ix2.add(li, lj, ld)
        calls = []
        got2 = D.exchange(comm, [li, lj, ld], lambda cols: (calls.append(int(cols[0].shape[0])), ix2.add(*cols)),
                          chunk=2)
        assert got2 == got and ix2.store == ix.store
        assert max(calls) <= 2 and sum(calls) == got
        # ---- tile partition: disjoint cover
        T = 37
        mine = set(D.tiles_of_rank(T, rank, world))
        counts = torch.zeros(T, dtype=torch.int64)
        counts[list(mine)] = 1
        dist.all_reduce(counts)
        assert counts.tolist() == [1] * T
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_dist_host_logic_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_reducer_requires_process_group():
    from annchor_b200 import dist as D
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("a process group is already initialised in this interpreter")
    with pytest.raises(RuntimeError):
        D.Comm()
