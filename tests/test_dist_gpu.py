"""Sharded fit() on 2 GPUs == fit() on 1 GPU ("multi-GPU = same result as
1 GPU").  One process per GPU, NCCL.  Skipped on boxes with a single GPU."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _blobs(n, d, seed):
    """The bench generator: 100-centre Gaussian blobs."""
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(100, d)) * (30.0 / np.sqrt(d))
    return (c[rng.integers(0, 100, size=n)] + rng.normal(size=(n, d))).astype(np.float32)


def _fit(X, comm, device, **kw):
    import annchor_b200 as ab
    from annchor_b200.annchor import Annchor
    ctx = ab.default_context(device)
    a = Annchor(X, "euclidean", ctx=ctx, comm=comm, **kw).fit()
    return a


def _worker(rank, world, port, out_dir, n, kw):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from annchor_b200.dist import Comm
        comm = Comm(device=torch.device("cuda", rank))
        a = _fit(_blobs(n, 128, 42), comm, rank, **kw)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), idx=a.neighbor_graph[0], d=a.neighbor_graph[1],
                 A=a.A, evals=a.evals, n_tight=a.n_tightened, red=comm.n_reductions)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,kw", [
    (3000, dict(n_anchors=20, n_neighbors=15, n_samples=2000, p_work=0.1)),
    (20000, dict(n_anchors=30, n_neighbors=15, n_samples=5000, p_work=0.01)),  # pool-mode sampler
])
def test_two_gpu_fit_equals_one_gpu(tmp_path, n, kw):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), n, kw), nprocs=world, join=True)
    r0, r1 = [np.load(os.path.join(tmp_path, "r%d.npz" % r)) for r in range(world)]
    one = _fit(_blobs(n, 128, 42), None, 0, **kw)
    for r in (r0, r1):
        assert np.array_equal(r["A"], one.A)
        assert int(r["evals"]) == one.evals
        assert np.array_equal(r["idx"], one.neighbor_graph[0])
        assert np.array_equal(r["d"], one.neighbor_graph[1])
        assert int(r["red"]) > 0
    assert int(r0["n_tight"]) == one.n_tightened
