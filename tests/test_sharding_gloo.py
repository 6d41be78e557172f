"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU paths.

Frequencies are dealt round-robin with no data-path collective; the only exchange is the hand-off of
finished per-frequency results to the MUMPS host rank.  Here the per-rank 'assembly' is the CPU oracle
(this is a test of the plumbing, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from movfem_b200 import mesh
    from movfem_b200.sharding import frequency_shard, gather_to_root
    from oracle.oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = mesh.build_model("t", 4, 4, 8, 1000., 1000., 1000., 1, 1, 1, dirichlet=1, freqs=(0.1, 1.0, 10.0, 100.0, 1000.0))
    o = Oracle(m)
    mine = frequency_shard(len(m.freqs), rank, world)
    res = {}
    for ifreq in mine:
        r = o.assemble(m.omega(ifreq), m.sigma_for(ifreq), nthreads=1)
        res[ifreq] = torch.from_numpy(r["a"].view(np.float64).copy())
    gathered = gather_to_root(res, len(m.freqs), root=0)
    if rank == 0:
        ok = sorted(gathered) == list(range(1, len(m.freqs) + 1))
        for ifreq in gathered:      # root recomputes everything serially and compares bit for bit
            r = o.assemble(m.omega(ifreq), m.sigma_for(ifreq), nthreads=1)
            ok &= bool(np.array_equal(gathered[ifreq].numpy().view(np.complex128), r["a"]))
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_frequency_sharding_world2():
    from movfem_b200.sharding import frequency_shard
    assert frequency_shard(5, 0, 2) == [1, 3, 5] and frequency_shard(5, 1, 2) == [2, 4]
    assert sorted(sum((frequency_shard(32, r, 8) for r in range(8)), [])) == list(range(1, 33))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    ok = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in ps)


def test_slab_partition_covers_mesh():
    from movfem_b200.sharding import slab_partition
    for n, w in ((400, 8), (58, 4), (5, 2), (7, 8)):
        parts = [slab_partition(n, r, w) for r in range(w)]
        cover = [i for lo, hi in parts for i in range(lo, hi + 1) if hi >= lo]
        assert cover == list(range(1, n + 1))
