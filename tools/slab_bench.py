"""Slab-sharded assembly of config 5 (large linear mesh with topography) on N GPUs.

    python tools/slab_bench.py [--scale S] [--steps K]                      # 1 GPU, whole mesh
    torchrun --nproc-per-node N tools/slab_bench.py [--scale S] [--gather]  # N x-slabs

Each rank owns a range of ie (contiguous rows), computes its +x halo itself (no data-path collective) and, with
--gather, hands its value slice to rank 0 with NCCL send/recv.  Prints one JSON line on rank 0."""
import argparse, json, os, sys, time, copy
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION": os.environ["NCCL_DEBUG"] = "WARN"   # keep the banner off stdout
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movfem_b200 import mesh, host, abi
from movfem_b200.sharding import slab_partition

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.5); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--gather", action="store_true"); ap.add_argument("--config", type=int, default=5)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    _fd = os.dup(1); os.dup2(2, 1)          # NCCL's banner goes to stderr, stdout carries the JSON line
    dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(_fd, 1); os.close(_fd)
m = mesh.config(args.config, scale=args.scale)
if world > 1:
    m = copy.copy(m); m.ie_lo, m.ie_hi = slab_partition(m.g_nx - 1, rank, world)
t0 = time.perf_counter(); asm = host.Assembly(m, device=local); t_create = time.perf_counter() - t0
stream = torch.cuda.Stream(); asm.set_stream(stream.cuda_stream)
om, sg = m.omega(1), m.sigma_for(1)
d_sigma = torch.from_numpy(sg.view(np.float64).reshape(-1)).to(dev)
ms = []
for it in range(1 + args.steps):
    asm.reset_cache()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1: dist.barrier()
        e0.record(stream); asm.assemble_device(1, om, d_sigma.data_ptr(), abi.MODE_T2); p = asm.device_result(); e1.record(stream)
    torch.cuda.synchronize()
    if it: ms.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.mean(ms))], device=dev, dtype=torch.float64)
nz = torch.tensor([p[4]], device=dev, dtype=torch.int64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(nz)
gather_ms = None
if args.gather and world > 1:
    # values only (IRN/JCN are static): rank 0 receives every slice at its row-ordered offset
    import ctypes
    n_local = p[4]
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    mine = torch.empty(2 * n_local, dtype=torch.float64, device=dev)
    rt.cudaMemcpy(mine.data_ptr(), p[2], n_local * 16, 3)
    full = torch.empty(2 * sum(counts), dtype=torch.float64, device=dev) if rank == 0 else None
    def exchange():
        if rank == 0:
            full[: 2 * counts[0]] = mine
            off = 2 * counts[0]
            for r in range(1, world):
                dist.recv(full[off: off + 2 * counts[r]], src=r); off += 2 * counts[r]
        else:
            dist.send(mine, dst=0)
    exchange()                                   # warm-up: NCCL builds its peer connections lazily
    torch.cuda.synchronize(); dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(); exchange(); g1.record(); torch.cuda.synchronize()
    tg = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
    dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    gather_ms = float(tg.item())
if rank == 0:
    ne_total = (mesh.config(args.config, scale=args.scale)).ne if world > 1 else m.ne
    print(json.dumps({"workload": m.name, "scale": args.scale, "n_gpus": world, "elements": ne_total, "nne": asm.nne, "nnz_delivered": int(nz.item()),
                      "ms_per_assembly_max_over_ranks": float(t.item()), "elements_per_s": ne_total / (float(t.item()) * 1e-3),
                      "nnz_per_s": int(nz.item()) / (float(t.item()) * 1e-3), "create_s_rank0": t_create, "gather_values_to_rank0_ms": gather_ms,
                      "sharding": "x-slab, one-element +x halo computed locally, no data-path collective" if world > 1 else "none",
                      "stats_rank0": asm.stats()}))
asm.close()
if world > 1: dist.destroy_process_group()
