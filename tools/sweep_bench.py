"""Frequency-sharded sweep of config 4 (100x100x60 linear mesh, 32 frequencies 1e-3..1e3 Hz) on N GPUs.

    python tools/sweep_bench.py [--scale S]                                   # 1 GPU, all 32 frequencies
    torchrun --nproc-per-node N tools/sweep_bench.py [--scale S] [--gather]   # rank r takes frequencies r+1, r+1+N, ...

Every rank holds the mesh, the pattern and the K_e/M_e cache of the unstretched elements; its first frequency is a
cold assembly, the later ones recompute only the GPML layers and the RHS (the device-resident API, inputs in HBM).
No data-path collective; with --gather the finished value arrays go to rank 0 with NCCL send/recv (what a
centralised ZMUMPS would need).  Prints one JSON line on rank 0; time = max over ranks of the whole shard."""
import argparse, json, os, sys
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION": os.environ["NCCL_DEBUG"] = "WARN"   # keep the banner off stdout
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movfem_b200 import mesh, host, abi
from movfem_b200.sharding import frequency_shard

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--gather", action="store_true")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    _fd = os.dup(1); os.dup2(2, 1)          # NCCL's banner goes to stderr, stdout carries the JSON line
    dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(_fd, 1); os.close(_fd)
m = mesh.config(4, scale=args.scale)
nf = len(m.freqs)
mine = frequency_shard(nf, rank, world)
asm = host.Assembly(m, device=local)
stream = torch.cuda.Stream(); asm.set_stream(stream.cuda_stream)
# per-frequency g_sigma (the host's update_sigma history) resident in HBM before the timed region
sig = {f: torch.from_numpy(m.sigma_for(f).view(np.float64).reshape(-1)).to(dev) for f in mine}
best, per_freq = None, None
for rep in range(args.repeat):
    asm.reset_cache()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    stats = []
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for f in mine:
            asm.assemble_device(f, m.omega(f), sig[f].data_ptr(), abi.MODE_T2)
            p = asm.device_result()
            stats.append(asm.stats())
        ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if best is None or ms < best: best, per_freq = ms, stats
t = torch.tensor([best], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
gather_ms = None
if args.gather and world > 1:
    nz = p[4]
    mine_t = torch.empty(2 * nz, dtype=torch.float64, device=dev)
    import ctypes
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaMemcpy(mine_t.data_ptr(), p[2], nz * 16, 3)
    def exchange():     # one frequency's values from every rank to rank 0
        if rank == 0:
            for r in range(1, world):
                n = torch.zeros(1, dtype=torch.int64, device=dev); dist.recv(n, src=r)
                buf = torch.empty(int(n.item()), dtype=torch.float64, device=dev); dist.recv(buf, src=r)
        else:
            dist.send(torch.tensor([mine_t.numel()], dtype=torch.int64, device=dev), dst=0); dist.send(mine_t, dst=0)
    exchange(); torch.cuda.synchronize(); dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(); exchange(); g1.record(); torch.cuda.synchronize()
    tg = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64); dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    gather_ms = float(tg.item())
if rank == 0:
    ms = float(t.item())
    cold, warm = per_freq[0], per_freq[1:] or per_freq
    avg = lambda k: float(np.mean([s[k] for s in warm]))
    print(json.dumps({"workload": m.name, "scale": args.scale, "n_gpus": world, "frequencies": nf, "elements": m.ne, "nne": asm.nne,
                      "nnz_per_frequency": int(p[4]), "ms_sweep_max_over_ranks": ms,
                      "frequencies_per_s": nf / (ms * 1e-3), "elements_per_s": nf * m.ne / (ms * 1e-3), "nnz_per_s": nf * int(p[4]) / (ms * 1e-3),
                      "rank0_cold_frequency_ms": cold["ms_total"],
                      "rank0_cached_frequency_ms": {k: avg(k) for k in ("ms_total", "ms_node", "ms_element", "ms_geometry", "ms_contract", "ms_gather", "ms_finalize")},
                      "gather_one_frequency_per_rank_to_rank0_ms": gather_ms,
                      "sharding": "frequency round-robin, replicated mesh/pattern/K-M cache, no data-path collective"}))
asm.close()
if world > 1: dist.destroy_process_group()
