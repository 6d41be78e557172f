"""Host-link microbenchmark (explains the end-to-end scaling of bench.py at N > 1): pinned D2H bandwidth of one rank alone and of all
ranks at once, with the default placement and with the process pinned to the CPUs of its GPU's NUMA node before allocating.

    torchrun --nproc-per-node N tools/d2h_scaling.py
"""
import json, os, sys, time
import torch, torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def numa_cpus(gpu):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read())
        if node < 0:
            return None, node
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus, node
    except Exception as exc:
        return None, repr(exc)

def bw(nbytes, reps, bind):
    if bind:
        cpus, node = numa_cpus(local)
        if cpus:
            os.sched_setaffinity(0, cpus)
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.zero_()                      # first touch on the (possibly bound) CPUs
    out = {}
    for mode in ("alone", "all"):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        if mode == "alone" and rank != 0:
            if world > 1: dist.barrier()
            continue
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        out[mode] = nbytes * reps / (time.perf_counter() - t0) * 1e-9
        if mode == "alone" and world > 1: dist.barrier()
    return out

res = {"default": bw(2 << 30, 5, False), "numa_bound": bw(2 << 30, 5, True), "numa": numa_cpus(local)[1], "ncpu": os.cpu_count()}
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, res)
else:
    allr = [res]
if rank == 0:
    agg = {k: sum(r[k].get("all", 0) for r in allr) for k in ("default", "numa_bound")}
    print(json.dumps({"world": world, "rank0_alone_GBs": {k: allr[0][k].get("alone") for k in ("default", "numa_bound")}, "aggregate_all_ranks_GBs": agg,
                      "per_rank": allr}))
if world > 1:
    dist.destroy_process_group()
