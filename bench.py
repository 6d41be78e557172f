#!/usr/bin/env python
"""bench.py -- throughput of the MoVFEM_3DMT element assembly hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (default N=1)
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm (oracle port)

Headline workload at every N: BASELINE.json configs[4], the 400x400x200 linear-element mesh with topography
(32 M elements, 1.63 G delivered entries, GPML Fang) -- the largest configuration that fits one B200 (about 115 of
180 GB).  A "step" is one COLD assembly of one frequency: node fields, every element's K_e/M_e/b_e, the
deterministic gather, A = K + i*w32*M, float32 round trip, zero strip, RHS (the K/M cache is reset before every
step, nothing is skipped).  N > 1: the mesh is split into N x-slabs (ranges of ie own contiguous rows, SURVEY 8e;
the +x halo layer is recomputed, no data-path collective) -> STRONG scaling of the same 32 M elements.

value       device-timed (CUDA events on the launching stream), inputs resident in HBM, max over ranks
e2e         the same metric through the C-ABI host call movfem_assemble with pinned HOST buffers: H2D of g_sigma and
            D2H of irn/jcn/a/rhs inside the timed region; at N > 1 every rank delivers its row slice to host memory
            (the hand-off to the MUMPS host rank), max over ranks
per_config  (N = 1) configs[0..3] measured the same way, each with both roofline fractions of the step and of the
            kernel that does the algorithmic flops, its e2e and a CPU baseline
sweep       configs[3]: the 32-frequency sweep on the 100x100x60 mesh, frequencies dealt round-robin to the N GPUs
"""
import argparse
import copy
import ctypes
import glob
import json
import os
import subprocess
import sys
import threading
import time

# NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION and latches the level at its first call, which can
# happen while torch.distributed is imported; stdout carries exactly one JSON line, so lower the level before that
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from movfem_b200 import abi, mesh  # noqa: E402
from movfem_b200.sharding import frequency_shard, slab_partition  # noqa: E402

HEADLINE = 5     # BASELINE.json configs[4]
FLOPS_PER_ELEMENT = {12: 10944, 36: 250776, 54: 533628}   # 2*ngp*(18*me + 3*me*(me+1))
BYTES_PER_ELEMENT = {12: 1071, 36: 9900, 54: 22300}        # algorithmic bytes of a cold assembly, per element
BYTES_PER_NNZ_UPDATE = 32                                  # read K 8 + M 8, write A 16


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from this round's ncu --set full captures
    (profiles/r02_ncu_*.json, written by tools/summarize_ncu.py): {workload: {kernel: bytes}}"""
    out = {}
    for f in sorted(glob.glob(os.path.join(HERE, "profiles", "r02_ncu_*.json"))):
        try:
            d = json.load(open(f))
            out.setdefault(d.get("workload", os.path.basename(f)), {}).update(d.get("dram_bytes_per_launch", {}))
        except Exception:
            pass
    return out


def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    try:
        with open(os.path.join(HERE, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


PHASES = ("ms_node", "ms_element", "ms_geometry", "ms_contract", "ms_fused", "ms_exact", "ms_gather", "ms_finalize")


def step_roofs(me, ne, nz, phases, ms_step, fp64_peak, hbm_peak, traffic=None):
    """Both roofline fractions of the whole step and of its kernels, from ALGORITHMIC work: flops = the
    B^T D B count of the two contractions (Jacobian / basis / RHS / re-evaluation flops are not counted), bytes = what
    a cold assembly must read and write once.  Pure function of its arguments (tests/test_abi_host.py)."""
    flops, nbytes = FLOPS_PER_ELEMENT[me] * ne, BYTES_PER_ELEMENT[me] * ne
    sec = ms_step * 1e-3
    out = {"step": {"ms": ms_step, "tflops": flops / sec * 1e-12, "frac_fp64": flops / sec * 1e-12 / fp64_peak if fp64_peak else None,
                    "gbs": nbytes / sec * 1e-9, "frac_hbm": nbytes / sec * 1e-9 / hbm_peak,
                    "frac_fp64_of_nominal_37.2": flops / sec * 1e-12 / 37.2}}
    # the kernel(s) that execute the algorithmic flops: contract_kernel (+ geometry_kernel feeding it), or fused12_kernel
    ms_flop = (phases.get("ms_contract") or 0.0) + (phases.get("ms_fused") or 0.0)
    if ms_flop > 0:
        out["flop_kernels"] = {"kernels": "fused12_kernel (+ contract_kernel on the GPML layers)" if phases.get("ms_fused") else "contract_kernel",
                               "ms": ms_flop, "tflops": flops / (ms_flop * 1e-3) * 1e-12,
                               "frac_fp64": flops / (ms_flop * 1e-3) * 1e-12 / fp64_peak if fp64_peak else None}
    ms_ga = phases.get("ms_gather") or 0.0
    if ms_ga > 0:
        gb = BYTES_PER_NNZ_UPDATE * nz / (ms_ga * 1e-3) * 1e-9
        out["gather"] = {"kernels": "gather_finalize_kernel (+ rhs_kernel)", "ms": ms_ga, "gbs": gb, "frac_hbm": gb / hbm_peak, "bytes_per_nnz": BYTES_PER_NNZ_UPDATE}
    shares = {k[3:]: v / ms_step for k, v in phases.items() if k in ("ms_node", "ms_geometry", "ms_contract", "ms_fused", "ms_exact", "ms_gather", "ms_finalize") and v}
    out["share_of_step"] = shares
    if traffic:
        out["ncu_dram_bytes_per_launch"] = traffic
    return out


def dominant_roofline(ms_flop_kernels, ms_gather):
    """Which of the two roofline objects the top-level `roofline` key repeats: the kernels with the larger share of the step."""
    return "roofline_hbm" if (ms_gather or 0.0) >= (ms_flop_kernels or 0.0) else "roofline_fp64"


def pinned(n, dt):
    import torch
    return torch.empty(n, dtype=dt, pin_memory=True)


def measure_cold(model, local_rank, steps, warmup, barrier, e2e_steps, want_pageable=False):
    """Cold assemblies of one frequency of `model` (a whole mesh or this rank's x-slab): device-timed and end to end."""
    import torch
    from movfem_b200 import host
    dev = torch.device("cuda", local_rank)
    omega, sigma_np = model.omega(1), model.sigma_for(1)
    t0 = time.perf_counter()
    asm = host.Assembly(model, device=local_rank)
    create_s = time.perf_counter() - t0
    stream = torch.cuda.Stream(device=dev)
    asm.set_stream(stream.cuda_stream)
    sigma_dev = torch.from_numpy(sigma_np.view(np.float64).reshape(-1)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    phase = {k: [] for k in PHASES}
    launches, nz, nflag = 0, 0, 0
    with torch.cuda.stream(stream):
        for it in range(warmup + steps):
            if it == warmup:
                torch.cuda.synchronize(dev); barrier(); torch.cuda.synchronize(dev)
            flush.zero_()                                   # evict L2 between iterations
            asm.reset_cache()                               # every step is a cold, full assembly
            if it >= warmup:
                ev[it - warmup][0].record(stream)
            asm.assemble_device(1, omega, sigma_dev.data_ptr(), abi.MODE_T2)
            _, _, _, _, nz = asm.device_result()            # completes the step (zero strip if needed)
            if it >= warmup:
                ev[it - warmup][1].record(stream)
                st = asm.stats()
                for k in phase:
                    phase[k].append(st[k])
                launches += int(st["launches"]) + 1         # + the L2 flush fill
                nflag = int(st["nflagged"])
        torch.cuda.synchronize(dev); barrier(); torch.cuda.synchronize(dev)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    del flush, sigma_dev
    out = {"ms_per_step": dev_ms, "phases_ms": {k: float(np.mean(v)) for k, v in phase.items()}, "nz": int(nz), "nne": asm.nne,
           "nz_upper": asm.nz_upper, "launches": launches, "create_s": create_s, "elements_computed": model.ne, "nflagged": nflag}

    # ---- end to end: C-ABI host call, host buffers; H2D of g_sigma and D2H of irn/jcn/a/rhs inside the timed region ----
    if e2e_steps > 0:
        h_sigma = pinned(sigma_np.size * 2, torch.float64); h_sigma.numpy()[:] = sigma_np.view(np.float64).reshape(-1)
        h_irn, h_jcn = pinned(asm.nz_upper, torch.int32), pinned(asm.nz_upper, torch.int32)
        h_a, h_rhs = pinned(asm.nz_upper * 2, torch.float64), pinned(asm.nne * 4, torch.float64)
        sig_c = h_sigma.numpy().view(np.complex128)
        a_c, rhs_c = h_a.numpy().view(np.complex128), h_rhs.numpy().view(np.complex128)

        def loop(mode, irn, jcn, a, rhs, n):
            ts, nz_e = [], 0
            for it in range(1 + n):
                if it == 1:
                    torch.cuda.synchronize(dev); barrier()
                asm.reset_cache()
                t0 = time.perf_counter()
                _, _, _, _, nz_e = asm.global_vfem(1, omega, sig_c, mode=mode, irn=irn, jcn=jcn, a=a, rhs=rhs)
                if it >= 1:
                    ts.append(time.perf_counter() - t0)
            return float(np.mean(ts)), nz_e, asm.stats()
        s_full, nz_e, st_full = loop(abi.MODE_T2, h_irn.numpy(), h_jcn.numpy(), a_c, rhs_c, e2e_steps)
        s_keep, _, _ = loop(abi.MODE_T2 | abi.MODE_KEEP_PATTERN, h_irn.numpy(), h_jcn.numpy(), a_c, rhs_c, e2e_steps)
        out["e2e"] = {"s_per_step": s_full, "h2d_bytes_per_step": int(sigma_np.size * 16), "d2h_bytes_per_step": int(nz_e * 24 + asm.nrows * 32),
                      "ms_h2d": st_full["ms_h2d"], "ms_d2h": st_full["ms_d2h"], "host_memory": "pinned (cudaHostAlloc)",
                      "keep_pattern_variant": {"s_per_step": s_keep, "d2h_bytes_per_step": int(nz_e * 16 + asm.nrows * 32),
                                               "note": "MOVFEM_MODE_KEEP_PATTERN: the static irn/jcn are not re-sent (what the Fortran shim does after the first frequency)"}}
        if want_pageable:
            # what a Fortran `allocate`d array is: pageable memory (the library registers it with cudaHostRegister, cached per pointer)
            p_irn, p_jcn = np.empty(asm.nz_upper, np.int32), np.empty(asm.nz_upper, np.int32)
            p_a, p_rhs = np.empty(asm.nz_upper, np.complex128), np.empty(2 * asm.nne, np.complex128)
            s_page, _, _ = loop(abi.MODE_T2, p_irn, p_jcn, p_a, p_rhs, max(2, e2e_steps // 2))
            out["e2e"]["pageable_variant"] = {"s_per_step": s_page, "note": "caller arrays from numpy.empty (pageable), as a Fortran allocate gives"}
            del p_irn, p_jcn, p_a, p_rhs
        del h_sigma, h_irn, h_jcn, h_a, h_rhs
    asm.close()
    torch.cuda.empty_cache()
    return out


def measure_sweep(model, local_rank, rank, world, barrier, repeat=2):
    """configs[3]: every frequency of the list assigned to this rank (round-robin), device-resident; first one cold."""
    import torch
    from movfem_b200 import host
    dev = torch.device("cuda", local_rank)
    nf = len(model.freqs)
    mine = frequency_shard(nf, rank, world)
    asm = host.Assembly(model, device=local_rank)
    stream = torch.cuda.Stream(device=dev)
    asm.set_stream(stream.cuda_stream)
    sig = {f: torch.from_numpy(model.sigma_for(f).view(np.float64).reshape(-1)).to(dev) for f in mine}
    best, per_freq, nz = None, None, 0
    for _ in range(repeat):
        asm.reset_cache()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev); barrier(); torch.cuda.synchronize(dev)
        stats = []
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for f in mine:
                asm.assemble_device(f, model.omega(f), sig[f].data_ptr(), abi.MODE_T2)
                nz = asm.device_result()[4]
                stats.append(asm.stats())
            ev1.record(stream)
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1)
        if best is None or ms < best:
            best, per_freq = ms, stats
    nne, nzu = asm.nne, asm.nz_upper
    asm.close()
    del sig
    torch.cuda.empty_cache()
    cold, warm = per_freq[0], per_freq[1:] or per_freq
    return {"ms_shard": best, "frequencies": nf, "mine": len(mine), "nz": int(nz), "nne": nne, "nz_upper": nzu,
            "cold_frequency_ms": cold["ms_total"],
            "cached_frequency_ms": {k: float(np.mean([s[k] for s in warm])) for k in ("ms_total", "ms_node", "ms_geometry", "ms_exact", "ms_gather")}}


def cpu_port_sample(model, n_sample, note):
    """The oracle in `faithful` mode (reference loop structure: one alocal per pair, every redundant Jacobian rebuild),
    single thread -- the reference assembles sequentially on rank 0 (MoVFEM_3DMT.f90:63) -- on a bounded run of elements."""
    from oracle.oracle import Oracle
    o = Oracle(model)
    omega, sigma = model.omega(1), model.sigma_for(1)
    n_sample = min(n_sample, model.ne)
    lo = 1 + (model.ne // 2 // n_sample) * n_sample
    r = o.assemble(omega, sigma, faithful=True, nthreads=1, want_t1=False, want_t2=False, ide_range=(lo, lo + n_sample - 1))
    nt = os.cpu_count() or 1
    n_all = min(model.ne, max(n_sample, 4000))
    lo2 = 1 + (model.ne // 2 // n_all) * n_all
    r2 = o.assemble(omega, sigma, faithful=False, nthreads=nt, want_t1=False, want_t2=False, ide_range=(lo2, lo2 + n_all - 1))
    return {"value": n_sample / r["seconds"], "unit": "elements/s", "cores": 1, "kind": "port",
            "sample": f"{n_sample} consecutive elements (ide {lo}..{lo + n_sample - 1}) of {note}, oracle faithful mode (reference loop structure), "
                      f"g++ -O1 -ffp-contract=off, {r['seconds']:.1f} s",
            "host_cores_available": nt,
            "all_cores_port": {"value": n_all / r2["seconds"], "unit": "elements/s", "cores": nt, "sample_elements": n_all,
                               "what": "oracle with memoised Jacobians, OpenMP over elements (an optimised CPU port, not the reference's sequential loop)"}}


def cpu_sample_model(n):
    """The mesh the CPU sample runs on: the config itself, except config 5 whose 3.2 G structural entries the reference (and
    its port) cannot index -- there the 40x40x20 sub-mesh with the same spacing / topography."""
    if n == 5:
        return mesh.config5_submesh(), "the 40x40x20 sub-mesh of config 5 (same spacing, 300 m topography; the full mesh overflows the reference's 32-bit nzindx, SURVEY Q16)"
    m = mesh.config(n)
    return m, m.name


REF_SAMPLE = {1: 20000, 2: 300, 3: 150, 4: 20000, 5: 20000}   # elements per CPU sample: roughly 3-15 s of one core each


def run_reference(args):
    """The reference's own algorithm on the host cores: the oracle port in `faithful` mode, single thread because the
    reference assembles sequentially on rank 0 (MoVFEM_3DMT.f90:63).  No Fortran compiler exists in this image, so the
    reference itself cannot be built (oracle/_ref is impossible).  Workload: the headline's (config 5), on the
    bounded sample cpu_sample_model() describes; each step = one pass over that sample."""
    rank, _, world = rank_env()
    if rank != 0:
        return
    from oracle.oracle import Oracle
    model, note = cpu_sample_model(HEADLINE)
    o = Oracle(model)
    omega, sigma = model.omega(1), model.sigma_for(1)
    n_sample = min(args.ref_elements or REF_SAMPLE[HEADLINE], model.ne)
    lo = 1 + (model.ne // 2 // n_sample) * n_sample
    times = []
    for it in range(args.warmup + args.steps):
        r = o.assemble(omega, sigma, faithful=True, nthreads=1, want_t1=False, want_t2=False, ide_range=(lo, lo + n_sample - 1))
        if it >= args.warmup:
            times.append(r["seconds"])
    sec = float(np.mean(times))
    val = n_sample / sec
    full = mesh.config(HEADLINE)
    line = {"impl": "reference", "metric": "elements_assembled_per_s", "value": val, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f64/c128", "data": "synthetic",
            "nnz_per_s": val * (o.nz_upper / model.ne),
            "config": {"workload": full.name, "elements": full.ne, "element_type": f"{full.mn}-node/{full.me}-dof",
                       "note": "each step = a bounded sample of the workload: " + note},
            "cpu_baseline": {"value": val, "unit": "elements/s", "cores": 1, "kind": "port",
                             "sample": f"{n_sample} consecutive elements (ide {lo}..{lo + n_sample - 1}) of {note}, oracle faithful mode, g++ -O1 -ffp-contract=off",
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_graft(args):
    import torch
    import torch.distributed as dist
    from movfem_b200 import host

    rank, local_rank, world = rank_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL / c10d print while the communicator is built (the
        # "NCCL version ..." banner) goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return int(t.item())

    sampler = ClockSampler(local_rank); sampler.start()     # spans everything that is timed
    hbm_peak, hbm_src = measured_peaks()
    fp64_peak = host.fp64_peak_tflops(local_rank)
    traffic = ncu_traffic()
    e2e_steps = max(2, min(args.steps, 3))

    # ---------------- N = 1 only: the other BASELINE configs, same bar ----------------
    per_config = {}
    if world == 1 and not args.headline_only:
        for n in (1, 2, 3):
            m = mesh.config(n)
            r = measure_cold(m, local_rank, max(3, min(args.steps, 10)), args.warmup, barrier, e2e_steps, want_pageable=(n == 2))
            ms = r["ms_per_step"]
            e = r["e2e"]
            per_config[f"config{n}"] = {
                "workload": m.name, "elements": m.ne, "element_type": f"{m.mn}-node/{m.me}-dof", "nne": r["nne"], "nnz_delivered": r["nz"],
                "nnz_stripped_by_rem_zeros": r["nz_upper"] - r["nz"], "pairs_reevaluated_in_reference_order": r["nflagged"],
                "ms_per_step": ms, "value": m.ne / (ms * 1e-3), "unit": "elements/s", "nnz_per_s": r["nz"] / (ms * 1e-3), "phases_ms": r["phases_ms"],
                "roofline": step_roofs(m.me, m.ne, r["nz"], r["phases_ms"], ms, fp64_peak, hbm_peak, traffic.get(m.name)),
                "e2e": {"value": m.ne / e["s_per_step"], "unit": "elements/s", "ms_per_step": e["s_per_step"] * 1e3, "h2d_bytes_per_step": e["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": e["d2h_bytes_per_step"], "ms_h2d": e["ms_h2d"], "ms_d2h": e["ms_d2h"],
                        "keep_pattern_ms_per_step": e["keep_pattern_variant"]["s_per_step"] * 1e3,
                        **({"pageable_ms_per_step": e["pageable_variant"]["s_per_step"] * 1e3} if "pageable_variant" in e else {})},
            }
            if not args.no_cpu_baseline:
                cm, note = cpu_sample_model(n)
                per_config[f"config{n}"]["cpu_baseline"] = cpu_port_sample(cm, REF_SAMPLE[n], note)

    # ---------------- configs[3]: the frequency sweep, dealt round-robin to the N GPUs ----------------
    sweep = None
    if not args.headline_only:
        m4 = mesh.config(4)
        s = measure_sweep(m4, local_rank, rank, world, barrier)
        ms_sweep = allmax(s["ms_shard"])
        if rank == 0:
            nf = s["frequencies"]
            flops = FLOPS_PER_ELEMENT[m4.me] * m4.ne
            sweep = {"workload": m4.name, "elements": m4.ne, "frequencies": nf, "n_gpus": world, "sharding": "frequency round-robin, replicated mesh/pattern/K-M cache, no data-path collective",
                     "ms_sweep_max_over_ranks": ms_sweep, "ms_per_frequency": ms_sweep / nf * world, "value": nf * m4.ne / (ms_sweep * 1e-3), "unit": "elements/s",
                     "nnz_per_s": nf * s["nz"] / (ms_sweep * 1e-3), "scaling": "strong",
                     "rank0_cold_frequency_ms": s["cold_frequency_ms"], "rank0_cached_frequency_ms": s["cached_frequency_ms"],
                     "roofline": {"cold_frequency": {"frac_fp64": flops / (s["cold_frequency_ms"] * 1e-3) * 1e-12 / fp64_peak,
                                                     "frac_hbm": BYTES_PER_ELEMENT[m4.me] * m4.ne / (s["cold_frequency_ms"] * 1e-3) * 1e-9 / hbm_peak},
                                  "cached_frequency": {"frac_hbm_32B_per_nnz": BYTES_PER_NNZ_UPDATE * s["nz"] / (s["cached_frequency_ms"]["ms_total"] * 1e-3) * 1e-9 / hbm_peak,
                                                       "note": "K/M cached: a frequency = node fields + element RHS + A = K + i*w32*M streamed from the gathered cache"}}}
            if world == 1 and not args.no_cpu_baseline:
                cm, note = cpu_sample_model(4)
                sweep["cpu_baseline"] = cpu_port_sample(cm, REF_SAMPLE[4], note)

    # ---------------- headline: config 5, whole mesh (N = 1) or this rank's x-slab ----------------
    model = mesh.config(HEADLINE)
    ne_total = model.ne
    if world > 1:
        model = copy.copy(model)
        model.ie_lo, model.ie_hi = slab_partition(model.g_nx - 1, rank, world)
    r = measure_cold(model, local_rank, args.steps, args.warmup, barrier, e2e_steps)
    ms_per_step = allmax(r["ms_per_step"])
    e2e_s = allmax(r["e2e"]["s_per_step"])
    keep_s = allmax(r["e2e"]["keep_pattern_variant"]["s_per_step"])
    nz_total = allsum(r["nz"])
    launches = allsum(r["launches"])
    h2d, d2h = allsum(r["e2e"]["h2d_bytes_per_step"]), allsum(r["e2e"]["d2h_bytes_per_step"])
    clocks = sampler.stop()
    value = ne_total / (ms_per_step * 1e-3)

    if rank == 0:
        ph = r["phases_ms"]
        ne_rank = ne_total if world == 1 else (model.ie_hi - model.ie_lo + 1) * (model.g_ny - 1) * (model.g_nz - 1)   # owned elements (the +x halo layer is extra work)
        roofs = step_roofs(model.me, ne_rank, r["nz"], ph, r["ms_per_step"], fp64_peak, hbm_peak, traffic.get(model.name))
        flopk = roofs.get("flop_kernels", {})
        line = {
            "metric": "elements_assembled_per_s", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f64/c128", "data": "synthetic",
            "nnz_per_s": nz_total / (ms_per_step * 1e-3),
            "config": {"workload": mesh.config_name(HEADLINE), "elements": ne_total, "element_type": f"{model.mn}-node/{model.me}-dof",
                       "nnz_delivered": nz_total, "frequencies_per_step": 1,
                       "sharding": ("none (one GPU holds the whole mesh)" if world == 1 else
                                    f"{world} x-slabs (ranges of ie): SURVEY 8e's contiguous-row equivalent of north_star's z-slabs -- DOFs are numbered in (ie,je,ke) "
                                    "order, so x-slabs own contiguous row ranges; each rank recomputes its +x halo layer, no data-path collective"),
                       "l2": "flushed between iterations (256 MiB fill, outside the per-step event pair); the working set is far larger than L2 anyway",
                       "cache": "K_e/M_e cache reset every step: full cold assembly"},
            "phases_ms": ph,
            "e2e": {"value": ne_total / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "ms_h2d": r["e2e"]["ms_h2d"], "ms_d2h": r["e2e"]["ms_d2h"],
                    "api": "movfem_assemble (C ABI), pinned host buffers" + ("" if world == 1 else "; every rank delivers its row slice over its own PCIe link "
                                                                           "(the hand-off to the MUMPS host rank: slices of one host array), max over ranks"),
                    "keep_pattern_variant": {"value": ne_total / keep_s, "ms_per_step": keep_s * 1e3,
                                             "note": "MOVFEM_MODE_KEEP_PATTERN: static irn/jcn not re-sent (16 instead of 24 B/entry)"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline_step": roofs["step"],
        }
        # both kernels' rooflines (algorithmic work / device time); `roofline` is the one of the kernel that takes the
        # larger share of the step -- on the headline mesh that is the gather since fused12_kernel replaced geometry + contraction
        tr = traffic.get(model.name) or {}
        ga = roofs.get("gather", {})
        line["roofline_fp64"] = {"bound": "fp64", "kernel": flopk.get("kernels"), "achieved": flopk.get("tflops"), "peak": fp64_peak, "unit": "TFLOP/s",
                                 "frac": flopk.get("frac_fp64"), "traffic": tr.get("fused12_kernel"),
                                 "peak_source": "FP64 FMA-loop microbenchmark run in this process (movfem_fp64_peak); nominal 37.2 TFLOP/s",
                                 "flops_per_element": FLOPS_PER_ELEMENT[model.me], "bytes_per_element": BYTES_PER_ELEMENT[model.me], "ms_kernel": flopk.get("ms"),
                                 "note": "rank 0's launches; the kernel that executes the B^T D B flops SURVEY 8d counts"}
        moved = (tr.get("gather_finalize_kernel") or 0) + (tr.get("rhs_kernel") or 0)
        line["roofline_hbm"] = {"bound": "hbm", "kernel": "gather_finalize_kernel (+ rhs_kernel)", "achieved": ga.get("gbs"), "peak": hbm_peak, "unit": "GB/s",
                                "frac": ga.get("frac_hbm"), "traffic": moved or None, "peak_source": f"MEASURED_PEAKS.json ({hbm_src})",
                                "bytes_per_nnz": BYTES_PER_NNZ_UPDATE, "ms_kernel": ph.get("ms_gather"),
                                "moved": ({"gbs": moved / (ph["ms_gather"] * 1e-3) * 1e-9, "frac": moved / (ph["ms_gather"] * 1e-3) * 1e-9 / hbm_peak,
                                           "what": "DRAM bytes ncu counted for the two kernels (K/M contributions are 1.53 per entry and carry a 4-byte index each: "
                                                   "48.6 B per entry + the element right-hand sides) over the same device time"}
                                          if moved and ph.get("ms_gather") else None),
                                "note": "algorithmic bytes: SURVEY 8d's 32 B per delivered entry (read K 8 + M 8, write A 16)"}
        line["roofline"] = dict(line[dominant_roofline(flopk.get("ms"), ph.get("ms_gather"))],
                                dominant_of={"fp64_kernels_ms": flopk.get("ms"), "gather_ms": ph.get("ms_gather")}, whole=roofs)
        if per_config:
            line["per_config"] = per_config
        if sweep:
            line["sweep"] = sweep
        try:
            line["config"]["library"] = host.lib().movfem_version().decode()
        except Exception:   # pragma: no cover
            pass
        if world == 1 and not args.no_cpu_baseline:
            cm, note = cpu_sample_model(HEADLINE)
            line["cpu_baseline"] = cpu_port_sample(cm, args.ref_elements or REF_SAMPLE[HEADLINE], note)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--ref-elements", type=int, default=0, help="elements in the bounded CPU sample (0: per-config default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip per_config and the sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_graft(args)


if __name__ == "__main__":
    main()
