#!/usr/bin/env python
"""bench.py -- throughput of the MoVFEM_3DMT element assembly hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (default N=1)
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm (oracle port)

A "step" is one full assembly of one frequency (node fields, every element's K_e/M_e/b_e, the
deterministic gather, A = K + i*w32*M, float32 round trip, zero strip, RHS): the K/M cache is
reset before every step so no work is skipped.  Workload at every N: BASELINE.json configs[1]
(40x40x30 20-node elements, GPML Fang 1996, one frequency per GPU).  For N > 1 the frequency list
is sharded, one frequency per rank, no data-path collective -> weak scaling.

value   device-timed (CUDA events on the launching stream), inputs resident in HBM
e2e     the same metric through the C-ABI host call movfem_assemble with pinned HOST buffers:
        H2D of g_sigma and D2H of irn/jcn/a/rhs inside the timed region
"""
import argparse
import ctypes
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

WORKLOAD = 2     # BASELINE.json configs[1]
FLOPS_PER_ELEMENT = {12: 10944, 36: 250776, 54: 533628}   # 2*ngp*(18*me + 3*me*(me+1))
BYTES_PER_NNZ_UPDATE = 32                                  # read K 8 + M 8, write A 16
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this workload
# (profiles/r01_ncu_full_top_kernels.json, profiles/r01_summary.md); None for other workloads
NCU_TRAFFIC_BYTES = {"contract_kernel": None, "geometry_kernel": None, "gather_finalize_kernel": None}   # filled from profiles/r01_ncu_full.json
try:
    with open(os.path.join(HERE, "profiles", "r01_ncu_full.json")) as _f:
        for _k, _v in json.load(_f).get("dram_bytes_per_step", {}).items():
            NCU_TRAFFIC_BYTES[_k] = _v
except Exception:
    pass


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


def phase_roofs(model, nne, nz_upper, phases_ms, fp64_peak, hbm_peak):
    """Both roofline fractions for every phase of a step, from the ALGORITHMIC work of each phase:
    flops = the B^T D B count (contraction only; Jacobian/basis/RHS flops are deliberately not counted), bytes = the
    data a phase must read and write once.  Every element is taken as unstretched (the GPML layers move 51 instead of
    12 scratch components, so their share is understated).  Pure function of its arguments (tests/test_abi_host.py)."""
    me, ngp = model.me, (8 if model.me == 12 else 27)
    npairs = me * (me + 1) // 2
    node_in = 152 if model.nord == 2 else 1216              # new grid nodes per element x (z 8 + sigma 96 + mu 48) B
    scratch = 12 * ngp * 8                                  # Q (6) | T (6) per Gauss point
    work = {
        "node": (0, model.npt * (152 + 208)),               # read z, sigma, mu; write one node record
        "geometry": (0, model.ne * (node_in + scratch + me * 32)),          # nodes in, Q|T scratch and b_e out
        "contract": (FLOPS_PER_ELEMENT[me] * model.ne, model.ne * (scratch + npairs * 16)),   # scratch in, K_e/M_e out
        "gather": (0, nz_upper * BYTES_PER_NNZ_UPDATE + nne * 64),           # K, M in, A out; RHS rows out (2 columns)
    }
    out = {}
    for name, (flops, nbytes) in work.items():
        ms = phases_ms.get("ms_" + name)
        if not ms or ms <= 0:
            continue
        tf, gbs = flops / (ms * 1e-3) * 1e-12, nbytes / (ms * 1e-3) * 1e-9
        out[name] = {"ms": ms, "tflops": tf if flops else None, "frac_fp64": tf / fp64_peak if flops and fp64_peak else None,
                     "gbs": gbs, "frac_hbm": gbs / hbm_peak if hbm_peak else None}
    return out


def all_cores_port(o, model, omega, sigma, n_sample, lo):
    """Informational: the same arithmetic with the Jacobian memoised per Gauss point and the element matrices computed
    by all host threads (OpenMP; scattered serially in element order, same bits).  NOT the reference's structure -- the
    reference assembles sequentially with every redundant Jacobian rebuild -- so it is reported beside the baseline."""
    nt = os.cpu_count() or 1
    n_sample = min(model.ne, max(n_sample, 4000))          # enough elements per thread to amortise the fork/join
    lo = 1 + (model.ne // 2 // n_sample) * n_sample
    r = o.assemble(omega, sigma, faithful=False, nthreads=nt, want_t1=False, want_t2=False, ide_range=(lo, lo + n_sample - 1))
    return {"value": n_sample / r["seconds"], "unit": "elements/s", "cores": nt, "sample_elements": n_sample,
            "what": "oracle with memoised Jacobians, OpenMP over elements (an optimised CPU port, not the reference's sequential loop)"}


def run_reference(args):
    """The reference's own algorithm on the host cores: the oracle port in `faithful` mode (reference loop
    structure, one alocal per pair, every redundant Jacobian rebuild), single thread because the reference
    assembles sequentially on rank 0 (MoVFEM_3DMT.f90:63).  No Fortran compiler exists in this image, so the
    reference itself cannot be built (oracle/_ref is impossible)."""
    rank, _, world = rank_env()
    if rank != 0:
        return
    from oracle.oracle import Oracle
    model = mesh.config(WORKLOAD)
    o = Oracle(model)
    omega, sigma = model.omega(1), model.sigma_for(1)
    n_sample = min(args.ref_elements, 400)               # ~4 s of CPU work per step
    lo = 1 + (model.ne // 2 // n_sample) * n_sample      # a run of elements from the middle of the mesh
    times = []
    for it in range(args.warmup + args.steps):
        r = o.assemble(omega, sigma, faithful=True, nthreads=1, want_t1=False, want_t2=False, ide_range=(lo, lo + n_sample - 1))
        if it >= args.warmup:
            times.append(r["seconds"])
    sec = float(np.mean(times))
    val = n_sample / sec
    nnz_per_el = o.nz_upper / model.ne
    line = {"impl": "reference", "metric": "elements_assembled_per_s", "value": val, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64/c128", "data": "synthetic",
            "nnz_per_s": val * nnz_per_el,
            "config": {"workload": model.name, "elements": model.ne, "element_type": f"{model.mn}-node/{model.me}-dof",
                       "note": "each step = bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": "elements/s", "cores": 1, "kind": "port",
                             "sample": f"{n_sample} consecutive elements (ide {lo}..{lo + n_sample - 1}) of {model.name}, oracle faithful mode, g++ -O1 -ffp-contract=off",
                             "host_cores_available": os.cpu_count(),
                             "all_cores_port": all_cores_port(o, model, omega, sigma, n_sample, lo)},
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

    model = mesh.config(WORKLOAD)
    # frequency sharding: one frequency per rank, replicated mesh, no collective on the data path
    if world > 1:
        model.freqs = np.logspace(-1, 1, world)
    ifreq = rank + 1
    omega = model.omega(ifreq)
    sigma_np = model.sigma_for(ifreq)

    asm = host.Assembly(model, device=local_rank)
    stream = torch.cuda.Stream(device=dev)
    asm.set_stream(stream.cuda_stream)

    sigma_dev = torch.from_numpy(sigma_np.view(np.float64).reshape(-1)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- device-timed loop ----------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase = {"ms_node": [], "ms_element": [], "ms_geometry": [], "ms_contract": [], "ms_gather": [], "ms_finalize": []}
    launches = 0
    sampler = None
    with torch.cuda.stream(stream):
        for it in range(args.warmup + args.steps):
            if it == 0:
                sampler = ClockSampler(local_rank); sampler.start()     # spans warm-up, the timed region and the e2e loop
            if it == args.warmup:
                torch.cuda.synchronize(dev); barrier(); torch.cuda.synchronize(dev)
                t_wall0 = time.perf_counter()
            flush.zero_()                                   # evict L2 between iterations
            asm.reset_cache()                               # every step is a cold, full assembly
            if it >= args.warmup:
                ev[it - args.warmup][0].record(stream)
            asm.assemble_device(1, omega, sigma_dev.data_ptr(), abi.MODE_T2)
            _, _, _, _, nz = asm.device_result()            # completes the step (zero strip if needed)
            if it >= args.warmup:
                ev[it - args.warmup][1].record(stream)
                st = asm.stats()
                for k in phase:
                    phase[k].append(st[k])
                launches += int(st["launches"]) + 1         # + the L2 flush fill
        torch.cuda.synchronize(dev); barrier(); torch.cuda.synchronize(dev)
        t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = world * model.ne / (ms_per_step * 1e-3)

    # ---------------- end-to-end loop: C-ABI host call with pinned host buffers ----------------
    pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()   # noqa: E731
    h_sigma = pin(sigma_np.size * 2, torch.float64); h_sigma.numpy()[:] = sigma_np.view(np.float64).reshape(-1)
    h_irn, h_jcn = pin(asm.nz_upper, torch.int32), pin(asm.nz_upper, torch.int32)
    h_a, h_rhs = pin(asm.nz_upper * 2, torch.float64), pin(asm.nne * 4, torch.float64)
    sig_c = h_sigma.numpy().view(np.complex128)
    a_c, rhs_c = h_a.numpy().view(np.complex128), h_rhs.numpy().view(np.complex128)
    e2e_t = []
    nz_e2e = 0
    e2e_steps = max(3, min(args.steps, 10))
    for it in range(2 + e2e_steps):
        if it == 2:
            torch.cuda.synchronize(dev); barrier()
        asm.reset_cache()
        t0 = time.perf_counter()
        _, _, _, _, nz_e2e = asm.global_vfem(1, omega, sig_c, mode=abi.MODE_T2, irn=h_irn.numpy(), jcn=h_jcn.numpy(), a=a_c, rhs=rhs_c)
        t1 = time.perf_counter()
        if it >= 2:
            e2e_t.append(t1 - t0)
    # opt-in variant: the caller keeps irn/jcn between frequencies (MOVFEM_MODE_KEEP_PATTERN), 16 instead of 24 B/entry D2H
    keep_t = []
    for it in range(2 + e2e_steps):
        asm.reset_cache()
        t0 = time.perf_counter()
        asm.global_vfem(1, omega, sig_c, mode=abi.MODE_T2 | abi.MODE_KEEP_PATTERN, irn=h_irn.numpy(), jcn=h_jcn.numpy(), a=a_c, rhs=rhs_c)
        t1 = time.perf_counter()
        if it >= 2:
            keep_t.append(t1 - t0)
    te = torch.tensor([float(np.mean(e2e_t))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_stats = asm.stats()
    clocks = sampler.stop()
    h2d = sigma_np.size * 16
    d2h = nz_e2e * 24 + asm.nne * 32

    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        fp64_peak = host.fp64_peak_tflops(local_rank)
        ms_el = float(np.mean(phase["ms_element"])); ms_ga = float(np.mean(phase["ms_gather"]))
        ms_con = float(np.mean(phase["ms_contract"])); ms_geo = float(np.mean(phase["ms_geometry"]))
        flops = FLOPS_PER_ELEMENT[model.me] * model.ne
        ach = flops / (ms_con * 1e-3) * 1e-12          # the contraction kernel executes exactly the flops SURVEY 8d counts
        ach_path = flops / (ms_el * 1e-3) * 1e-12      # same count over geometry + contraction
        ga_bytes = BYTES_PER_NNZ_UPDATE * asm.nz_upper
        line = {
            "metric": "elements_assembled_per_s", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64/c128", "data": "synthetic",
            "nnz_per_s": world * nz / (ms_per_step * 1e-3),
            "config": {"workload": model.name, "elements": model.ne, "element_type": f"{model.mn}-node/{model.me}-dof",
                       "nne": asm.nne, "nnz_delivered": int(nz), "frequencies_per_gpu": 1, "sharding": "frequency" if world > 1 else "none",
                       "l2": "flushed between iterations (256 MiB fill, outside the per-step event pair)",
                       "cache": "K_e/M_e cache reset every step: full cold assembly"},
            "phases_ms": {k: float(np.mean(v)) for k, v in phase.items()},
            "wall_s_timed_region": t_wall,
            "e2e": {"value": world * model.ne / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3, "ms_h2d": e2e_stats["ms_h2d"], "ms_d2h": e2e_stats["ms_d2h"],
                    "api": "movfem_assemble (C ABI) with pinned host buffers",
                    "keep_pattern_variant": {"value": model.ne / float(np.mean(keep_t)), "ms_per_step": float(np.mean(keep_t)) * 1e3,
                                             "d2h_bytes_per_step": int(nz_e2e * 16 + asm.nne * 32),
                                             "note": "rank 0, MOVFEM_MODE_KEEP_PATTERN: irn/jcn (static pattern) not re-sent; opt-in, not the headline"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "contract_kernel (plain + GPML launches): the B^T D B contractions SURVEY 8d counts",
                         "achieved": ach, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None, "traffic": NCU_TRAFFIC_BYTES["contract_kernel"],
                         "traffic_note": "dram bytes per step over both launches, ncu --set full capture (profiles/r01_ncu_full.json)",
                         "peak_source": "FP64 FMA-loop microbenchmark run in this process (movfem_fp64_peak); nominal 37.2",
                         "flops_per_element": FLOPS_PER_ELEMENT[model.me], "ms_kernel": ms_con,
                         "element_path": {"kernels": "geometry_kernel + contract_kernel", "ms": ms_el, "ms_geometry": ms_geo, "achieved": ach_path,
                                          "frac": ach_path / fp64_peak if fp64_peak else None,
                                          "note": "same algorithmic flop count over the whole per-element path (geometry flops are not counted)"}},
            "roofline_hbm": {"bound": "hbm", "kernel": "gather_finalize_kernel", "achieved": ga_bytes / (ms_ga * 1e-3) * 1e-9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": ga_bytes / (ms_ga * 1e-3) * 1e-9 / hbm_peak, "traffic": NCU_TRAFFIC_BYTES["gather_finalize_kernel"],
                             "peak_source": f"MEASURED_PEAKS.json ({hbm_src})", "bytes_per_nnz": BYTES_PER_NNZ_UPDATE, "ms_kernel": ms_ga},
        }
        try:    # which library produced the numbers (A/B builds announce their switches in the version string)
            line["config"]["library"] = host.lib().movfem_version().decode()
        except Exception:   # pragma: no cover
            pass
        try:    # informational table; never allowed to cost the line
            line["phase_roofs"] = phase_roofs(model, asm.nne, asm.nz_upper, line["phases_ms"], fp64_peak, hbm_peak)
        except Exception as exc:   # pragma: no cover
            line["phase_roofs"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            from oracle.oracle import Oracle
            o = Oracle(model)
            n_sample = args.ref_elements
            lo = 1 + (model.ne // 2 // n_sample) * n_sample
            r = o.assemble(omega, sigma_np, faithful=True, nthreads=1, want_t1=False, want_t2=False, ide_range=(lo, lo + n_sample - 1))
            line["cpu_baseline"] = {"value": n_sample / r["seconds"], "unit": "elements/s", "cores": 1, "kind": "port",
                                    "sample": f"{n_sample} consecutive elements (ide {lo}..{lo + n_sample - 1}) of {model.name}, oracle faithful mode "
                                              f"(reference loop structure), g++ -O1 -ffp-contract=off, {r['seconds']:.1f} s",
                                    "host_cores_available": os.cpu_count(),
                                    "all_cores_port": all_cores_port(o, model, omega, sigma_np, n_sample, lo)}
        print(json.dumps(line), flush=True)
    asm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--ref-elements", type=int, default=4000, help="elements in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_graft(args)


if __name__ == "__main__":
    main()
