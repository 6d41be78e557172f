"""Summarise an `ncu --set full` report: per kernel launch the metrics DESIGN.md / bench.py quote.

    python tools/summarize_ncu.py gpurun_out/r02_full_config5.ncu-rep config5_large_topography > profiles/r02_ncu_config5.json

bench.py reads `workload` and `dram_bytes_per_launch` (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured
launches of a kernel) for the `traffic` fields of its roofline objects.
"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__inst_executed.sum": "warp_inst",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
NAMES = ("fused12_kernel", "contract_kernel", "geometry_kernel", "gather_finalize_kernel", "exact_kernel", "node_kernel", "compact_kernel", "rhs_kernel", "narrow_kernel")


def main(rep, workload):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kernels, tot, cnt = [], {}, {}
    stall_cols = [(i, h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] != "":
                v = float(r[i].replace(",", ""))
                k[WANT[h]] = v * UNIT.get(units[i], 1.0) if units[i] in UNIT else v
        st = sorted(((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for i, h in stall_cols if r[i] not in ("", "n/a")), reverse=True)[:4]
        k["top_stalls_per_issue"] = {n: round(v, 2) for v, n in st}
        kernels.append(k)
        name = next((n for n in NAMES if n in k["kernel"]), None)
        if name and k.get("duration_us", 0.0) >= 20.0:   # (a variant that stood down on a device flag returns within microseconds: not a launch of the kernel's work)
            tot[name] = tot.get(name, 0.0) + k.get("dram_read", 0.0) + k.get("dram_write", 0.0)
            cnt[name] = cnt.get(name, 0) + 1
    print(json.dumps({"source": rep, "workload": workload, "note": "one cold assembly; ncu replays each launch (cold cache, serialised): compare shares, not absolutes",
                      "dram_bytes_per_launch": {n: tot[n] / cnt[n] for n in tot}, "launches_captured": cnt, "launches": kernels}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
