"""Summarise an `ncu --set full` report: per kernel launch the metrics DESIGN.md / bench.py quote.

    python tools/summarize_ncu.py gpurun_out/r01_full.ncu-rep > profiles/r01_ncu_full.json
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
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}

def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kernels, tot = [], {}
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] != "":
                v = float(r[i].replace(",", ""))
                k[WANT[h]] = v * UNIT.get(units[i], 1.0) if units[i] in UNIT else v
        kernels.append(k)
        name = "contract_kernel" if "contract_kernel" in k["kernel"] else "geometry_kernel" if "geometry_kernel" in k["kernel"] else \
            "gather_finalize_kernel" if "gather_finalize" in k["kernel"] else None
        if name:
            tot[name] = tot.get(name, 0.0) + k.get("dram_read", 0.0) + k.get("dram_write", 0.0)
    print(json.dumps({"source": rep, "note": "one cold assembly of config 2; dram_bytes_per_step sums the launches of a kernel within the step",
                      "dram_bytes_per_step": tot, "launches": kernels}, indent=1))

if __name__ == "__main__":
    main(sys.argv[1])
