"""Turns an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_spmm.ncu-rep "title line" > profiles/rN_ncu_full_xxx.txt
"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}")
        rd = wr = dur = None
        for w in WANT:
            if w not in idx:
                continue
            print(f"{w} [{units[idx[w]]}] = {r[idx[w]]}")
        try:
            def to_bytes(name):
                v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
            def to_sec(name):
                v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
                return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[u]
            b = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            s = to_sec("gpu__time_duration.sum")
            print(f"derived: dram traffic {b / 1e9:.3f} GB in {s * 1e6:.1f} us = {b / s / 1e9:.0f} GB/s")
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
