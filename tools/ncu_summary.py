"""Summarises an `ncu --page raw --csv` export: one line per captured launch with duration, DRAM traffic, achieved
DRAM throughput, occupancy and registers.  Usage: python tools/ncu_summary.py gpurun_out/prof_X.raw.csv [> profiles/...]"""
import csv
import sys

COLS = {"gpu__time_duration.sum": "dur", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
        "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "lts__t_bytes.sum": "l2_bytes",
        "launch__occupancy_limit_shared_mem": "occ_lim_smem", "launch__occupancy_limit_registers": "occ_lim_regs",
        "l1tex__t_bytes.sum": "l1_bytes", "smsp__cycles_active.avg": "cyc"}


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main(path):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    print("| kernel | grid×block | regs | dur µs | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 MB | occ % | sm % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0][-60:]
        g = {}
        for m, k in COLS.items():
            if m in idx:
                v = to_num(r[idx[m]])
                u = units[idx[m]]
                if v is not None and u in scale and k in ("dur", "rd", "wr", "l2_bytes", "l1_bytes"):
                    v *= scale[u]
                g[k] = v
        dur = g.get("dur") or 0.0
        rd, wr = g.get("rd") or 0.0, g.get("wr") or 0.0
        gbs = (rd + wr) / dur / 1e3 if dur else 0.0
        print(f"| {name} | {int(g.get('grid') or 0)}×{int(g.get('block') or 0)} | {int(g.get('regs') or 0)} | {dur:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
              f"{gbs:.0f} | {g.get('dram_pct') or 0:.1f} | {(g.get('l2_bytes') or 0) / 1e6:.1f} | {g.get('occ_pct') or 0:.1f} | {g.get('sm_pct') or 0:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
