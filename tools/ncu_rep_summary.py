"""Reads an `ncu --set full` report (.ncu-rep) HERE (no GPU needed) and prints / writes the per-kernel evidence the docs cite:
duration, DRAM bytes, issue utilisation, warps active, registers, top stall reasons, instruction mix.

    python tools/ncu_rep_summary.py gpurun_out/x.ncu-rep [--traffic-json profiles/ncu_traffic.json] [--md profiles/x_summary.md]
"""
import argparse
import collections
import csv
import json
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def short(name):
    m = re.search(r"([A-Za-z_0-9]+)(<.*)?\(", name)
    return m.group(1) if m else name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--traffic-json")
    ap.add_argument("--md")
    a = ap.parse_args()
    rows = page(a.rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            v = float(r[col[k]].replace(",", ""))
        except Exception:
            return None
        u = units[col[k]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
        return v * scale.get(u, 1)
    per = collections.OrderedDict()
    for r in data:
        k = short(r[col["Kernel Name"]])
        d = per.setdefault(k, dict(launches=0, us=0.0, dram=0.0, issue=[], warps=[], regs=None, stalls=collections.Counter(), inst=0.0))
        d["launches"] += 1
        d["us"] += f(r, "gpu__time_duration.sum") or 0
        d["dram"] += (f(r, "dram__bytes_read.sum") or 0) + (f(r, "dram__bytes_write.sum") or 0)
        d["issue"].append(f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0)
        d["warps"].append(f(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0)
        d["regs"] = r[col["launch__registers_per_thread"]]
        d["inst"] += f(r, "smsp__inst_executed.sum") or 0
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    d["stalls"][h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] += float(r[i])
                except Exception:
                    pass
    lines = ["| kernel | launches | us/launch | DRAM MB/launch | issue % | warps active % | regs | warp-inst/launch | top stalls (per issued inst) |", "|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for k, d in per.items():
        n = d["launches"]
        st = ", ".join(f"{s} {v / n:.2f}" for s, v in d["stalls"].most_common(4) if s != "selected")
        lines.append(f"| `{k}` | {n} | {d['us'] / n:.1f} | {d['dram'] / n / 1e6:.1f} | {sum(d['issue']) / n:.0f} | {sum(d['warps']) / n:.0f} | {d['regs']} | {d['inst'] / n:.3g} | {st} |")
        traffic[k] = {"launches": n, "dram_bytes_per_launch": int(d["dram"] / n), "dur_us_per_launch": round(d["us"] / n, 2)}
    text = "\n".join(lines)
    print(text)
    if a.md:
        with open(a.md, "w") as fh:
            fh.write(f"ncu --set full summary of `{a.rep}` (tools/ncu_rep_summary.py)\n\n" + text + "\n")
    if a.traffic_json:
        with open(a.traffic_json, "w") as fh:
            json.dump(traffic, fh, indent=1)


if __name__ == "__main__":
    sys.exit(main())
