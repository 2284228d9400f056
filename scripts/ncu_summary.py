"""Per-kernel summary of an ncu report: python scripts/ncu_summary.py report.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def g(r, name):
    return r[hdr.index(name)] if name in hdr else None
def f(r, name):
    v = g(r, name)
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}
def si(r, name):
    v = f(r, name)
    return None if v is None else v * SCALE.get(units[hdr.index(name)], 1)
out = []
for r in rows[2:]:
    st = []
    for i, h in enumerate(hdr):
        if "stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    st = sorted(st, reverse=True)[:5]
    t = si(r, "gpu__time_duration.sum")
    rd, wr = si(r, "dram__bytes_read.sum"), si(r, "dram__bytes_write.sum")
    d = {"kernel": g(r, "Kernel Name"), "grid": g(r, "Grid Size"), "block": g(r, "Block Size"),
         "duration_us": t * 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr,
         "dram_gbs": (rd + wr) / t / 1e9,
         "dram_pct_of_peak": f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed") or f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "l2_hit_pct": f(r, "lts__t_sector_hit_rate.pct"),
         "lts_throughput_pct": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "sm_throughput_pct": f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
         "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "achieved_occupancy_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
         "warp_inst_executed": f(r, "smsp__inst_executed.sum"),
         "threads_per_inst": f(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
         "registers": f(r, "launch__registers_per_thread"),
         "top_stalls": {h: v for v, h in st}}
    out.append(d)
    print(f"{d['kernel'][:44]:44s} {d['duration_us']:8.1f}us dram {d['dram_gbs']:7.0f} GB/s (R {rd/1e9:.3f} W {wr/1e9:.3f} GB) lts {d['lts_throughput_pct']} sm {d['sm_throughput_pct']} issue {d['issue_active_pct']:.1f}% occ {d['achieved_occupancy_pct']:.0f}% inst {d['warp_inst_executed']:.3g} thr/inst {d['threads_per_inst']:.1f}")
    print("        stalls:", ", ".join(f"{h}={v:.1f}" for v, h in st))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
