"""Hot source lines of one kernel in an ncu report (captured with --import-source on, built with -lineinfo):
python scripts/ncu_hot_lines.py report.ncu-rep <kernel regex> [top N]
Per source line: stall samples, warp instructions executed, average active threads, dominant stall reasons."""
import csv, io, subprocess, sys


def num(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0


rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0] == "Kernel Name" and lines:
        break    # first launch only
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        stalls = sorted(((num(v), k[6:]) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:3]
        lines.append((num(d["# Samples"]), num(d["Instructions Executed"]), d["Avg. Threads Executed"], cur_file, r[0],
                      r[1].strip()[:90], ", ".join(f"{k}={v}" for v, k in stalls if v)))
ts, ti = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"total samples {ts}, warp instructions {ti}")
for s, i, thr, f, ln, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst thr {thr:>3} {f}:{ln:<5} {src}   [{st}]")
