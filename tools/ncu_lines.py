"""Per-source-line warp-stall samples of one `ncu --set full --import-source on` report (built with -lineinfo):
  python tools/ncu_lines.py report.ncu-rep [min_percent]"""
import csv, subprocess, sys
path = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) >= 8 and r[0].isdigit():
        lines.append((cur_file, int(r[0]), r[1], int(r[6]) if r[6].isdigit() else 0, int(r[7]) if r[7].isdigit() else 0))
tot = sum(l[3] for l in lines)
print("total samples", tot)
for f, n, src, s, ex in lines:
    if s >= tot * minp / 100:
        print("%-14s %5d %5.1f%% %11d  %s" % (f, n, 100 * s / tot, ex, src.strip()[:110]))
