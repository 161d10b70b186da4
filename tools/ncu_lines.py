"""Top CUDA source lines of a kernel from an .ncu-rep (uses `--page source --print-source cuda,sass`).
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = []
fname = ""
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        wi = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and len(r) >= len(hdr) and r[2] == "-":        # a CUDA source line (its SASS rows carry an address)
        try:
            data.append((int(r[ii] or 0), int(r[wi] or 0), fname, r[0], r[1]))
        except ValueError:
            pass
ti, tw = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print("total warp instr", ti, "stall samples", tw)
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{100 * d[0] / ti:5.1f}%i {100 * d[1] / tw:5.1f}%s  {d[2]}:{d[3]:>4}  {d[4][:105]}")
