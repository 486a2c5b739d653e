"""Aggregate warp-stall samples of an ncu report per CUDA source line: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[1][1][:100])
lines = [r for r in rows[3:] if len(r) > 7 and r[0].strip().isdigit() and r[6].strip().isdigit()]
tot = sum(int(r[6]) for r in lines)
print("total samples", tot)
for r in sorted(lines, key=lambda r: -int(r[6]))[:top]:
    print(f"{int(r[6]):6d} {100 * int(r[6]) / tot:5.1f}%  L{r[0]:>4} insts={r[7]:>10}  {r[1][:120]}")
