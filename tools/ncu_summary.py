"""Summarise an `ncu --set full` report: one column per captured launch, the metrics DESIGN.md / VERDICT.md talk about.

    python tools/ncu_summary.py report.ncu-rep [--json out.json label ...]
With --json the per-launch DRAM bytes are also written as {"<label>": {"dram_bytes_read": .., "dram_bytes_write": ..}} for the
labels given in launch order (bench.py reads profiles/r2_ncu_traffic.json for `roofline.traffic`)."""
import csv
import io
import json
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__cluster_max_active", "max active clusters"),
        ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "MUFU (xu) pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__cycles_active.avg", "smsp cycles active"), ("sm__cycles_elapsed.avg", "sm cycles elapsed")]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print(f"# {rep}: {len(data)} launches")
    for i, r in enumerate(data):
        print(f"# launch {i}: {r[kn][:110]}")
    for key, label in WANT:
        if key not in hdr:
            continue
        c = hdr.index(key)
        print(f"{label:28s} [{units[c]:>14s}] " + "  ".join(f"{r[c]:>12s}" for r in data))
    if "--json" in sys.argv:
        j = sys.argv.index("--json")
        out, labels = sys.argv[j + 1], sys.argv[j + 2:]
        cr, cw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        try:
            d = json.load(open(out))
        except Exception:
            d = {}
        for lab, r in zip(labels, data):
            if lab != "-":
                d[lab] = {"dram_bytes_read": float(r[cr]) * scale[units[cr]], "dram_bytes_write": float(r[cw]) * scale[units[cw]],
                          "report": rep.split("/")[-1], "kernel": r[kn][:80]}
        json.dump(d, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
