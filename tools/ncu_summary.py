#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summaries kept under profiles/:
   <out>_summary.md (key metrics per kernel), <out>_instruction_mix_and_stalls.txt (SASS opcode mix, stall reasons,
   most-sampled instructions) and the per-launch DRAM bytes for profiles/r1_dram_traffic.json.
   usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_tma_v9 "<command that produced it>" """
import collections, csv, io, json, re, subprocess, sys

rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def run(*a):
    return subprocess.run(["ncu", "-i", rep] + list(a), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
h, units = raw[0], raw[1]
md = ["# ncu --set full, kernels of one solver iteration at 256^3 (B200, --clock-control none)", "Command: `%s`" % cmd,
      "Report: `%s` (scratch); values are per launch. ncu times are cold-cache and serialised; the live CUDA-event times of the "
      "same kernels are in `bench.py`'s `kernel_ms`." % rep, ""]
traffic = {}
names = []
for r in raw[2:]:
    name = r[h.index("Kernel Name")]
    short = "pass_a" if "pass_a" in name else ("pass_b" if "pass_b" in name else name[:40])
    names.append((short, name))
    md += ["## " + name.split("(")[0], "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in h:
            md.append("| %s | %s | %s |" % (k, r[h.index(k)], units[h.index(k)]))
    md.append("| stalls per issue (warps) | " + ", ".join(
        "%s %.2f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
        for i, n in enumerate(h) if "issue_stalled" in n and "per_issue_active" in n and r[i] and float(r[i]) >= 0.05) + " | |")
    md.append("")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    traffic[short] = int(float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]])
open(out + "_summary.md", "w").write("\n".join(md))

txt = []
for short, _ in names:
    rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + short))))
    hdr, data = rows[1], rows[2:]
    iA, iI, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot, sam = sum(int(r[iI]) for r in data), sum(int(r[iS]) for r in data)
    op = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iA])
        op[m.group(2).split(".")[0] if m else "?"] += int(r[iI])
    txt.append("## %s\ntotal warp inst %d (%.1f thread-instructions per voxel at 256^3) samples %d" % (short, tot, tot * 32 / 256 ** 3, sam))
    txt += ["  %-12s %10d %5.1f%%" % (k, v, 100.0 * v / tot) for k, v in op.most_common(20)]
    st = collections.Counter()
    for i, n in enumerate(hdr):
        if n.startswith("stall_") and "Not Issued" not in n:
            st[n] = sum(int(r[i] or 0) for r in data)
    txt.append("stalls:")
    txt += ["  %-28s %6d %5.1f%%" % (k, v, 100.0 * v / max(1, sam)) for k, v in st.most_common(10)]
    txt.append("top sampled instructions:")
    txt += ["  %6d %9s %s" % (int(r[iS]), r[iI], r[iA].strip()[:90]) for r in sorted(data, key=lambda r: -int(r[iS]))[:16]]
open(out + "_instruction_mix_and_stalls.txt", "w").write("\n".join(txt) + "\n")
print(json.dumps(traffic))
