#!/usr/bin/env bash
# registers / spills per kernel of one .cu file of sobfu_b200/csrc (no GPU needed): tools/ptxas_report.sh solver_tiled.cu [extra nvcc flags]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
src="$1"; shift || true
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --ftz=true --prec-div=false --prec-sqrt=false -Xcompiler -fPIC \
    -I"$ROOT/include" -I"$ROOT/sobfu_b200/csrc" "$@" -Xptxas -v -c "$ROOT/sobfu_b200/csrc/$src" -o /dev/null 2>&1 | python3 -c '
import re, subprocess, sys
name = None
for line in sys.stdin:
    m = re.search(r"Compiling entry function .(\w+). for", line)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip(); name = re.sub(r"\(anonymous namespace\)::", "", d.split("(CUtensorMap")[0])[-60:]
        spill = ""
    elif "spill" in line:
        spill = line.strip()
    elif "Used" in line and name:
        r = re.search(r"Used (\d+) registers", line).group(1)
        print("%-62s regs %3s  %s" % (name, r, spill.replace("bytes ", "B ")))
        name = None
'
