"""Registers / spills per kernel of the library build (ptxas -v), names demangled.
    python benchmarks/ptxas_report.py [-Dmacro=value ...]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from devis_b200 import build  # noqa: E402

cmd = [build.nvcc_path()] + build.NVCC_FLAGS + ["-Xptxas", "-v"] + sys.argv[1:] + ["-o", "/tmp/ptxas_report.so"] + build.SOURCES
err = subprocess.run(cmd, capture_output=True, text=True).stderr
name, spill = None, ""
for line in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if "spill" in line:
        spill = line.strip()
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        short = re.sub(r"devis::|\(devis::.*", "", name)
        print(f"{int(m.group(1)):4d} regs  {spill:60s} {short}")
        name = None
