#!/usr/bin/env python
"""Registers / stack / static shared memory of every kernel in the product library (cuobjdump -res-usage)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "procgen_b200/libprocgen_b200.so"
out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
name = None
for line in out.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        continue
    if name and "REG:" in line:
        short = re.sub(r"\(anonymous namespace\)::|pg::", "", name)
        short = re.sub(r"\(.*\)$", "", short)
        print(f"{short:60s} {line.strip()}")
        name = None
