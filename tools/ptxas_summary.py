"""Compile phb200.cu with -Xptxas -v and print registers / spills / smem per step kernel."""
import re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "phonomena_b200", "csrc", "phb200.cu")
cmd = ["nvcc", "-O3", "-std=c++20", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
       "--expt-relaxed-constexpr", "-Xptxas", "-v", "-c", "-o", "/tmp/phb_v.o", src] + sys.argv[1:]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
cur = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1); continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        spill = (m.group(1), m.group(2), m.group(3)); continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"phb::", "", name)
        name = re.sub(r"\(.*", "", name)
        rows.append((name, int(m.group(1)), spill))
        cur = None
for name, regs, spill in sorted(rows):
    if "k_step" in name:
        print("%-90s regs=%3d stack=%s spill_st=%s spill_ld=%s" % (name[:90], regs, *spill))
