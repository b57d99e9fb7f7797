#!/usr/bin/env python
"""Static SASS of one kernel grouped by CUDA source line (no GPU needed).
Usage: sass_lines.py lib.so kernel_substring first_line last_line
Needs -lineinfo at compile time; uses cuobjdump -xelf + nvdisasm -g."""
import glob, os, re, subprocess, sys, tempfile

so, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, stdout=subprocess.DEVNULL)
txt = "".join(subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout for f in glob.glob(d + "/*.cubin"))
sec, cur, counts = False, None, {}
for l in txt.split("\n"):
    if l.strip().startswith(".section"):
        sec = kern in l and ".text." in l
        continue
    if not sec:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = int(m.group(2))
        continue
    if re.search(r"/\*[0-9a-f]{4}\*/", l) and cur is not None:
        counts[cur] = counts.get(cur, 0) + 1
        if lo <= cur <= hi:
            print("%4d  %s" % (cur, re.sub(r"\s+", " ", l.split("*/", 1)[1]).strip()[:110]))
print("static instructions per line:", {k: v for k, v in sorted(counts.items()) if lo <= k <= hi})
print("total static instructions:", sum(counts.values()))
