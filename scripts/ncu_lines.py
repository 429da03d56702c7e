"""Per-source-line stall samples of one kernel: joins the SASS page of an .ncu-rep (ncu -i rep
--page source --csv) with nvdisasm -g line info of the same kernel in the built library.
usage: ncu_lines.py rep.ncu-rep cubin kernel_substring [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = rows[hdr].index("# Samples")
sass = [(r[1].strip(), int(r[col] or 0)) for r in rows[hdr + 1:] if len(r) > col and r[0].startswith("0x")]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn, src = [], 0, False, ""
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        src, cur = m.group(1).split("/")[-1], int(m.group(2))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines.append((src, cur, m.group(1)))
print("sass rows in report: %d, instructions in cubin: %d" % (len(sass), len(lines)))
if len(sass) != len(lines):
    print("MISMATCH: the report was taken on a different build; joining by order anyway")
per = defaultdict(int)
for (s, n), (f, l, _) in zip(sass, lines):
    per[(f, l)] += n
tot = sum(per.values()) or 1
for (f, l), n in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    print("%6.2f%%  %s:%d" % (100.0 * n / tot, f, l))
if len(sys.argv) > 5:   # ranges "a-b,c-d" of the main file summed
    for rg in sys.argv[5].split(","):
        a, b = (int(v) for v in rg.split("-"))
        print("lines %d-%d: %.2f%%" % (a, b, 100.0 * sum(n for (f, l), n in per.items() if a <= l <= b and f.startswith(kern.split("_")[0][:4]) or a <= l <= b) / tot))
