"""SASS opcode summary per kernel of the built objects (build/csrc/*.o) -> profiles/r02_sass_summary.txt.
Evidence for the tensor-core / TMA / vector-width claims: UTCHMMA / UTCBAR / LDTM / STTM (tcgen05 + TMEM), UBLKCP
(bulk-async stores), LDG.E.128 / STG.E.128 (16-byte gathers), RED / ATOM (atomics). Runs on the build host (no GPU)."""
import collections
import glob
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDG.E.128", "LDG.E.64", "LDG.E",
        "STG.E.128", "STG.E", "LDS.128", "STS.128", "RED.E", "REDG", "ATOMG", "ATOMS", "ATOM.E", "FFMA", "HMMA", "BAR.SYNC", "SYNCS", "CCTL", "PREFETCH")
for obj in sorted(glob.glob(os.path.join(root, "build", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, per = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
                    break
    out.append(f"== {os.path.basename(obj)}")
    for fn, cnt in per.items():
        name = re.sub(r"\(.*", "", fn)
        sel = ", ".join(f"{k} {v}" for k, v in cnt.items() if k != "_total")
        out.append(f"  {name[:110]:110s} instr {cnt['_total']:6d} | {sel}")
text = "\n".join(out) + "\n"
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles", "r02_sass_summary.txt")
open(dst, "w").write(text)
print(text[:3000])
