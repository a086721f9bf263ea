"""SASS opcode census of the built library (cuobjdump -sass): DMMA / UTMALDG / UBLKCP / SYNCS / LDGSTS / DFMA per kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "admm_project_b200", "libadmm_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "LDS", "STS", "MEMBAR", "ATOM", "RED"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + "."):
                per[cur][o] += 1
                break
        else:
            if op.startswith("ATOMG") or op.startswith("ATOMS"):
                per[cur]["ATOM"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in per.values():
    tot.update(c)
tag = sys.argv[1] if len(sys.argv) > 1 else "round 2"
print("SASS opcode census of admm_project_b200/libadmm_b200.so (cuobjdump -sass, sm_100a), %s" % tag)
print("TMA = UTMALDG (cp.async.bulk.tensor) / UBLKCP (cp.async.bulk 1-D); SYNCS = mbarrier ops; DMMA = FP64 tensor-core MMA; LDGSTS = cp.async\n")
print("whole library: " + ", ".join("%s %d" % (o, tot[o]) for o in ops) + "\n")
print("%-84s %6s %7s %6s %6s %6s %6s" % ("kernel", "DMMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "DFMA"))
for (k, c), name in zip(per.items(), demangle):
    name = re.sub(r"\(.*", "", name)
    if not any(c[o] for o in ("DMMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS")) and c["DFMA"] < 40:
        continue
    print("%-84s %6d %7d %6d %6d %6d %6d" % (name[:84], c["DMMA"], c["UTMALDG"], c["UBLKCP"], c["SYNCS"], c["LDGSTS"], c["DFMA"]))
