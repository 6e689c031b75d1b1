"""Per-kernel counts of the tcgen05 / TMEM / TMA instructions in the built library (cuobjdump -sass):
UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor),
UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "kelpie_b200", "libkelpie_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCCP|SYNCS|HMMA|FFMA|MUFU\.EX2)(\.[A-Z0-9_.]+)?")
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        kernels[m.group(1)] = cur
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        op = m.group(1)
        mods = m.group(2) or ""
        tag = op + (".2CTA" if ".2CTA" in mods else "") + (".MULTICAST" if "MULTICAST" in mods else "")
        cur[tag] += 1
print("# cuobjdump -sass kelpie_b200/libkelpie_b200.so: kernels that use the 5th-gen tensor cores / TMEM / TMA")
print("# (instruction counts in the SASS of one kernel, not executions)")
for name, c in kernels.items():
    if not any(k.startswith(("UTCHMMA", "LDTM", "STTM", "UTMALDG")) for k in c):
        continue
    d = demangle(name)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d)
    keys = [k for k in sorted(c) if not k.startswith(("FFMA", "SYNCS"))]
    print(f"{d:45s} " + "  ".join(f"{k}={c[k]}" for k in keys) + f"  SYNCS={c['SYNCS']}  FFMA={c['FFMA']}")
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print("total: " + "  ".join(f"{k}={tot[k]}" for k in sorted(tot) if k.startswith(("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR"))))
