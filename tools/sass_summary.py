#!/usr/bin/env python
"""Per-kernel SASS evidence of the built library: counts of the Blackwell-specific mnemonics (B200_PROFILING.md: UTC*MMA =
tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor loads, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,
SYNCS = mbarrier), cluster barriers (UCGABAR) and FP64 FMAs.    python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "autoinst_b200", "lib", "libautoinst_ncuts.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "DFMA", "HMMA", "ATOMS", "RED.E", "LDGSTS"]
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = name.replace("void ", "").replace("ancuts::", "")
        rows[cur] = collections.Counter()
        continue
    if cur:
        for p in pats:
            if re.search(r"\b" + re.escape(p), line):
                rows[cur][p] += 1
        rows[cur]["_instr"] += 1 if re.search(r"^\s+/\*[0-9a-f]{4,6}\*/\s+[A-Z@]", line) else 0
print("# SASS summary of libautoinst_ncuts.so (sm_100a), `cuobjdump -sass | grep -c` per kernel\n")
print("Only kernels with at least one of the listed mnemonics besides DFMA are shown in full; the library holds "
      f"{len(rows)} kernels.\n")
print("| kernel | instr | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
for k, c in rows.items():
    if any(c[p] for p in pats if p not in ("DFMA", "RED.E", "ATOMS")) or "lanczos" in k or "affinity" in k or "pair" in k:
        print(f"| `{k[:70]}` | {c['_instr']} | " + " | ".join(str(c[p]) if c[p] else "" for p in pats) + " |")
