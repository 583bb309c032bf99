#!/usr/bin/env python
"""Per-kernel SASS evidence of libpcf_b200.so: tcgen05 (UTC*MMA, LDTM / STTM, UTCBAR), bulk-copy / TMA (UBLKCP, UTMA*),
mbarrier (SYNCS), programmatic dependent launch (ACQBULK-free `griddepcontrol` shows up as DEPBAR / the .wait on SR), legacy
mma.sync (HMMA: must be absent).  Runs on CPU (cuobjdump): python scripts/sass_summary.py r02 -> profiles/sass_<tag>.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ml-pointconvformer_b200", "libpcf_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts, sizes = None, collections.OrderedDict(), {}
pats = {"UTCHMMA/UTCxMMA": r"\bUTC\w*MMA", "LDTM": r"\bLDTM", "UTCBAR": r"\bUTCBAR", "UBLKCP": r"\bUBLKCP", "UTMA": r"\bUTMA\w+",
        "SYNCS(mbarrier)": r"\bSYNCS", "LDGSTS(cp.async)": r"\bLDGSTS", "FFMA2": r"\bFFMA2", "HMMA(legacy)": r"\bHMMA", "ACQBULK/griddep": r"ACQBULK|GRIDDEP"}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*$", "", cur).replace("void ", "").replace("pcfb::", "")
        counts[cur] = collections.Counter(); sizes[cur] = 0
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        sizes[cur] += 1
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
out = os.path.join(ROOT, "profiles", "sass_%s.txt" % tag)
with open(out, "w") as f:
    f.write("# cuobjdump -sass of ml-pointconvformer_b200/libpcf_b200.so (%s): instruction counts per kernel\n" % ", ".join(sorted(set(re.findall(r"sm_\w+", arch)))))
    f.write("%-64s %7s  %s\n" % ("kernel", "instrs", "  ".join(pats)))
    tot = collections.Counter()
    for k, c in counts.items():
        tot.update(c)
        if any(c[x] for x in ("UTCHMMA/UTCxMMA", "LDTM", "UBLKCP", "UTMA", "HMMA(legacy)", "SYNCS(mbarrier)")):
            f.write("%-64s %7d  %s\n" % (k[:64], sizes[k], "  ".join("%d" % c[p] for p in pats)))
    f.write("%-64s %7d  %s\n" % ("TOTAL over %d kernels" % len(counts), sum(sizes.values()), "  ".join("%d" % tot[p] for p in pats)))
print(open(out).read())
