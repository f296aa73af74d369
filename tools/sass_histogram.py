"""SASS opcode histogram per kernel of the built library (evidence of the Blackwell-native path: UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = bulk copy, STSM/LDSM, MUFU, HFMA2 ...).
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "uformer_b200", "lib", "liblewin_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "STSM", "LDSM", "LDGSTS", "MUFU", "HFMA2", "HMUL2", "HMNMX2", "FFMA2", "FFMA",
       "SYNCS", "LDG", "STG", "LDS", "STS", "HMMA", "BAR", "USETMAXREG"]
cur, hist = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass), selected opcodes")
print("kernel".ljust(46) + " ".join(k.rjust(8) for k in KEY) + "   total")
for k, c in hist.items():
    print(k[:45].ljust(46) + " ".join(str(c.get(o, 0)).rjust(8) for o in KEY) + f"   {sum(c.values())}")
tot = collections.Counter()
for c in hist.values():
    tot.update(c)
print("ALL".ljust(46) + " ".join(str(tot.get(o, 0)).rjust(8) for o in KEY) + f"   {sum(tot.values())}")
