"""Opcode histogram per kernel of the built library (cuobjdump -sass), filtered to the mnemonics that prove the
Blackwell-native paths: tcgen05 (UTCHMMA / UTCBAR / LDTM / UTCATOM...), TMA (UTMALDG / UBLKCP / UTMAPF), mbarrier (SYNCS),
mma.sync / ldmatrix (HMMA / LDSM), cluster ops (UCGABAR / MAPA), PDL (ACQBULK / DEPBAR-like griddepcontrol).
usage: python scripts/sass_opcodes.py [lib.so] > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hqtransformer_b200", "libhqgraft.so")
INTEREST = re.compile(r"^(UTC|UTMA|UBLK|LDTM|STTM|SYNCS|HMMA|LDSM|UCGABAR|MAPA|ACQBULK|REDUX|ERRBAR|CCTL|MEMBAR|ATOMG|RED|BAR|UTMAPF|ELECT|FENCE|NANOSLEEP)")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name, hist, total = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        hist[name], total[name] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and name:
        total[name] += 1
        op = m.group(1)
        if INTEREST.match(op):
            hist[name][op] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a); only async / tensor / sync mnemonics listed")
agg = collections.Counter()
for k in sorted(hist):
    if not hist[k]:
        continue
    agg.update(hist[k])
    print(f"\n{k}   [{total[k]} instructions]")
    print("   " + "  ".join(f"{op} x{n}" for op, n in sorted(hist[k].items())))
print("\n# whole library")
print("   " + "  ".join(f"{op} x{n}" for op, n in sorted(agg.items())))
