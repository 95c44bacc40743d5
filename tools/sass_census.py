"""Per-kernel SASS census of libsift_b200.so (cuobjdump -sass): the opcodes that prove the Blackwell-specific paths.

    python tools/sass_census.py > profiles/r02_sass_census.txt

  UTMALDG   cp.async.bulk.tensor (TMA) loads            UTMASTG  TMA stores (none: the copy probe found STG.64 as fast)
  UTCIMMA   tcgen05.mma kind::i8 (5th-gen tensor cores) LDTM     tcgen05.ld (TMEM -> registers)
  UTCBAR    tcgen05.commit onto an mbarrier             SYNCS    mbarrier operations
  FFMA2 / FADD2 / FMUL2   packed f32x2 arithmetic       FMNMX3 / VIMNMX3  three-input float / integer min-max
  ACQBULK   griddepcontrol.wait (programmatic dependent launch: first statement of every kernel of a group)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "sift_features_b200", "libsift_b200.so")
OPS = ["UTMALDG", "UTMASTG", "UTCIMMA", "LDTM", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "VIMNMX3",
       "MATCH", "MUFU", "DFMA", "LDS", "STS", "LDG", "STG", "BAR", "ACQBULK"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
names = iter(demangle)
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(names)
        cur = cur.replace("(int)", "").replace("(bool)", "").replace("(anonymous namespace)::", "")
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("sb::", "")
        while cur in counts:
            cur += "'"
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_all"] += 1
        if op in OPS:
            counts[cur][op] += 1
            total[op] += 1
print("# SASS census of sift_features_b200/libsift_b200.so (sm_100a), per kernel: static instruction counts")
print("# " + subprocess.run(["cuobjdump", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1])
cols = [o for o in OPS if total[o]]
print(f"{'kernel':46s} {'instrs':>7s} " + " ".join(f"{c:>7s}" for c in cols))
for k, c in counts.items():
    print(f"{k[:46]:46s} {c['_all']:7d} " + " ".join(f"{c[o]:7d}" if c[o] else f"{'.':>7s}" for o in cols))
print(f"{'TOTAL':46s} {sum(c['_all'] for c in counts.values()):7d} " + " ".join(f"{total[o]:7d}" for o in cols))
print("# UTMASTG = 0: results are stored with STG.64 / STG.128 (tools/probe/copy_probe.cu measured TMA stores no faster for these access patterns)")
