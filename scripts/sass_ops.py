"""Count the Blackwell-specific SASS instructions per kernel of libvipe_ba.so (cuobjdump -sass): the evidence file under
profiles/.  Usage: python scripts/sass_ops.py > profiles/rNN_sass_blackwell_ops.txt"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, ".")
from vipe_b200.build import SO  # noqa: E402

OPS = ["UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "LDGMC", "STG.SYS", "FFMA2", "FMUL2", "FADD2", "DMMA",
       "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                counts[cur][o] += 1
        if op.startswith("STG.") and ".STRONG.SYS" in op:  # multimem.st: a system-scope store to the multicast address
            counts[cur]["STG.SYS"] += 1
print("# SASS evidence, libvipe_ba.so: Blackwell-specific instructions per kernel (cuobjdump -sass, counted per function)")
print("# UBLKCP = cp.async.bulk (TMA, 1-D bulk copy); UTCHMMA = tcgen05.mma kind::tf32; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit;")
print("# UTCATOMSWS = tcgen05.alloc/dealloc; SYNCS = mbarrier ops; LDGSTS = cp.async; LDGMC = multimem.ld_reduce, STG.SYS = STG.*.STRONG.SYS = multimem.st to the multicast address (NVSwitch);")
print("# FFMA2/FMUL2/FADD2 = packed fp32; DMMA = fp64 mma.sync; LDL/STL = local memory")
for fn, c in counts.items():
    if c:
        print(f"\n{fn}\n   " + "  ".join(f"{o}={c[o]}" for o in OPS if c[o]))
