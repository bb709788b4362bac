"""cuobjdump -sass of libls_b200.so -> opcode counts per kernel (profiles/r2_sass_opcode_counts.txt).
usage: python tools/sass_counts.py > profiles/r2_sass_opcode_counts.txt"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "livelyspeaker_b200", "csrc", "libls_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "FFMA2", "FADD2", "FMUL2", "MUFU", "SYNCS", "LDS", "STS", "SHFL", "LDL", "STL"]
print("# SASS evidence (round 2, final build): cuobjdump -sass livelyspeaker_b200/csrc/libls_b200.so, opcode counts per kernel")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA engine), UTCBAR = tcgen05.commit,")
print("# FADD2/FMUL2/FFMA2 = packed fp32x2 arithmetic (sm_100a), SYNCS = mbarrier ops, LDL/STL = register spills\n")
name, counts, total = None, {}, 0
def flush():
    if name:
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", demangled.replace("(anonymous namespace)::", ""))[:90]
        print("%-92s total=%6d  %s" % (short, total, "  ".join("%s=%d" % (k, counts[k]) for k in KEYS if counts.get(k))))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, total = m.group(1), {}, 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        total += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[k] = counts.get(k, 0) + 1
flush()
