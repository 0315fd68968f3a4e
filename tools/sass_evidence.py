"""per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (profiles/rNN_sass_evidence.txt).
  python tools/sass_evidence.py qwen_inference_engine_b200/libqie_b200.so r02 > profiles/r02_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

so, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "rNN")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.OrderedDict()
pat = [("UTCHMMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
       ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("LDSM", r"\bLDSM"), ("LDGSTS", r"\bLDGSTS"), ("FFMA2", r"\bFFMA2"),
       ("FMUL2", r"\bFMUL2"), ("FADD2", r"\bFADD2"), ("MUFU.EX2", r"\bMUFU\.EX2")]
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["instructions"] += 1
        for k, p in pat:
            if re.search(p, line):
                counts[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {tag}: SASS evidence per kernel of {so} (cuobjdump -sass, sm_100a build of this round)")
print("# counts of the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG/UBLKCP = TMA, SYNCS = mbarrier, HMMA = legacy mma.sync (deliberate where the reference's wmma accumulation must be reproduced),")
print("# FFMA2/FMUL2/FADD2 = packed fp32x2 (sm_100), LDGSTS = cp.async\n")
for name, (k, c) in zip(dem, counts.items()):
    print(name)
    print("    instructions %d  " % c["instructions"] + "  ".join(f"{p} {c[p]}" for p, _ in pat if c[p]))
