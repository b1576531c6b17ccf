"""SASS opcode histogram of every object of libsug_b200 (cuobjdump -sass sug_b200/build/*.o): the evidence that the
tensor-core / TMA / TMEM paths are the Blackwell-native ones (UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA,
LDTM / STTM = tcgen05.ld / st; no HMMA = no legacy mma.sync).  python tools/sass_histogram.py > profiles/<tag>_sass_opcodes.md"""
import collections, glob, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "SYNCS",
       "HMMA", "IMMA", "FMNMX3", "FFMA2", "FADD2", "RED", "ATOMG", "ATOMS", "ATOM", "LDGSTS", "DADD", "DFMA", "ELECT", "FFMA", "LDS", "STS", "LDG", "STG")
print("# SASS opcode histogram (sm_100a), per translation unit\n")
print("`cuobjdump -sass sug_b200/build/<unit>.o`, instruction mnemonics before the first `.`; Blackwell-specific ones first.\n")
for obj in sorted(glob.glob(os.path.join(root, "sug_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cnt, kernels = collections.Counter(), 0
    for ln in out.splitlines():
        if "Function :" in ln:
            kernels += 1
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            cnt[m.group(1)] += 1
    total = sum(cnt.values())
    hot = ", ".join(f"{k} {cnt[k]}" for k in KEY if cnt[k])
    rest = ", ".join(f"{k} {v}" for k, v in cnt.most_common(8))
    print(f"* **{os.path.basename(obj)[:-2]}.cu** ({kernels} kernels, {total} instructions): {hot}\n  * most frequent: {rest}")
