"""Opcode evidence for the built library: per kernel, the count of the SASS mnemonics that prove (or would disprove) a
Blackwell-native path -- tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, mbarrier = SYNCS,
legacy tensor path = HMMA (must be absent), memory barriers = MEMBAR.*.
    python profiles/sass_histogram.py kandinsky-2_b200/libk2b200.so > profiles/sass_histogram_r2.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "MEMBAR.ALL.GPU",
        "MEMBAR.ALL.CTA", "LDG", "STG", "LDS", "STS", "FFMA", "BAR"]
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"k2::\(anonymous namespace\)::|\(.*", "", cur).replace("void ", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k.endswith(".2CTA") and ".2CTA" in op and op.startswith("UTCHMMA")):
                per[cur][k] += 1
tot = collections.Counter()
print(f"{'kernel':58s} " + " ".join(f"{k.replace('MEMBAR.ALL.', 'MB.'):>8s}" for k in ["total"] + KEYS))
for name, c in per.items():
    tot.update(c)
    print(f"{name[:58]:58s} " + " ".join(f"{c[k]:8d}" for k in ["total"] + KEYS))
print(f"{'ALL KERNELS':58s} " + " ".join(f"{tot[k]:8d}" for k in ["total"] + KEYS))
