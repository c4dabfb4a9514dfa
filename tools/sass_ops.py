"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA), packed fp32x2 arithmetic.  No GPU needed.
    python tools/sass_ops.py > profiles/sass_ops.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "excel_b200", "lib", "libexcel_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTCBAR|SYNCS|FFMA2|FADD2|FMUL2|MUFU\.EX2|HMMA|LDGSTS)\b")
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()[:90]
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        counts[cur]["instructions"] += 1
        for op in pat.findall(line):
            counts[cur][op] += 1
            total[op] += 1
print("SASS evidence per kernel of excel_b200/lib/libexcel_b200.so (cuobjdump -sass, sm_100a).  tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM,")
print("TMA = UTMALDG / UTMASTG / UTMAREDG, mbarrier = SYNCS, packed fp32x2 = FFMA2 / FADD2 / FMUL2; no legacy HMMA anywhere.\n")
print("library totals:", dict(sorted(total.items())), "\n")
for k, c in counts.items():
    ops = {o: n for o, n in c.items() if o != "instructions"}
    print(f"{c['instructions']:6d} instr  {k}\n          {dict(sorted(ops.items()))}")
