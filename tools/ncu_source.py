"""Per-opcode executed-instruction and stall-sample summary of one kernel from an .ncu-rep (needs --import-source on).
   python tools/ncu_source.py report.ncu-rep kernel_regex"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iW, iWi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
tot = 0; byop = {}; samp = {}; data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or not r[iE].isdigit():
        if r and r[0] == "Kernel Name": break
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    e, s = int(r[iE]), int(r[iSm])
    byop[op] = byop.get(op, 0) + e; samp[op] = samp.get(op, 0) + s; tot += e
    data.append((e, s, r[iS].strip(), r[iW], r[iWi]))
print('total warp-instructions', tot, ' total samples', sum(samp.values()))
for op, e in sorted(byop.items(), key=lambda x: -x[1])[:16]:
    print(f"{op:10s} {e:12d} {100*e/tot:5.1f}%  samples {samp[op]}")
print('--- top sampled instructions (samples, executed, sass, smem wavefronts, ideal)')
for e, s, src, w, wi in sorted(data, key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print(s, e, src[:100], w, wi)
