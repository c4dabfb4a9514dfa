"""Dev: summarise a clock trace of attn_pv_kernel (tools/attn_probe.py run with mask 4096): python tools/pv_trace.py gpurun_out/pv_trace_4096.npy"""
import collections, sys
import numpy as np
b = np.load(sys.argv[1])
def dec(seg):
    seg = seg[seg != 0]
    return (seg >> np.uint64(56)).astype(int), (seg & np.uint64((1 << 56) - 1)).astype(np.int64)
names = {1: 'pre xy_full', 2: 'got xy_full', 3: 'S issued', 4: 'pre p_ready', 5: 'got p_ready', 6: 'got v_full', 7: 'PV issued',
         8: 'pre s_full', 9: 'got s_full', 10: 'math done', 11: 'arrived', 12: 'flush start', 13: 'chunk out', 14: 'flush end'}
for nm, seg in (("MMA warp", b[:4096]), ("epilogue warp, group 0", b[4096:8192]), ("epilogue warp, group 1", b[8192:])):
    t, c = dec(seg)
    if len(t) < 2:
        continue
    d = collections.defaultdict(list)
    for i in range(1, len(t)):
        d[(t[i - 1], t[i])].append(c[i] - c[i - 1])
    print(f"{nm}: {len(t)} events, span {c[-1] - c[0]} clk")
    for k, v in sorted(d.items()):
        v = np.array(v)
        print(f"  {names[k[0]]:>12s} -> {names[k[1]]:<12s} n={len(v):4d} mean {v.mean():7.0f} median {np.median(v):7.0f} total {v.sum():8d}")
