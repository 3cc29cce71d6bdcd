#!/usr/bin/env python
"""Text timeline of the CUPTI trace written by tools/trace_pcg.py: the last operator apply and the last complete PCG
iteration of rank 0 (kernel, stream, start relative to the first kernel of the window, duration, gap to the previous
kernel on the same stream), plus busy / idle time of the window.
   python tools/trace_summary.py gpurun_out/<tag>_trace.json"""
import json
import sys

ev = json.load(open(sys.argv[1]))
ev = [e for e in ev if e["dur"] > 0]


def short(n):
    n = n.replace("(anonymous namespace)::", "").replace("libp_b200::", "").replace("void ", "")
    return n.split("(")[0][:70]


def window(evs, title):
    if not evs:
        return
    t0 = evs[0]["start"]
    t1 = max(e["start"] + e["dur"] for e in evs)
    print(f"== {title}: {len(evs)} kernels, {t1 - t0:.1f} us")
    last = {}
    busy = []
    for e in evs:
        s = e.get("stream")
        gap = e["start"] - last[s] if s in last else 0.0
        last[s] = e["start"] + e["dur"]
        busy.append((e["start"], e["start"] + e["dur"]))
        print(f"  +{e['start'] - t0:9.1f} us  {e['dur']:8.1f} us  stream {s}  gap {gap:6.1f}  {short(e['name'])}")
    busy.sort()
    tot, cur_s, cur_e = 0.0, busy[0][0], busy[0][1]
    for s, e in busy[1:]:
        if s > cur_e:
            tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    tot += cur_e - cur_s
    print(f"  GPU busy (union over streams) {tot:.1f} us = {100 * tot / (t1 - t0):.1f} % of the window")


names = [short(e["name"]) for e in ev]
# operator applies: windows that start with the masked zero-fill (or the chain kernel on the main stream)
ax = [i for i, n in enumerate(names) if "ax_hex3d_chain_kernel" in n]
zf = [i for i, n in enumerate(names) if n.startswith("zero_fill_kernel")]
if zf:
    i0 = zf[-1]
    end = max([i for i in ax if i >= i0] + [i0])
    # include exchange kernels that belong to the apply (until the next zero fill or the end)
    j = end
    while j + 1 < len(ev) and "zero_fill" not in names[j + 1] and ev[j + 1]["start"] < ev[end]["start"] + ev[end]["dur"] + 50:
        j += 1
    window(ev[i0:j + 1], "last operator apply")
# PCG iterations: delimited by the p-update kernel
pu = [i for i, n in enumerate(names) if "pupdate_kernel" in n]
if len(pu) >= 3:
    window(ev[pu[-3]:pu[-2]], "one Jacobi-PCG iteration (p-update to p-update)")
