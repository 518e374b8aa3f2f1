#!/usr/bin/env python
"""SASS-level stall view of one kernel from an .ncu-rep (ncu --set full --import-source on): warp-stall samples per
warp role (grouped by how often an instruction executes), per mbarrier wait site, and the top instructions by samples.
usage: python profiles/sassstat.py x.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
print(rows[hi - 1][1] if hi else "")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    try:
        data.append((r[1].strip(), int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]]), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data)
print("samples", tot, "warp instructions", sum(d[2] for d in data))
# roles: the most common execution count is the consumer warps' (once per warp-tile)
cnt = collections.Counter(d[2] for d in data if d[2] > 0)
per_tile = max(cnt, key=lambda c: cnt[c] * c)
print("consumer warp-tiles", per_tile)
groups = collections.OrderedDict()
for name, smp, ins, r in data:
    g = "consumer (>= once per warp-tile)" if per_tile * 0.9 <= ins <= per_tile * 3 else ("other roles / slow paths" if ins > 0 else "never executed")
    G = groups.setdefault(g, [0, 0, collections.Counter()])
    G[0] += smp
    G[1] += ins
    for s in stalls:
        G[2][s[6:]] += int(r[idx[s]] or 0)
for g, (s, i, c) in groups.items():
    print("%-34s %5.1f%% of samples, %.1f instr per warp-tile; stalls: %s" % (g, 100.0 * s / max(tot, 1), i / per_tile, ", ".join("%s %.1f%%" % (k, 100.0 * v / max(s, 1)) for k, v in c.most_common(8))))
print("mbarrier wait sites (TRYWAIT followed by the branch that holds the samples):")
for n, (name, smp, ins, r) in enumerate(data):
    if "TRYWAIT" in name and n + 1 < len(data) and ins > 0:
        nxt = max(data[n + 1:n + 6], key=lambda d: d[1])
        print("  %-70s x%-8d samples %5d (%4.1f%%)" % (name[:70], ins, nxt[1], 100.0 * nxt[1] / max(tot, 1)))
print("top instructions by samples:")
for name, smp, ins, r in sorted(data, key=lambda d: -d[1])[:top]:
    st = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("  %5.1f%%  x%-9d %-64s %s" % (100.0 * smp / max(tot, 1), ins, name[:64], " ".join("%s=%d" % (b, a) for a, b in st)))
