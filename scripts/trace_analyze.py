"""Read a PNR_TRACE event log (scripts/gpu_trace.sh) and print the timeline of a few layer hand-offs of CTA pair 0."""
import sys
from collections import defaultdict

NAMES = {0: "W_FULL", 5: "W_EMPTY0", 6: "W_EMPTY1", 7: "W_EMPTY2", 8: "W_EMPTY3", 9: "W_EMPTY4", 10: "IN_READY", 11: "IN_FREE", 12: "X_FULL0", 13: "X_FULL1", 14: "H_FULL0", 15: "H_FULL1",
         16: "RDY0", 17: "RDY1", 18: "RDY2", 19: "RDY3", 20: "X_FREE0", 21: "X_FREE1", 22: "LAND0", 23: "LAND1", 24: "OUT_FREE", 25: "C0_FREE"}


def tagname(t):
    if t in (1, 2, 3, 4):
        return ["", "commit X_FULL0", "commit X_FULL1", "commit H_FULL0", "commit H_FULL1"][t]
    if 0x10 <= t < 0x40:
        return "wait> " + NAMES.get(t - 0x10, str(t - 0x10))
    if 0x40 <= t < 0x70:
        return "wait< " + NAMES.get(t - 0x40, str(t - 0x40))
    return {0x70: "fenced", 0x71: "stage issued", 0x72: "weights ok", 0x80: "tmem read", 0x81: "remote stored", 0x82: "local stored",
            0x83: "published", 0x84: "tma issued", 0xFF: "start"}.get(t, hex(t))


launches = []
for line in open(sys.argv[1]):
    if line.startswith("#"):
        launches.append(defaultdict(list))
        continue
    r, c, t = line.split()
    launches[-1][int(r)].append((int(c), int(t)))
L = launches[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
off = L[2][0][0] - L[1][0][0]      # CTA1 clock - CTA0 clock at the cluster start
print("clock offset CTA1 - CTA0 at start:", off)
ev = []
for r, lst in L.items():
    for c, t in lst:
        ev.append((c - (off if r in (2, 3, 8, 9, 10) else 0), r, t))
ev.sort()
t0 = ev[0][0]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 60000
role = {0: "MMA ", 1: "epi0", 2: "epi1", 3: "rely", 4: "gath", 5: "prd0", 6: "prd1", 7: "prd2", 8: "PRD0", 9: "PRD1", 10: "PRD2"}
skip_stage = "--stages" not in sys.argv
for c, r, t in ev:
    if lo <= c - t0 <= hi:
        if skip_stage and t in (0x71, 0x72):
            continue
        print(f"{c - t0:9d}  {'    ' * r}{role[r]} {tagname(t)}")
