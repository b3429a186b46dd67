"""Weight-ring round trip from a PNR_TRACE log with PNR_TRACE_STAGES=1: per stage, MMA-thread period, time from 'stage issued' to the
producer seeing the slot's W_EMPTY, producer reaction, and TMA issue -> W_FULL observed (only stages the MMA thread actually waited for)."""
import sys
from collections import defaultdict
import statistics as st

def load(path, launch=0):
    launches = []
    for line in open(path):
        if line.startswith("#"):
            launches.append(defaultdict(list)); continue
        r, c, t = line.split()
        launches[-1][int(r)].append((int(c), int(t)))
    return launches[launch]

for path in sys.argv[1:]:
    L = load(path)
    mma = L[0]
    ok = [c for c, t in mma if t == 0x72]
    issued = [c for c, t in mma if t == 0x71]
    n = min(len(ok), len(issued))
    tma = {}
    seen = {}
    for p in range(3):
        ev = L[5 + p]
        j = 0
        last_wait_end = None
        for c, t in ev:
            if 0x40 <= t < 0x70:
                last_wait_end = c
            elif t == 0x84:
                stage = p + 3 * j
                tma[stage] = c
                seen[stage] = last_wait_end
                j += 1
    # CTA 1's producers (roles 8..10), on CTA 0's clock
    off = L[2][0][0] - L[1][0][0]
    tma1, seen1 = {}, {}
    for p in range(3):
        j = 0
        last = None
        for c, t in L.get(8 + p, []):
            if 0x40 <= t < 0x70:
                last = c - off
            elif t == 0x84:
                tma1[p + 3 * j] = c - off
                seen1[p + 3 * j] = last
                j += 1
    lo, hi = 600, min(n, 3000)
    period = [ok[j + 1] - ok[j] for j in range(lo, hi - 1)]
    waited = [j for j in range(lo, hi) if j in tma and ok[j] - issued[j - 1] > 150]
    t_tma = [ok[j] - tma[j] for j in waited]
    t_rel = [seen[j + 5] - issued[j] for j in range(lo, hi - 5) if j + 5 in seen and seen[j + 5] is not None]
    t_react = [tma[j] - seen[j] for j in range(lo, hi) if j in tma and seen.get(j) is not None]
    wait = [ok[j] - issued[j - 1] for j in range(lo, hi)]
    lag_seen = [seen1[j] - seen[j] for j in range(lo, hi) if j in seen1 and seen1[j] is not None and seen.get(j) is not None]
    lag_tma = [tma1[j] - tma[j] for j in range(lo, hi) if j in tma1 and j in tma]
    q = lambda v: (round(st.median(v)), round(st.mean(v))) if v else None
    print(f"{path.split('/')[-1]:28s} stages {n}  period med/mean {q(period)}  wait-for-weights {q(wait)}  issued->producer sees W_EMPTY {q(t_rel)}  "
          f"CTA1 - CTA0: sees W_EMPTY {q(lag_seen) if lag_seen else None}, TMA issue {q(lag_tma) if lag_tma else None}  "
          f"producer reaction {q(t_react)}  TMA issue->W_FULL seen (waited stages: {len(waited)}) {q(t_tma)}")
