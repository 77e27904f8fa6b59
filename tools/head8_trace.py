"""Prints the phase durations recorded by ISLE_HEAD8_TRACE (spmm_head_i8.cu): per super-chunk of CTA 0, clock stamps of the
TMA thread (slot 0: stage free), the MMA thread (1: operands ready, 2: MMAs issued), worker warp 0 of the owning group
(3: operands ready, 4: STTM issued, 5: stores complete, 6: drain done)."""
import sys
import numpy as np
launches, cur = [], []
for line in open(sys.argv[1]):
    if line.startswith("#"):
        if cur: launches.append(np.array(cur, dtype=np.int64))
        cur = []; print(line.strip()); continue
    cur.append([int(x) for x in line.split()])
if cur: launches.append(np.array(cur, dtype=np.int64))
for li, a in enumerate(launches[-2:]):
    t0 = a[a > 0].min()
    r = np.where(a > 0, a - t0, -1)
    print(f"launch {li}: n  tma_free  mma_ready mma_issued | w_ready w_sttm w_stdone w_drain | d(mma_ready) d(w_ready) expand stwait")
    for n in range(min(40, len(r))):
        dm = r[n, 1] - r[n - 1, 1] if n else 0
        dw = r[n, 3] - r[n - 2, 3] if n >= 2 else 0
        print(f"{n:3d} {r[n,0]:8d} {r[n,1]:8d} {r[n,2]:8d} | {r[n,3]:8d} {r[n,4]:8d} {r[n,5]:8d} {r[n,6]:8d} | {dm:6d} {dw:6d} {r[n,4]-r[n,3]:6d} {r[n,5]-r[n,4]:6d}")
    d = np.diff(r[8:80, 1])
    print("steady-state cycles per super-chunk (MMA ready to ready): mean %.0f median %.0f" % (d.mean(), np.median(d)))
