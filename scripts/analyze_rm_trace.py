"""per-level timing of the row-major warp-per-row sweep from the 4-stamps-per-row trace (NGSAMG_B200_TRACE_FILE)"""
import sys
import numpy as np
fn = sys.argv[1]
raw = np.fromfile(fn, dtype=np.int64)
nr, nl = int(raw[0]), int(raw[1])
ls = raw[2:2 + nl]
t = raw[2 + nl:].reshape(nr, 12).astype(np.float64)
bwd = fn.endswith("bwd")
t0 = t[t[:, 0] > 0, 0].min()
t = (t - t0) / 1e3   # us
print("rows", nr, "levels", nl - 1, "total %.1f us" % t[:, 3].max())
prev_pub = 0.0
order = range(nl - 1) if not bwd else range(nl - 2, -1, -1)
rows = []
for k, lv in enumerate(order):
    a, b = int(ls[lv]), int(ls[lv + 1])
    if b <= a: continue
    seg = t[a:b]
    crit = seg[:, 3].argmax()
    pub = seg[crit, 3]
    rows.append((lv, b - a, pub - prev_pub, seg[crit, 0] - prev_pub, seg[crit, 1] - seg[crit, 0], seg[crit, 2] - max(seg[crit, 1], prev_pub), seg[crit, 3] - seg[crit, 2]))
    prev_pub = pub
rows = np.array(rows)
print("level rows hop  pickup-prevpub  stage-wait  poll-after-ready  epilogue   [us, critical row of the level]")
for r in rows[:12]: print("%4d %5d %6.2f %8.2f %8.2f %8.2f %8.2f" % tuple(r))
print("...")
print("mean hop %.2f us; mean: pickup-prevpub %.2f stage-wait %.2f poll-after-ready %.2f epilogue %.2f" % tuple(rows[3:, 2:].mean(axis=0)))

# fine stamps of the critical rows: 0 top, 4 after wait_all, 5 after lds ptr, 6 after stage_row, 1 after stage_ptr, 7 after gate, 2 polls done, 8 butterfly done, 9 scalars read, 3 stored
seq = [0, 4, 5, 6, 1, 7, 2, 8, 9, 3]
names = ["wait_all", "lds_ptr", "stage_row", "stage_ptr", "gate", "polls", "butterfly", "lds_scal", "store"]
acc = np.zeros(len(names)); cnt = 0
for k, lv in enumerate(order):
    a, b = int(ls[lv]), int(ls[lv + 1])
    if b <= a or k < 3: continue
    seg = t[a:b]; crit = seg[:, 3].argmax()
    acc += np.diff(seg[crit, seq]); cnt += 1
print("critical rows, mean us per section:", "  ".join("%s %.2f" % (n, v) for n, v in zip(names, acc / max(cnt, 1))))
