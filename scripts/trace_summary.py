"""Summarise gpurun_out/trace_rank*.npy (FS_DIST_TRACE=1): per kernel of one PCG iteration, mean time from kernel
start to 'flags seen', 'boundary rows sent', end, and the gap to the next kernel.  python scripts/trace_summary.py [dir]"""
import glob, os, sys
import numpy as np
d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
NAMES = {1: "A*p", 2: "xr", 3: "p-update"}
for f in sorted(glob.glob(os.path.join(d, "trace_rank*.npy"))):
    tr = np.load(f)
    tag, t = tr[:, 0].astype(np.int64), tr[:, 1].astype(np.int64)
    order = np.argsort(t, kind="stable")
    tag, t = tag[order], t[order]
    n = len(tag)
    tag, t = tag[n // 2:], t[n // 2:]            # the last (timed) pass
    kid = np.where(tag >= 100, tag // 10, tag // 10)      # kernel id: SELL tags are id*10+k; xr 20..23 -> 2; p 31..33 -> 3
    ev = tag % 10
    starts = np.nonzero((ev == 0) | ((kid == 3) & (ev == 1)))[0]      # kernel starts as seen by CTA 0
    rows = {}
    seq = []
    for a, b in zip(starts, list(starts[1:]) + [len(tag)]):
        k = int(kid[a])
        e = {int(ev[j]): int(t[j]) for j in range(a, b) if kid[j] == k}
        t0 = t[a]
        end = e.get(3, t0)
        nxt = t[b] if b < len(tag) else end
        last = e.get(6, end)
        rows.setdefault(k, []).append((e.get(1, t0) - t0, e.get(2, t0) - t0, end - t0, nxt - max(end, last), e.get(4, t0) - t0, e.get(5, t0) - t0, last - t0))
        seq.append(k)
    print(f"== {os.path.basename(f)}: {len(starts)} kernels traced")
    print(f"{'kernel':>14s} {'count':>6s} {'wait/ar us':>10s} {'bnd done':>10s} {'total':>10s} {'gap next':>10s} {'pre-fence':>10s} {'post-fence':>10s} {'last CTA':>10s}")
    tot = 0.0
    for k in sorted(rows):
        r = np.array(rows[k], dtype=np.float64) / 1e3
        name = NAMES.get(k, f"L{(k - 10) // 2} {'down' if (k - 10) % 2 == 0 else 'up'}" if k >= 10 else str(k))
        m = np.median(r, axis=0)
        print(f"{name:>14s} {len(r):6d} {m[0]:10.2f} {m[1]:10.2f} {m[2]:10.2f} {m[3]:10.2f} {m[4]:10.2f} {m[5]:10.2f} {m[6]:10.2f}")
        tot += max(m[2], m[6]) + m[3]
    print(f"sum of medians (kernel + gap): {tot:.1f} us")
