#!/usr/bin/env python3
"""Generate a regular (dv, dc) non-binary alist over GF(q) (format of SystemC/NB-LDPC/src/alist.cpp:23-56: header `N M q`, then
(index value) pairs).  The reference ships GF(2) / GF(4) / GF(8) codes only; BASELINE.json configs[4] names GF(16), so the GF(16)
fixture codes/NB/gf16.reg.1536.768.alist is made here (seeded, socket construction with 4-cycle avoidance by re-drawing).
Usage: python tools/make_nb_code.py N dv dc q seed out"""
import sys

import numpy as np


def make(N, dv, dc, q, seed):
    M = N * dv // dc
    rng = np.random.default_rng(seed)
    for attempt in range(200):
        sockets = np.repeat(np.arange(M), dc)
        rng.shuffle(sockets)
        rows = sockets.reshape(N, dv)
        if all(len(set(r)) == dv for r in rows):
            # no two columns may share two checks (no 4-cycles)
            seen, ok = set(), True
            for r in rows:
                for a in range(dv):
                    for b in range(a + 1, dv):
                        k = (min(r[a], r[b]), max(r[a], r[b]))
                        if k in seen:
                            ok = False
                        seen.add(k)
            if ok:
                break
    else:
        raise SystemExit("no simple graph found")
    rows = np.sort(rows, axis=1)
    vals = rng.integers(1, q, size=(N, dv))
    mrows = [[] for _ in range(M)]
    for i in range(N):
        for s in range(dv):
            mrows[rows[i, s]].append((i, vals[i, s]))
    return M, rows, vals, mrows


def main():
    N, dv, dc, q, seed, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    M, rows, vals, mrows = make(N, dv, dc, q, seed)
    with open(out, "w") as f:
        f.write("%d %d %d\n%d %d\n" % (N, M, q, dv, dc))
        f.write(" ".join([str(dv)] * N) + "\n" + " ".join([str(dc)] * M) + "\n")
        for i in range(N):
            f.write(" \t".join("%d %d" % (rows[i, s] + 1, vals[i, s]) for s in range(dv)) + " \t\n")
        for j in range(M):
            f.write(" \t".join("%d %d" % (i + 1, v) for i, v in sorted(mrows[j])) + " \t\n")


if __name__ == "__main__":
    main()
