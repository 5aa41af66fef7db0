#!/usr/bin/env python3
"""Import the reference's parity-check matrices and codeword files as DATA fixtures.

The hot path keeps the reference's file formats (alist under codes/, data.enc codewords), and
the tests / bench need the very matrices the reference ships because /root/reference is not
present on the GPU box.  Every alist is parsed, validated (nlist == mlist^T, ascending rows)
and re-emitted in the zero-padded layout the reference's default loader needs
(src/alist.cpp:71-91); codeword files are copied verbatim.  No source code is imported.
"""
import os
import shutil
import sys

REF = os.environ.get("LDPC_REFERENCE", "/root/reference/C_implementations")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "codes")

ALISTS = [
    "PEGReg504x1008/PEGReg504x1008.alist",
    "802_3/802_3_H.alist",
    "802_3/802_3.alist",
    "4000.2000.4.244/4000.2000.4.244.alist",
    "4376.282.4.9598/4376.282.4.9598.alist",
    "dvbs2_1_2/dvbs2_1_2.alist",
]
ENCS = ["PEGReg504x1008/data.enc", "4000.2000.4.244/data.enc"]


def parse(path):
    toks = [int(t) for t in open(path).read().split()]
    N, M, dv, dc = toks[:4]
    p = 4
    num_n = toks[p:p + N]; p += N
    num_m = toks[p:p + M]; p += M
    nl = [toks[p + i * dv: p + (i + 1) * dv] for i in range(N)]; p += N * dv
    ml = [toks[p + j * dc: p + (j + 1) * dc] for j in range(M)]; p += M * dc
    assert p == len(toks), "not a zero-padded alist"
    cols = [set() for _ in range(N)]
    for j in range(M):
        row = [v for v in ml[j] if v]
        assert len(row) == num_m[j] and row == sorted(row)
        for v in row:
            cols[v - 1].add(j + 1)
    for i in range(N):
        col = [v for v in nl[i] if v]
        assert len(col) == num_n[i] and col == sorted(col) and set(col) == cols[i], "nlist != mlist^T"
    return N, M, dv, dc, num_n, num_m, nl, ml


def emit(path, N, M, dv, dc, num_n, num_m, nl, ml):
    with open(path, "w") as f:
        f.write("%d %d\n%d %d\n" % (N, M, dv, dc))
        f.write(" ".join(map(str, num_n)) + "\n")
        f.write(" ".join(map(str, num_m)) + "\n")
        for r in nl:
            f.write(" ".join(map(str, r)) + "\n")
        for r in ml:
            f.write(" ".join(map(str, r)) + "\n")


def main():
    for rel in ALISTS:
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        emit(out, *parse(os.path.join(REF, "codes", rel)))
        print("alist", rel)
    for rel in ENCS:
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(REF, "codes", rel), out)
        os.chmod(out, 0o644)
        print("enc  ", rel)


if __name__ == "__main__":
    sys.exit(main())
