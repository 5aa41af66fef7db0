"""a1: the parity-check file format (src/alist.cpp:22-95): host loader of the C ABI and of the oracle
on the shipped matrices, plus the malformed-input cases the reference's loader silently mis-parses."""
import os

import numpy as np
import pytest

from ldpcsimulation_b200 import abi, capi
from oracle.oracle_api import Oracle, code_path

DIMS = {   # name: N, M, E, dv_max, dc_max   
    "PEG": (1008, 504, 3024, 3, 8),
    "802_3_H": (2048, 384, 12288, 6, 32),
    "802_3": (2048, 325, 10400, 6, 32),
    "4000": (4000, 2000, 16000, 4, 8),
    "4376": (4376, 282, 17504, 4, 63),
    "dvbs2": (64800, 32400, 226799, 8, 7),
}


@pytest.mark.parametrize("name", sorted(DIMS))
def test_shipped_codes_load(name):
    c = capi.Code(code_path(name))
    assert (c.N, c.M, c.E, c.dv_max, c.dc_max) == DIMS[name]
    if name != "dvbs2":
        o = Oracle(name)
        assert (o.N, o.M, o.E, o.dv_max, o.dc_max) == DIMS[name]


def _toy(padded=True, transpose_header=False, break_symmetry=False, short=False):
    # H = [[1 1 0 1],[0 1 1 1]]  N=4 M=2
    cols = [[1], [1, 2], [2], [1, 2]]
    rows = [[1, 2, 4], [2, 3, 4]]
    if break_symmetry:
        rows[0] = [1, 2, 3]
    L = ["2 4" if transpose_header else "4 2", "2 3", "1 2 1 2", "3 3"]
    for c in cols:
        L.append(" ".join(map(str, c + ([0] * (2 - len(c)) if padded else []))))
    for r in rows:
        L.append(" ".join(map(str, r)))
    if short:
        L = L[:-1]
    return "\n".join(L) + "\n"


def test_padded_and_unpadded_rows(tmp_path):
    for padded in (True, False):       # default loader needs padding, -DCPPSTYLE reads unpadded (src/alist.cpp:26-62)
        p = tmp_path / ("toy%d.alist" % padded)
        p.write_text(_toy(padded))
        c = capi.Code(str(p))
        assert (c.N, c.M, c.E, c.dv_max, c.dc_max) == (4, 2, 6, 2, 3)
        o = Oracle(str(p))
        assert (o.N, o.M, o.E) == (4, 2, 6)


def test_create_from_loadfile_arrays():
    # exactly loadFile()'s outputs: 1-based, zero padded
    num_n, num_m = [1, 2, 1, 2], [3, 3]
    nl = [1, 0, 1, 2, 2, 0, 1, 2]
    ml = [1, 2, 4, 2, 3, 4]
    c = capi.Code(arrays=(4, 2, 2, 3, num_n, num_m, nl, ml))
    assert c.E == 6


@pytest.mark.parametrize("kw,msg", [(dict(transpose_header=True), "transposed"), (dict(break_symmetry=True), "transpose"),
                                    (dict(short=True), "line count")])
def test_malformed_files_are_rejected(tmp_path, kw, msg):
    p = tmp_path / "bad.alist"
    p.write_text(_toy(**kw))
    with pytest.raises(capi.LdpcGpuError) as e:
        capi.Code(str(p))
    assert e.value.code == abi.ERR_BAD_CODE and msg in str(e.value)
    with pytest.raises(ValueError):
        Oracle(str(p))


def test_missing_file_is_an_error_not_a_crash(tmp_path):
    # the reference dereferences fopen()'s NULL (src/alist.cpp:71-74)
    with pytest.raises(capi.LdpcGpuError) as e:
        capi.Code(str(tmp_path / "nope.alist"))
    assert e.value.code == abi.ERR_IO
