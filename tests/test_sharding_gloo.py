"""(e) on the CPU: two ranks (gloo) each take their shard of the global frame-id range, sum their
counters with an all-reduce, and the result equals one run over the whole range.  The decode itself
is stood in for by the oracle here (no GPU in this test); the sharding and reduction code is the
product's (ldpcsimulation_b200/shard.py)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import torch, torch.distributed as dist
    import cases
    from ldpcsimulation_b200 import shard
    from oracle.oracle_api import Oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = cases.cfg_for("decodeSMNGDBF")
    orc = Oracle("PEG")
    begin, n = shard.shard_range(1000, 101, rank, world)
    r = orc.simulate(cfg, 4.0, 0.5, 5, begin, n)
    flat = torch.tensor(shard.pack_counters(r.counters, (r.iter_hist, r.error_weight_hist)), dtype=torch.int64)
    dist.all_reduce(flat)
    tot, (ith, ew) = shard.unpack_counters(flat.tolist(), (len(r.iter_hist), len(r.error_weight_hist)))
    if rank == 0:
        whole = orc.simulate(cfg, 4.0, 0.5, 5, 1000, 101)
        ok = tot == whole.counters and ith == whole.iter_hist.tolist() and ew == whole.error_weight_hist.tolist()
        print(json.dumps({"ok": ok, "tot": tot, "ranges": [shard.shard_range(1000, 101, g, world) for g in range(world)]}))
    dist.destroy_process_group()
''')


def test_two_rank_shards_sum_to_the_whole_run(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"], out
    assert out["ranges"] == [[1000, 50], [1050, 51]]
    assert out["tot"]["totalWords"] == 101


def test_shard_ranges_partition_exactly():
    from ldpcsimulation_b200 import shard
    for n in (0, 1, 7, 100, 10 ** 8 + 3):
        for world in (1, 2, 3, 8):
            parts = [shard.shard_range(5, n, g, world) for g in range(world)]
            assert sum(p[1] for p in parts) == n
            pos = 5
            for b, c in parts:
                assert b == pos
                pos += c
    seen = set()
    for s in range(4):
        for g in range(4):
            b, c = shard.step_range(s, g, 4, 10)
            assert not (set(range(b, b + c)) & seen)
            seen |= set(range(b, b + c))
    assert seen == set(range(160))
