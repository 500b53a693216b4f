"""CPU: the N>1 path's host logic on gloo, world_size 2: the batch is partitioned by image with
no data-path collective; the only communication is the barrier and the max-over-ranks of the
timing / sum of the counts that bench.py does."""
import os
import socket
import subprocess
import sys

import pytest

from libmodjpeg_b200.batch import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1250, 10000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["MJ_ROOT"])
import torch, torch.distributed as dist
from libmodjpeg_b200.batch import shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1001, rank, world)
owned = torch.zeros(1001, dtype=torch.int64); owned[lo:hi] = 1
dist.all_reduce(owned)                      # every image owned exactly once
t = torch.tensor([1.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench.py's max-over-ranks
n = torch.tensor([hi - lo]); dist.all_reduce(n)
dist.barrier()
if rank == 0:
    print(json.dumps({"owned_once": bool((owned == 1).all()), "tmax": t.item(), "total": int(n.item()), "world": world}))
dist.destroy_process_group()
"""


def test_two_rank_gloo_partition(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MJ_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json

    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r == {"owned_once": True, "tmax": 2.0, "total": 1001, "world": 2}
