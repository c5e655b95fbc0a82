"""N>1 host logic on CPU: two gloo ranks shard the 101 consecutive pairs of a 102-frame sequence with
no overlap and no gap, and the reported time is the max over ranks."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from papteam_opticalflow_b200.shard import consecutive_pairs, pairs_for_rank


def test_shard_partition_properties():
    for n in (0, 1, 7, 101):
        for world in (1, 2, 4, 8):
            shards = [pairs_for_rank(n, r, world) for r in range(world)]
            flat = sorted(p for s in shards for p in s)
            assert flat == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert consecutive_pairs(4) == [(0, 1), (1, 2), (2, 3)] and consecutive_pairs(1) == []
    with pytest.raises(ValueError):
        pairs_for_rank(5, 2, 2)


@pytest.mark.slow
def test_two_gloo_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == 2 and r["npairs"] == 101 and r["slowest"] == 2.0
    flat = sorted(p for s in r["shards"] for p in s)
    assert flat == list(range(101)) and len(r["shards"][0]) == 51 and len(r["shards"][1]) == 50
