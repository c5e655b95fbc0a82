"""world_size-2 worker (gloo, CPU) for tests/test_multi_gloo.py."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papteam_opticalflow_b200.shard import consecutive_pairs, max_over_ranks, pairs_for_rank

dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pairs = consecutive_pairs(102)                   # the 1920 collection has 102 frames -> 101 pairs
mine = pairs_for_rank(len(pairs), rank, world)
fake_seconds = 1.0 + rank                        # rank 1 is the slow one
slowest = max_over_ranks(dist, fake_seconds)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "npairs": len(pairs), "shards": gathered, "slowest": slowest}))
dist.destroy_process_group()
