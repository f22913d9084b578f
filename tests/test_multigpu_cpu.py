"""N > 1 path on CPU: world_size-2 gloo run of the sharding + logits all-gather logic (hawq_b200.engine)."""
import os
import subprocess
import sys

import pytest
import torch

from hawq_b200.engine import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from hawq_b200.engine import shard_range, all_gather_logits
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = 10
g = torch.Generator().manual_seed(0)
full = torch.randn(total, 1000, generator=g)            # "logits" of the whole batch, same on every rank
lo, hi = shard_range(total, rank, world)
assert hi - lo == total // world
mine = full[lo:hi].clone()                               # what this rank's engine would have produced for its images
out = all_gather_logits(mine)
assert out.shape == (total, 1000)
assert torch.equal(out, full), "gathered logits differ from the unsharded batch"
# timing reduction used by bench.py: MAX over ranks
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t) == world
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_shard_range_partitions_every_image_once():
    for total in (1, 7, 8, 128, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_all_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("ok") >= 2
