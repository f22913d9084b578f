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


EXIT_WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
bench.start_watchdog(60, "rank %%d" %% rank)               # armed like a real multi-rank run; must not fire
# the collective parity verdict of bench.py: every rank checks its shard, MIN over ranks, rank 0 reports
ok = torch.tensor([1], dtype=torch.int32)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"value": 1.0, "parity": {"bit_equal": bool(int(ok)), "ranks_checked": world}}))
# the way bench.py leaves a multi-rank run: final barrier, flush, os._exit(0) (no communicator teardown)
dist.barrier()
sys.stdout.flush(); sys.stderr.flush()
os._exit(0)
'''


def test_multi_rank_bench_exit_protocol(tmp_path):
    """bench.py ends a multi-rank run with barrier + flush + os._exit(0): under torch.distributed.run that is a clean exit (status
    0), rank 0's JSON line is on stdout, and the armed watchdog thread does not interfere."""
    script = tmp_path / "exit_worker.py"
    script.write_text(EXIT_WORKER % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and '"bit_equal": true' in lines[0] and '"ranks_checked": 2' in lines[0], r.stdout[-2000:]


def test_watchdog_ends_a_stuck_rank(tmp_path):
    """A rank that stops making progress is ended by bench.start_watchdog with exit status 3 and a message on stderr."""
    script = tmp_path / "stuck.py"
    script.write_text("import sys, time\nsys.path.insert(0, %r)\nimport bench\nbench.start_watchdog(1, 'rank 0 of 1')\ntime.sleep(30)\n" % ROOT)
    r = subprocess.run([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 3 and "watchdog" in r.stderr, (r.returncode, r.stderr[-500:])
