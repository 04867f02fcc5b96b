"""Env-index sharding (the N>1 path): pure host logic + a world_size-2 gloo run on CPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_oracle
from so100_mujoco_rl_b200.sharding import shard_range


@pytest.mark.parametrize("total,world", [(65536, 8), (1000, 3), (7, 8), (0, 2), (1 << 20, 8)])
def test_shards_partition_the_env_range(total, world):
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def test_bad_arguments():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        shard_range(-1, 0, 1)


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["SO100_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SO100_ROOT"], "tests"))
from conftest import make_oracle
from so100_mujoco_rl_b200.sharding import dist_env, max_over_ranks, shard_range, sum_over_ranks
rank, local, world = dist_env()
dist.init_process_group("gloo", rank=rank, world_size=world)
total = 12
lo, hi = shard_range(total, rank, world)
o = make_oracle(2, hi - lo, seed=21, env_offset=lo)          # this rank's shard, RNG keyed by the global env id
obs = o.reset()
rng = np.random.default_rng(0)
acts = rng.uniform(-1, 1, (5, total, 6)).astype(np.float32)   # same global action table on every rank
for t in range(5):
    obs, rew, *_ = o.step(acts[t, lo:hi])
buf = [torch.zeros(1) for _ in range(world)]
gathered = [None] * world
dist.all_gather_object(gathered, (lo, hi, obs, rew))
assert abs(max_over_ranks(float(rank + 1)) - world) < 1e-12
assert abs(sum_over_ranks(float(hi - lo)) - total) < 1e-12
if rank == 0:
    full = make_oracle(2, total, seed=21)
    fo = full.reset()
    for t in range(5):
        fo, fr, *_ = full.step(acts[t])
    for lo_, hi_, ob_, rw_ in gathered:
        assert np.array_equal(fo[lo_:hi_], ob_) and np.array_equal(fr[lo_:hi_], rw_)
    print("SHARDING_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_reproduce_the_single_process_run(tmp_path):
    """world_size 2 over gloo: each rank steps its env shard; the union equals the unsharded run bit for bit
    (no collective on the step path: torch.distributed only gathers results and agrees on timing)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SO100_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "SHARDING_OK" in r.stdout


PPO_WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["SO100_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SO100_ROOT"], "tests"))
from test_ppo import ToyEnv
from so100_mujoco_rl_b200.ppo import PPO, PPOConfig, pack_params
from so100_mujoco_rl_b200.sharding import dist_env
rank, local, world = dist_env()
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.set_num_threads(1)
algo = PPO(ToyEnv(32, limit=8, seed=100 + rank), PPOConfig(n_steps=8, n_epochs=2, n_minibatches=2, seed=5, cuda_graph=False))
p0 = pack_params(algo.policy).clone()
g = [torch.zeros_like(p0) for _ in range(world)]
dist.all_gather(g, p0)
assert all(torch.equal(g[0], x) for x in g), "ranks must start from identical weights"
algo.learn(total_samples=32 * 8 * world * 3, log_every=0, callback=lambda rec: None)
p1 = pack_params(algo.policy)
dist.all_gather(g, p1)
assert all(torch.equal(g[0], x) for x in g), "data-parallel update must keep the replicas identical"
assert not torch.equal(p0, p1)
assert algo.stats.samples == 32 * 8 * world * 3          # samples are counted over all ranks
if rank == 0:
    print("PPO_DP_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_data_parallel_ppo_keeps_replicas_identical(tmp_path):
    """world_size 2 over gloo: different env shards per rank, one flat gradient all-reduce per minibatch -> the two
    replicas stay bit-identical (the learner's only collective; the env step path has none)."""
    script = tmp_path / "ppo_worker.py"
    script.write_text(PPO_WORKER)
    env = dict(os.environ, SO100_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "PPO_DP_OK" in r.stdout
