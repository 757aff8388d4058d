"""CPU tier: the N>1 host logic (partition -> per-rank tracking -> ordered gather) with 2 gloo
processes.  The per-rank compute is the CPU oracle here (tests may use it); on a GPU box the same
harness drives the CUDA path (bench.py --gpus N, tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_seq, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import klt_oracle as O
    from visual_odom_pipeline_b200 import sharding, synth as S
    r, lr, w = sharding.init_process_group("gloo")
    assert (r, w) == (rank, world)
    lo, hi = sharding.shard_range(n_seq, world, rank)
    res = []
    for s in range(lo, hi):   # sequence s = its own seed, like BASELINE configs[3]
        a, b = S.frame_pair(60, 80, seed=100 + s)
        p = S.uniform_points(16, 60, 80, seed=s)
        q, st, er = O.calc_optical_flow_pyr_lk(a, b, p, None, (9, 9), 2, (3, 30, 0.01))
        res.append(np.concatenate([q.reshape(-1, 2), st.astype(np.float32), er], 1))
    local = torch.from_numpy(np.stack(res)) if res else torch.zeros((0, 16, 4))
    sharding.barrier()
    t = sharding.max_over_ranks(float(rank + 1))
    assert t == float(world)
    assert sharding.sum_over_ranks(hi - lo) == n_seq
    full = sharding.gather_shards(local, n_seq)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_seq", [5, 2])
def test_two_rank_sharding_matches_unsharded(tmp_path, oracle, n_seq):
    from visual_odom_pipeline_b200 import synth as S
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_seq, out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == (n_seq, 16, 4)
    for s in range(n_seq):
        a, b = S.frame_pair(60, 80, seed=100 + s)
        p = S.uniform_points(16, 60, 80, seed=s)
        q, st, er = oracle.calc_optical_flow_pyr_lk(a, b, p, None, (9, 9), 2, (3, 30, 0.01))
        want = np.concatenate([q.reshape(-1, 2), st.astype(np.float32), er], 1)
        assert np.array_equal(got[s].view(np.uint32), want.view(np.uint32)), s
