"""The N>1 path on CPU: two gloo ranks shard the envs, each forms its partial observation moments (here with the
C oracle standing in for the kernel), the packed fp64 buffer is all-reduced once, and every rank finalises to the
same RunningNorm state as a single process over all envs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_npz


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from puffer_phc_b200.dist import EpisodeMetrics, allreduce_packed, init_from_env, shard_range
    r, _, w = init_from_env(backend="gloo")
    obs = load_npz("cmu_step.npz")["obs"].astype(np.float64)
    lo, hi = shard_range(obs.shape[0], r, w)
    mine = obs[lo:hi]
    moments = torch.from_numpy(np.concatenate([[hi - lo], mine.sum(0), (mine ** 2).sum(0)]))
    metrics = EpisodeMetrics("cpu")
    S = load_npz("cmu_step.npz")
    metrics.add(torch.from_numpy(S["reward"][lo:hi]), torch.from_numpy(S["reward_raw"][lo:hi]),
                torch.from_numpy(S["reset_train"][lo:hi]), torch.from_numpy(S["terminated_train"][lo:hi]))
    allreduce_packed([moments, metrics.buf])
    q.put((r, moments.numpy().copy(), metrics.buf.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_moment_allreduce_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    S = load_npz("cmu_step.npz")
    obs = S["obs"].astype(np.float64)
    want = np.concatenate([[obs.shape[0]], obs.sum(0), (obs ** 2).sum(0)])
    for _, m, met in res:
        np.testing.assert_allclose(m, want, rtol=1e-13, atol=1e-13)
        assert met[0] == obs.shape[0]
        np.testing.assert_allclose(met[1], S["reward"].astype(np.float64).sum(), rtol=1e-12)
        assert met[7] == S["reset_train"].sum() and met[8] == S["terminated_train"].sum()
    np.testing.assert_array_equal(res[0][1], res[1][1])
    # finalising the reduced moments reproduces the reference's RunningNorm.update on the full batch
    R = load_npz("rms.npz")
    n = want[0]
    mean = want[1:935] / n
    var = want[935:] / n - mean ** 2
    np.testing.assert_allclose(mean.astype(np.float32), R["mean1"][0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(var.astype(np.float32), R["var1"][0], rtol=1e-5, atol=1e-9)
