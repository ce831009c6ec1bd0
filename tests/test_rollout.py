"""Device-resident rollout buffer against a literal replay of the reference's Experience.store /
sort_training_data (reference puffer_phc/clean_pufferl/structs.py:113-145) and the GAE call site (core.py:213-259)."""
import numpy as np
import pytest
import torch

from conftest import assert_equal
from puffer_phc_b200.rollout import RolloutBuffer


def reference_replay(values, rewards, dones, masks, batch_size):
    """Experience.store + sort_training_data in plain Python (arrival-order rows, sort by (env_id, step))."""
    T, N = masks.shape
    rows_v, rows_r, rows_d, keys = [], [], [], []
    ptr = 0
    for step in range(T):
        if ptr >= batch_size:
            break
        indices = np.where(masks[step])[0][: batch_size - ptr]
        rows_v += list(values[step][indices]); rows_r += list(rewards[step][indices]); rows_d += list(dones[step][indices])
        keys += [(int(i), step) for i in indices]
        ptr += len(indices)
    idxs = np.asarray(sorted(range(len(keys)), key=keys.__getitem__), dtype=np.int64)
    return (np.asarray(rows_v, np.float32), np.asarray(rows_r, np.float32), np.asarray(rows_d, np.float32), idxs)


def make_case(N, T, batch, p_mask, seed):
    g = np.random.default_rng(seed)
    return (g.standard_normal((T, N)).astype(np.float32), g.random((T, N)).astype(np.float32),
            (g.random((T, N)) < 0.05).astype(np.float32), g.random((T, N)) >= p_mask, batch)


CASES = [(16, 12, 128, 0.1, 0), (64, 40, 2048, 0.02, 1), (7, 9, 40, 0.3, 2), (32, 8, 256, 0.0, 3)]


@pytest.mark.parametrize("N,T,batch,p_mask,seed", CASES)
def test_sorted_arrival_indices_match_reference(N, T, batch, p_mask, seed):
    v, r, d, m, B = make_case(N, T, batch, p_mask, seed)
    rv, rr, rd, idxs = reference_replay(v, r, d, m, B)
    buf = RolloutBuffer(N, B, device="cpu")
    for t in range(T):
        buf.store(torch.from_numpy(v[t]), torch.from_numpy(r[t]), torch.from_numpy(d[t]), torch.zeros(N), torch.from_numpy(m[t]))
    got = buf.sort_training_data().numpy()
    assert_equal(got, idxs, "sorted arrival indices")
    assert_equal(buf._sorted(buf.values).numpy(), rv[idxs], "values in sorted order")
    assert_equal(buf._sorted(buf.dones).numpy(), rd[idxs], "dones in sorted order")
    assert buf.full == (m.sum() >= B)


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,batch,p_mask,seed", CASES + [(4096, 34, 131072, 0.01, 4)])
def test_advantages_match_reference_flat_gae(N, T, batch, p_mask, seed):
    from oracle import c_oracle as co
    v, r, d, m, B = make_case(N, T, batch, p_mask, seed)
    rv, rr, rd, idxs = reference_replay(v, r, d, m, B)
    buf = RolloutBuffer(N, B, device="cuda:0")
    for t in range(T):
        buf.store(*(torch.from_numpy(x[t]).cuda() for x in (v, r, d)), torch.zeros(N, device="cuda:0"), torch.from_numpy(m[t]).cuda())
    adv, ret = buf.compute_advantages(0.98, 0.2)
    want = co.gae(rd[idxs], rv[idxs], rr[idxs], 0.98, 0.2)          # core.py:249 on the sorted arrays
    assert_equal(adv.cpu().numpy().view(np.uint32), want.view(np.uint32), "advantages (bit-exact)")
    assert_equal(ret.cpu().numpy().view(np.uint32), (want + rv[idxs]).view(np.uint32), "returns")
    assert_equal(buf.idxs.cpu().numpy(), idxs, "idxs")
