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


CASES = [(16, 12, 128, 0.1, 0), (64, 40, 2048, 0.02, 1), (7, 9, 40, 0.3, 2), (32, 8, 256, 0.0, 3), (1, 5, 3, 0.2, 5), (33, 3, 99, 0.0, 6),
         (1500, 6, 7001, 0.15, 7), (1024, 3, 10 ** 6, 0.5, 8)]


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,batch,p_mask,seed", CASES + [(40000, 4, 123457, 0.03, 9), (50, 150, 5000, 0.1, 10), (37, 70, 10 ** 6, 0.05, 11)])
def test_sorted_arrival_indices_match_reference(N, T, batch, p_mask, seed):
    """phc_rollout_store + phc_rollout_sort against the literal replay: arrival indices in (env, step) order, the sorted arrays, the
    running row count; batch_size cutting a step in half, N not a multiple of 32 / 1024, more than 1024 env groups, nothing cut."""
    v, r, d, m, B = make_case(N, T, batch, p_mask, seed)
    rv, rr, rd, idxs = reference_replay(v, r, d, m, B)
    dev = "cuda:0"
    buf = RolloutBuffer(N, B, max_steps=2, device=dev)                 # grows on demand
    for t in range(T):
        if t % 2:       # bool flags straight from the env adapter ...
            buf.store(torch.from_numpy(v[t]).to(dev), torch.from_numpy(r[t]).to(dev), torch.from_numpy(d[t] != 0).to(dev),
                      torch.zeros(N, dtype=torch.bool, device=dev), torch.from_numpy(m[t]).to(dev))
        else:           # ... or float vectors
            buf.store(torch.from_numpy(v[t]).to(dev), torch.from_numpy(r[t]).to(dev), torch.from_numpy(d[t]).to(dev), None,
                      torch.from_numpy(m[t]).to(dev))
    got = buf.sort_training_data().cpu().numpy()
    assert_equal(got, idxs, "sorted arrival indices")
    assert_equal(buf._sorted(buf.values).cpu().numpy(), rv[idxs], "values in sorted order (generic gather)")
    assert_equal(buf._sorted_buf["values"][: len(idxs)].cpu().numpy(), rv[idxs], "values in sorted order")
    assert_equal(buf._sorted_buf["dones"][: len(idxs)].cpu().numpy(), rd[idxs], "dones in sorted order")
    assert_equal(buf._sorted_buf["rewards"][: len(idxs)].cpu().numpy(), rr[idxs], "rewards in sorted order")
    assert buf.full == (m.sum() >= B)
    assert int(buf._stored) == int(m.sum())
    # a second rollout in the same buffer
    buf.reset()
    for t in range(T - 1, -1, -1):
        buf.store(torch.from_numpy(v[t]).to(dev), torch.from_numpy(r[t]).to(dev), torch.from_numpy(d[t]).to(dev), None, torch.from_numpy(m[t]).to(dev))
    rv2, rr2, rd2, idxs2 = reference_replay(v[::-1], r[::-1], d[::-1], m[::-1], B)
    assert_equal(buf.sort_training_data().cpu().numpy(), idxs2, "second rollout")
    assert_equal(buf._sorted_buf["rewards"][: len(idxs2)].cpu().numpy(), rr2[idxs2], "second rollout rewards")


def test_rollout_buffer_is_cuda_only():
    with pytest.raises(RuntimeError):
        RolloutBuffer(8, 16, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,batch,p_mask,seed", CASES[:4] + [(4096, 34, 131072, 0.01, 4)])
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


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,batch,p_mask", [(45, 5, 100, 0.2), (2050, 3, 5000, 0.0), (1024, 70, 80000, 0.1)])
def test_rollout_kernels_stay_inside_their_buffers(N, T, batch, p_mask):
    """compute-sanitizer is not available on the pool: call the two entry points directly on buffers with guard zones on both sides
    (rows, group counts, sorted outputs, scratch) and check that no guard word changed."""
    from puffer_phc_b200 import _ffi
    lib = _ffi.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + T)
    G = 64                                              # guard elements on each side
    groups = -(-N // 32)

    def guarded(n, dtype, fill):
        t = torch.full((n + 2 * G,), fill, dtype=dtype, device=dev)
        return t, t[G:G + n]

    bufs = {k: guarded(T * N, torch.float32, -7.0) for k in ("values", "rewards", "dones", "truncs")}
    bufs["mask"] = guarded(T * N, torch.uint8, 77)
    bufs["subrank"] = guarded(T * N, torch.uint8, 77)
    bufs["gcounts"] = guarded(T * groups, torch.int32, -7)
    bufs["rcounts"] = guarded(T, torch.int32, -7)
    bufs["rcounts"][1].zero_()
    stored = torch.zeros(1, dtype=torch.int64, device=dev)
    masks = torch.rand(T, N, generator=g) >= p_mask
    for t in range(T):
        v, r = torch.randn(N, generator=g).to(dev), torch.rand(N, generator=g).to(dev)
        d, m = (torch.rand(N, generator=g) < 0.1).to(dev), masks[t].to(dev)
        row = lambda k, w=N: _ffi.C.c_void_p(bufs[k][1].data_ptr() + t * w * bufs[k][1].element_size())
        _ffi.check(lib.phc_rollout_store(_ffi.ptr(v), _ffi.ptr(r), _ffi.ptr(d), None, 0, _ffi.ptr(m), N, row("values"), row("rewards"),
                                         row("dones"), row("truncs"), row("mask"), row("subrank"), row("gcounts", groups), row("rcounts", 1),
                                         _ffi.ptr(stored), _ffi.stream_ptr()), "store")
    rows = min(batch, int(masks.sum()))
    out = {k: guarded(rows, torch.float32, -7.0) for k in ("sd", "sv", "sr")}
    out["idxs"] = guarded(rows, torch.int64, -7)
    out["pos"] = guarded(rows, torch.int64, -7)
    nbytes = int(lib.phc_rollout_scratch_bytes(N, T))
    scratch = guarded((nbytes + 7) // 8, torch.int64, -7)
    meta = torch.zeros(4, dtype=torch.int64, device=dev)
    _ffi.check(lib.phc_rollout_sort(_ffi.ptr(bufs["dones"][1]), _ffi.ptr(bufs["values"][1]), _ffi.ptr(bufs["rewards"][1]), _ffi.ptr(bufs["mask"][1]),
                                    _ffi.ptr(bufs["subrank"][1]), _ffi.ptr(bufs["gcounts"][1]), _ffi.ptr(bufs["rcounts"][1]), N, T, batch,
                                    _ffi.ptr(scratch[1]), _ffi.ptr(meta), _ffi.ptr(out["sd"][1]), _ffi.ptr(out["sv"][1]), _ffi.ptr(out["sr"][1]),
                                    _ffi.ptr(out["idxs"][1]), _ffi.ptr(out["pos"][1]), _ffi.stream_ptr()), "sort")
    torch.cuda.synchronize()
    assert int(meta[2]) == rows and int(stored) == int(masks.sum())
    for name, (full, inner) in list(bufs.items()) + list(out.items()) + [("scratch", scratch)]:
        fill = full[0].item()
        assert bool((full[:G] == fill).all()) and bool((full[G + inner.numel():] == fill).all()), f"{name}: guard zone overwritten"
    assert bool((out["idxs"][1] >= 0).all()) and bool((out["idxs"][1] < rows).all()) and out["idxs"][1].unique().numel() == rows
