"""The torch CPU restatement used as the timed CPU baseline (oracle/torch_port.py) against the golden vectors."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_equal
from oracle import torch_port as tp


@pytest.mark.parametrize("tn,sn", [("cmu_tables", "cmu_step"), ("synth_tables", "synth_step")])
def test_port_step_vs_reference(golden, tn, sn):
    T = {k: torch.from_numpy(v) for k, v in golden[tn].items()}
    G = golden[sn]
    S = {k[3:]: torch.from_numpy(v) for k, v in G.items() if k.startswith("in_")}
    out = tp.step(T, S)
    assert_close(out["obs"].numpy(), G["obs"], what="obs", row_scale=True)
    assert_close(out["reward"].numpy(), G["reward"], what="reward")
    assert_close(out["reward_raw"].numpy(), G["reward_raw"], what="reward_raw")
    assert_equal(out["reset"].numpy(), G["reset_train"], "reset")
    assert_equal(out["terminated"].numpy(), G["terminated_train"], "terminated")
    ms = tp.motion_state(T, S["motion_ids"], torch.from_numpy(G["t1"]), S["global_offset"])
    for k in ("dof_pos", "dof_vel", "rb_rot", "rg_pos", "motion_aa"):
        assert_close(ms[k].numpy(), G[f"t1_{k}"], what=k)


def test_port_rms(golden):
    R, x1 = golden["rms"], torch.from_numpy(golden["cmu_step"]["obs"])
    m, v, c = tp.rms_update(x1, torch.zeros(1, 934), torch.ones(1, 934), torch.ones(1))
    assert_close(m.numpy(), R["mean1"], what="mean")
    assert_close(v.numpy(), R["var1"], rtol=1e-5, atol=1e-9, what="var")
    y = tp.rms_forward(torch.from_numpy(R["fwd_in"]), torch.from_numpy(R["mean2"]), torch.from_numpy(R["var2"]))
    assert_equal(y.numpy().view(np.uint32), R["fwd_out"].view(np.uint32), "forward bits")
