"""Drop-in details of the reference's call surface: the TorchScript-ed RunningNorm, the top-level ``c_gae``
module, and the AMP history buffer step."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_close, assert_equal

DEV = "cuda:0"


def test_running_norm_scripts_on_a_cpu_only_box_and_keeps_the_reference_state_dict():
    """reference puffer_phc/policies/discriminator_policy.py:21 does ``torch.jit.script(RunningNorm(n))``."""
    from puffer_phc_b200.policies.running_norm import RunningNorm
    m = RunningNorm(934)
    sm = torch.jit.script(m)
    assert "phc_b200.rms_forward" in sm.code
    assert list(m.state_dict().keys()) == ["running_mean", "running_var", "count"]       # running_norm.py:9-11
    assert hasattr(sm, "update") and hasattr(sm, "finalize")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):          # no CPU kernel is registered: the dispatcher refuses host tensors
            sm(torch.zeros(2, 934))


def test_top_level_c_gae_import_resolves_to_the_drop_in():
    """reference puffer_phc/clean_pufferl/core.py:33-36: pyximport.install(...); from c_gae import compute_gae."""
    import subprocess
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import pyximport; pyximport.install(setup_args={'include_dirs': np.get_include()});"
            "import puffer_phc_b200; puffer_phc_b200.install_c_gae_shim(); from c_gae import compute_gae; print(compute_gae.__module__)" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "puffer_phc_b200.c_gae"


@pytest.mark.gpu
def test_scripted_running_norm_forward_and_update_match_the_eager_module(golden):
    from puffer_phc_b200.policies.running_norm import RunningNorm
    R = golden["rms"]                                   # the reference's module: update(cmu obs), update(synth obs[:200]), forward
    sm = torch.jit.script(RunningNorm(934).to(DEV))
    x1, x2 = torch.from_numpy(golden["cmu_step"]["obs"]).to(DEV), torch.from_numpy(golden["synth_step"]["obs"][:200]).to(DEV)
    sm.update(x1)
    sm.update(x2)
    assert_close(sm.running_mean.cpu().numpy(), R["mean2"], what="scripted running_mean")
    assert_close(sm.running_var.cpu().numpy(), R["var2"], what="scripted running_var", atol=1e-9)
    assert float(sm.count[0]) == float(R["count2"][0])
    sm.running_mean.copy_(torch.from_numpy(R["mean2"]))          # forward parity is stated for identical statistics
    sm.running_var.copy_(torch.from_numpy(R["var2"]))
    y = sm(torch.from_numpy(R["fwd_in"]).to(DEV))
    assert_close(y.cpu().numpy(), R["fwd_out"], rtol=1e-5, atol=1e-6, what="scripted forward vs the reference module")
    x2 = torch.from_numpy(R["fwd_in"]).to(DEV)
    # inside a scripted parent, as the reference's policy holds it
    class Policy(torch.nn.Module):
        def __init__(self, norm):
            super().__init__()
            self.obs_norm = norm

        def forward(self, x):
            return self.obs_norm(x) * 2.0
    p = torch.jit.script(Policy(sm))
    assert torch.equal(p(x2), y * 2.0)


@pytest.mark.gpu
def test_c_gae_shim_numpy_in_numpy_out(golden):
    import puffer_phc_b200
    d = puffer_phc_b200.install_c_gae_shim()
    try:
        sys.modules.pop("c_gae", None)
        from c_gae import compute_gae
        G = golden["gae"]
        for tag in ("a", "b", "c", "d", "e"):
            gm, lm = (float(v) for v in G[f"{tag}_gamma_lambda"])
            adv = compute_gae(G[f"{tag}_dones"], G[f"{tag}_values"], G[f"{tag}_rewards"], gm, lm)
            assert isinstance(adv, np.ndarray) and adv.dtype == np.float32
            assert_equal(adv.view(np.uint32), G[f"{tag}_adv"].view(np.uint32), f"c_gae shim vs the reference's compiled c_gae.pyx (golden {tag})")
    finally:
        sys.path.remove(d)
        sys.modules.pop("c_gae", None)


@pytest.mark.gpu
def test_amp_history_step_equals_shift_plus_compute(golden_amp):
    """amp_obs_history_step == _update_hist_amp_obs (humanoid_phc.py:1339-1348) + _compute_amp_observations (:1123-1174)."""
    from puffer_phc_b200.envs import common
    A = golden_amp
    t = lambda k: torch.from_numpy(np.ascontiguousarray(A[k])).to(DEV)   # noqa: E731
    args = [t(k) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_pos", "dof_vel", "key_pos", "shape", "limb", "subset")]
    for name in ("default", "full_dof", "with_params"):
        flags = [bool(x) for x in A[f"flags_{name}"]]
        want_row = torch.from_numpy(A[f"amp_{name}"]).to(DEV)
        N, W = want_row.shape
        S = 10
        buf = torch.rand(N, S, W, generator=torch.Generator().manual_seed(3)).to(DEV)
        old = buf.clone()
        common.amp_obs_history_step(buf, *args, *flags)
        assert torch.equal(buf[:, 1:], old[:, :-1]), f"{name}: history rows"
        assert_close(buf[:, 0].cpu().numpy(), want_row.cpu().numpy(), what=f"{name}: current row")
        got_row = common.build_amp_observations_smpl(*args, *flags)
        assert torch.equal(buf[:, 0], got_row), f"{name}: current row vs the stand-alone function"
