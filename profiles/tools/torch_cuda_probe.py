"""How does torch-CUDA round the flag-deciding reductions of the hot path?  (hard part 1.)

Runs on the GPU box; dumps raw inputs / outputs so that the candidate formulas can be fitted offline:
  * torch.norm(x, dim=-1) over 3 components (compute_humanoid_im_reset, common.py:343-350)
  * .mean(dim=-1) over 20 / 24 distances (the eval variant, common.py:342-346)
  * tensor / python-scalar (sample_time_interval, motion_lib.py:533)
and then the whole reference step (oracle/_ref, the reference's own files) on torch-CPU and torch-CUDA against the kernels.

    python profiles/tools/torch_cuda_probe.py gpurun_out/probe
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(prefix):
    dev = "cuda:0"
    g = torch.Generator().manual_seed(7)
    out = {}
    x = torch.randn(1 << 18, 3, generator=g) * 0.2
    out["norm_in"] = x.numpy()
    out["norm_cuda"] = torch.norm(x.to(dev), dim=-1).cpu().numpy()
    out["norm_cpu"] = torch.norm(x, dim=-1).numpy()
    x24 = (torch.randn(1 << 13, 24, 3, generator=g) * 0.2)
    out["norm24_in"] = x24.numpy()
    out["norm24_cuda"] = torch.norm(x24.to(dev), dim=-1).cpu().numpy()
    for n in (20, 24):
        d = torch.rand(1 << 15, n, generator=g)
        out[f"mean{n}_in"] = d.numpy()
        out[f"mean{n}_cuda"] = d.to(dev).mean(dim=-1).cpu().numpy()
        out[f"mean{n}_cpu"] = d.mean(dim=-1).numpy()
    p = torch.rand(1 << 18, generator=g) * 9.7
    out["div_in"] = p.numpy()
    out["div_cuda"] = (p.to(dev) / (1 / 30)).cpu().numpy()
    out["div_cpu"] = (p / (1 / 30)).numpy()
    s4 = torch.randn(1 << 16, 4, generator=g)
    out["sum4_in"] = s4.numpy()
    out["sum4_cuda"] = s4.to(dev).sum(dim=-1).cpu().numpy()          # slerp's dot product: torch.sum(q0*q1, dim=-1)
    # the same reductions the way the env calls them: operands are strided views of a [N,24,13] AoS buffer, inside TorchScript
    from oracle import ref_runner as rr0
    R0 = rr0.boot()
    aos = torch.randn(1 << 13, 24, 13, generator=g) * 0.2
    refp = torch.randn(1 << 13, 24, 3, generator=g) * 0.2
    prog = torch.full((1 << 13,), 5, dtype=torch.int16)
    pt = torch.zeros(1 << 13, dtype=torch.bool)
    td = torch.full((24,), 0.25)
    rb = torch.ones(1 << 13, dtype=torch.bool)
    cf = torch.zeros(1 << 13, 24, 3)
    ci = torch.zeros(4, dtype=torch.long)
    out["jit_aos"], out["jit_ref"] = aos.numpy(), refp.numpy()
    for name, dv in (("cpu", "cpu"), ("cuda", dev)):
        a_, r_ = aos.to(dv), refp.to(dv)
        for _ in range(3):
            rs, tm = R0.common.compute_humanoid_im_reset(rb.to(dv), prog.to(dv), cf.to(dv), ci.to(dv), a_[..., 0:3], r_, pt.to(dv), True, td.to(dv), False)
        out[f"jit_term_{name}"] = tm.cpu().numpy()
        out[f"jit_dist_{name}"] = torch.norm(a_[..., 0:3] - r_, dim=-1).cpu().numpy()
    np.savez_compressed(prefix + "_reductions.npz", **out)

    # ---- the reference's own step on CPU and CUDA vs the kernels --------------------------------------------
    from oracle import ref_runner as rr
    from puffer_phc_b200 import synth
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T = synth.make_motion_library(2000, seed=0, device="cpu", other_fps_fraction=0.2)
    N = 16384
    S = synth.make_env_state(T, N, seed=1)
    lib_c = rr.lib_from_tables(T, "cpu")
    lib_g = rr.lib_from_tables(T, dev)
    Sg = {k: v.to(dev) for k, v in S.items()}
    ref_c = rr.step(lib_c, S, with_blend=True)
    for _ in range(3):                           # TorchScript profiling executor: let it specialise / fuse before the kept run
        ref_g = rr.step(lib_g, Sg, with_blend=True)
    mg = rr.flag_margins(lib_g, Sg)
    mc = rr.flag_margins(lib_c, S)
    ours_lib = MotionLibSMPL.from_tables({k: v.to(dev) for k, v in T.items()}, device=dev)
    fs = FusedStep(ours_lib, N, StepConfig())
    o = fs(Sg["body_state"], Sg["progress"], Sg["start_time"], Sg["start_offset"], Sg["motion_ids"], Sg["global_offset"], Sg["dof_force"], Sg["dof_vel"])
    torch.cuda.synchronize()
    rep = {}
    for k in ("reset", "terminated"):
        a, c, gq = o[k].cpu().numpy().astype(bool), ref_c[k].numpy(), ref_g[k].cpu().numpy()
        rep[k] = {"ours_vs_ref_cpu": int((a != c).sum()), "ours_vs_ref_cuda": int((a != gq).sum()), "ref_cpu_vs_ref_cuda": int((c != gq).sum())}
    for k in ("t0_idx0", "t0_idx1", "t1_idx0", "t1_idx1"):
        rep[k] = {"ref_cpu_vs_ref_cuda": int((ref_c[k].numpy() != ref_g[k].cpu().numpy()).sum())}
    for k in ("t0_blend", "t1_blend"):
        rep[k] = {"ref_cpu_vs_ref_cuda_bits": int((ref_c[k].numpy().view(np.uint32) != ref_g[k].cpu().numpy().view(np.uint32)).sum())}
    for k in ("obs", "reward", "reward_raw"):
        a, c, gq = o[k].cpu().numpy().astype(np.float64), ref_c[k].numpy().astype(np.float64), ref_g[k].cpu().numpy().astype(np.float64)
        f = lambda u, v: float((np.abs(u - v) / (1e-5 * np.abs(v) + 2e-6)).max())      # noqa: E731
        rep[k] = {"ours_vs_ref_cpu_err_over_tol": f(a, c), "ours_vs_ref_cuda_err_over_tol": f(a, gq), "ref_cuda_vs_ref_cpu_err_over_tol": f(gq, c)}
    BL = (("root_h", 0, 1), ("self_pos", 1, 70), ("self_rot", 70, 214), ("self_vel", 214, 286), ("self_ang", 286, 358), ("d_pos", 358, 430),
          ("d_rot", 430, 574), ("d_vel", 574, 646), ("d_ang", 646, 718), ("l_pos", 718, 790), ("l_rot", 790, 934))
    oc, og, oo = ref_c["obs"].numpy().astype(np.float64), ref_g["obs"].cpu().numpy().astype(np.float64), o["obs"].cpu().numpy().astype(np.float64)
    rep["obs_blocks"] = {}
    for name, a0, a1 in BL:
        f = lambda u, v: (np.abs(u - v) / (1e-5 * np.abs(v) + 2e-6))[:, a0:a1]      # noqa: E731
        e_gc, e_oc, e_og = f(og, oc), f(oo, oc), f(oo, og)
        rep["obs_blocks"][name] = {"cuda_vs_cpu_max": float(e_gc.max()), "cuda_vs_cpu_n_over": int((e_gc > 1).sum()),
                                   "ours_vs_cpu_max": float(e_oc.max()), "ours_vs_cpu_n_over": int((e_oc > 1).sum()),
                                   "ours_vs_cuda_max": float(e_og.max()), "ours_vs_cuda_n_over": int((e_og > 1).sum()), "n": int(e_gc.size)}
    # envs where torch-CUDA and torch-CPU disagree: is it the slerp fall-back knife edge (|sin| < 0.001, torch_utils.py:128)?
    bad = np.argwhere((np.abs(og - oc) / (1e-5 * np.abs(oc) + 2e-6))[:, 790:934] > 3)
    rows = sorted(set(int(b[0]) for b in bad))[:64]
    detail = []
    full_c = rr.step(lib_c, S, full_state=True, with_blend=True)
    for e in rows:
        cols = [int(b[1]) for b in bad if b[0] == e]
        bodies = sorted(set(c // 6 for c in cols))
        idm = int(S["motion_ids"][e])
        f0 = int(full_c["t1_idx0"][e]) + int(T["length_starts"][idm]); f1 = int(full_c["t1_idx1"][e]) + int(T["length_starts"][idm])
        for j in bodies[:4]:
            q0, q1 = T["grs"][f0, j], T["grs"][f1, j]
            pr = (q0 * q1)
            c_seq = float(((pr[0] + pr[1]) + pr[2]) + pr[3]); c_tree = float((pr[0] + pr[2]) + (pr[1] + pr[3])); c_tree2 = float((pr[0] + pr[1]) + (pr[2] + pr[3]))
            c_cuda = float(torch.sum((q0.to(dev) * q1.to(dev)), dim=-1).cpu())
            c_cpu = float(torch.sum(q0 * q1, dim=-1))
            detail.append({"env": e, "body": j, "blend": float(full_c["t1_blend"][e]), "c_seq": c_seq, "c_tree_02_13": c_tree, "c_tree_01_23": c_tree2,
                           "c_torch_cuda": c_cuda, "c_torch_cpu": c_cpu, "s_seq": float(np.sqrt(np.float32(1) - np.float32(c_seq) * np.float32(c_seq))),
                           "s_cuda": float(np.sqrt(np.float32(1) - np.float32(c_cuda) * np.float32(c_cuda)))})
    rep["slerp_knife_edge_rows"] = detail[:48]
    rep["rows_cuda_vs_cpu_over_3tol_in_l_rot"] = len(rows)
    rep["dist_bits_ref_cpu_vs_ref_cuda"] = int((mg["dist"].cpu().numpy().view(np.uint32) != mc["dist"].numpy().view(np.uint32)).sum())
    rep["dist_total"] = int(mc["dist"].numel())
    print(json.dumps(rep, indent=1))
    json.dump(rep, open(prefix + "_step.json", "w"), indent=1)
    np.savez_compressed(prefix + "_step.npz", dist_cuda=mg["dist"].cpu().numpy(), dist_cpu=mc["dist"].numpy(),
                        body_pos=S["body_state"][:, :24, 0:3].numpy(),
                        rg_pos_cuda=rr.step(lib_g, Sg, full_state=True)["t0_rg_pos"].cpu().numpy(),
                        rg_pos_cpu=rr.step(lib_c, S, full_state=True)["t0_rg_pos"].numpy(),
                        reset_cuda=ref_g["reset"].cpu().numpy(), reset_cpu=ref_c["reset"].numpy(), reset_ours=o["reset"].cpu().numpy(),
                        term_cuda=ref_g["terminated"].cpu().numpy(), term_cpu=ref_c["terminated"].numpy(), term_ours=o["terminated"].cpu().numpy())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/probe")
