"""The whole post-physics half of ``HumanoidPHC.step`` as ONE kernel launch.

Reference flow (puffer_phc/envs/humanoid_phc.py:136-149): ``_compute_reward`` (:1228-1303) ->
``_compute_reset`` (:1311-1333) -> ``_compute_observations`` (:935-959), with two ``get_motion_state`` queries
(t and t+1) behind them and ``RunningNorm.forward`` applied later by the policy.  ``FusedStep`` takes the same
per-env buffers the env holds (PhysX rigid-body tensor, ``progress_buf``, ``_motion_start_times``,
``_motion_start_times_offset``, ``_sampled_motion_ids``, ``_global_offset``, ``dof_force_tensor``, ``_dof_vel``)
and fills ``obs_buf``, ``rew_buf``, ``reward_raw``, ``reset_buf``, ``_terminate_buf`` -- optionally also the
RMS-normalised observation and the fp64 column moments ``RunningNorm.update`` needs -- in one pass over HBM
(csrc/step_fused.cu).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import torch

from . import _ffi
from .motion_lib import MotionLibBase
from .policies.running_norm import RunningNorm

NUM_BODIES, OBS_DIM = 24, 934
EVAL_BODY_IDS = tuple(j for j in range(24) if j not in (4, 8, 18, 23))   # body_sets.py:42,57


@dataclass
class StepConfig:
    """Kernel constants; defaults = the reference's config.py defaults."""
    dt: float = 1.0 / 30.0                      # isaacgym_env.py:39-41
    k_pos: float = 100.0                        # RewardConfig, config.py:25-32
    k_rot: float = 10.0
    k_vel: float = 0.1
    k_ang_vel: float = 0.1
    w_pos: float = 0.5
    w_rot: float = 0.3
    w_vel: float = 0.1
    w_ang_vel: float = 0.1
    use_power_reward: bool = True               # config.py:37
    rew_power_coef: float = 0.0005              # config.py:96
    enable_early_termination: bool = True       # config.py:84
    termination_distance: float = 0.25          # config.py:85
    reset_bodies: Sequence[int] = field(default_factory=lambda: tuple(range(24)))
    use_mean: bool = False                      # flag_im_eval (humanoid_phc.py:1332)
    ref_device: Optional[str] = None            # "cuda" / "cpu": whose torch rounding the flag-deciding reductions reproduce
                                                # (include/phc_b200.h PHC_REF_DEVICE_*); None = the package setting (default cuda)

    def eval_mode(self) -> "StepConfig":
        """toggle_eval_mode (humanoid_phc.py:1421-1438): 0.5 m, mean over the 20 eval bodies."""
        import dataclasses
        return dataclasses.replace(self, termination_distance=0.5, reset_bodies=EVAL_BODY_IDS, use_mean=True)


class HostStepHandle:
    """A step queued by ``FusedStep.step_host(..., wait=False)``: ``result()`` blocks until its results are in host memory."""

    def __init__(self, out: Dict[str, torch.Tensor], done: "torch.cuda.Event"):
        self._out, self._done = out, done

    def result(self) -> Dict[str, torch.Tensor]:
        self._done.synchronize()
        return self._out


class FusedStep:
    def __init__(self, motion_lib: MotionLibBase, num_envs: int, cfg: Optional[StepConfig] = None,
                 rms: Optional[RunningNorm] = None, normalize: bool = False, accumulate_moments: bool = False,
                 debug_ref: bool = False, defer_moments: bool = False, metrics: bool = False):
        """``accumulate_moments``: the kernel also produces the fp64 column sums ``RunningNorm.update`` needs.  By default they
        are folded into ``rms``' pending moments after every step (one tiny reduce launch).  ``defer_moments=True`` lets the
        kernel ADD every step's sums to its per-CTA slots instead; call ``flush_moments()`` once per rollout (before
        ``rms.finalize()``) -- no per-step reduce launch at all.

        ``metrics``: the kernel also sums the episode metrics the reference logs (reference puffer_phc/clean_pufferl/env.py:102-110:
        env-steps, reward, the five reward_raw columns, resets, terminations) into per-CTA fp64 slots; ``flush_moments()`` folds
        them into ``self.stats[1 + 2 * 934:]`` (``metric_values()``), the tail of the ONE buffer ``rms.finalize()`` all-reduces."""
        self.lib = _ffi.load()
        self.motion_lib = motion_lib
        self.cfg = cfg or StepConfig()
        self.N = int(num_envs)
        dev = motion_lib._device
        self.device = dev
        self.rms = rms
        self.normalize = bool(normalize)
        self.accumulate_moments = bool(accumulate_moments)
        self.defer_moments = bool(defer_moments) and (self.accumulate_moments or bool(metrics))
        self._pending_rows = 0
        if (normalize or accumulate_moments) and rms is None:
            raise ValueError("normalize / accumulate_moments need a RunningNorm")
        c = self.cfg
        self.raw_dim = 5 if c.use_power_reward else 4
        # the env's output buffers (humanoid_phc.py:554-575)
        self.obs_buf = torch.zeros((self.N, OBS_DIM), dtype=torch.float32, device=dev)
        self.obs_norm = torch.zeros((self.N, OBS_DIM), dtype=torch.float32, device=dev) if normalize else None
        self.rew_buf = torch.zeros(self.N, dtype=torch.float32, device=dev)
        self.reward_raw = torch.zeros((self.N, self.raw_dim), dtype=torch.float32, device=dev)
        self.reset_buf = torch.ones(self.N, dtype=torch.bool, device=dev)
        self.terminate_buf = torch.ones(self.N, dtype=torch.bool, device=dev)
        self.termination_distances = torch.full((NUM_BODIES,), float(c.termination_distance), dtype=torch.float32, device=dev)
        self.num_partials = int(self.lib.phc_step_num_partials())
        # one slot per CTA of the step kernel, then one per CTA of the auto-reset tail (its moment corrections, envs/reset.py)
        self.num_tail_partials = int(self.lib.phc_auto_reset_num_partials())
        self.partials = (torch.zeros((self.num_partials + self.num_tail_partials, 2, OBS_DIM), dtype=torch.float64, device=dev)
                         if accumulate_moments else None)
        self.metrics = bool(metrics)
        self.metric_partials = torch.zeros((self.num_partials, _ffi.NUM_METRICS), dtype=torch.float64, device=dev) if metrics else None
        self.row_adjust = torch.zeros(1, dtype=torch.float64, device=dev)
        self.tail_slots_used = False     # set by envs.reset.AutoReset: the tail's correction slots then take part in the folds
        # the rank's statistics buffer: [n, sum x, sum x^2 | episode metrics] -- moments and metrics travel in one all-reduce
        self.stats = torch.zeros(1 + 2 * OBS_DIM + _ffi.NUM_METRICS, dtype=torch.float64, device=dev)
        if rms is not None and (accumulate_moments or metrics):
            rms.attach_stats(self.stats)
        self.ref_t = torch.zeros((self.N, 312), dtype=torch.float32, device=dev) if debug_ref else None
        self.ref_t1 = torch.zeros((self.N, 312), dtype=torch.float32, device=dev) if debug_ref else None
        mask = 0
        for j in c.reset_bodies:
            mask |= 1 << int(j)
        self._ccfg = _ffi.StepCfg(
            float(torch.tensor(c.dt, dtype=torch.float32)),
            (C.c_float * 4)(c.k_pos, c.k_rot, c.k_vel, c.k_ang_vel), (C.c_float * 4)(c.w_pos, c.w_rot, c.w_vel, c.w_ang_vel),
            float(c.rew_power_coef), mask, int(c.enable_early_termination), int(c.use_mean),
            float(rms.epsilon) if rms is not None else 1e-5, float(rms.clip) if rms is not None else 10.0,
            _ffi.ref_device(c.ref_device))

    def set_termination_distances(self, d) -> None:           # humanoid_phc.py:1336-1337
        self.termination_distances[:] = d

    def __call__(self, body_state, progress_buf, motion_start_times, motion_start_times_offset, sampled_motion_ids,
                 global_offset, dof_force=None, dof_vel=None, out: Optional[Dict[str, torch.Tensor]] = None,
                 env_range: Optional[Sequence[int]] = None) -> Dict[str, torch.Tensor]:
        """Run the step.  ``body_state`` is the PhysX rigid-body tensor ``[N, bodies_per_env, 13]`` (or ``[N, S]``);
        returns the dict of output buffers (owned by this object unless ``out`` supplies them).  ``env_range = (lo, hi)``
        (both multiples of 8) restricts the launch to those envs of the same buffers (used to pipeline host transfers)."""
        _ffi.require_cuda(body_state, progress_buf, motion_start_times, motion_start_times_offset, sampled_motion_ids, global_offset)
        if env_range is not None:
            lo, hi = int(env_range[0]), int(env_range[1])
            if not (0 <= lo < hi <= self.N) or lo % 8 or (hi % 8 and hi != self.N):
                raise ValueError(f"env_range {env_range}: need 0 <= lo < hi <= {self.N}, multiples of 8")
            sl = lambda t: None if t is None else t[lo:hi]          # noqa: E731
            o = out or {}
            sub = {k: sl(o.get(k, d)) for k, d in (("obs", self.obs_buf), ("obs_norm", self.obs_norm), ("reward", self.rew_buf),
                                                   ("reward_raw", self.reward_raw), ("reset", self.reset_buf), ("terminated", self.terminate_buf))}
            return self._run(hi - lo, body_state.reshape(self.N, -1)[lo:hi], sl(progress_buf), sl(motion_start_times),
                             sl(motion_start_times_offset), sl(sampled_motion_ids), sl(global_offset), sl(dof_force), sl(dof_vel),
                             sub, lo)
        return self._run(self.N, body_state, progress_buf, motion_start_times, motion_start_times_offset, sampled_motion_ids,
                         global_offset, dof_force, dof_vel, out, 0)

    def _run(self, N, body_state, progress_buf, motion_start_times, motion_start_times_offset, sampled_motion_ids, global_offset,
             dof_force, dof_vel, out, row0):
        for name, t, dt in (("progress_buf", progress_buf, torch.int16), ("motion_start_times", motion_start_times, torch.float32),
                            ("motion_start_times_offset", motion_start_times_offset, torch.float32),
                            ("sampled_motion_ids", sampled_motion_ids, torch.int64), ("global_offset", global_offset, torch.float32),
                            ("dof_force", dof_force, torch.float32), ("dof_vel", dof_vel, torch.float32)):
            if t is None:
                continue
            if t.dtype != dt or not t.is_contiguous() or t.shape[0] != N:
                raise TypeError(f"{name}: expected a contiguous {dt} tensor with {N} rows (as the reference env holds it), "
                                f"got {t.dtype} {tuple(t.shape)} contiguous={t.is_contiguous()}")
        bs = body_state.reshape(N, -1)
        if bs.dtype != torch.float32 or bs.stride(1) != 1:
            bs = bs.float().contiguous()
        use_power = self.cfg.use_power_reward
        if use_power and (dof_force is None or dof_vel is None):
            raise ValueError("use_power_reward=True needs dof_force and dof_vel")
        o = out or {}
        obs = o.get("obs", self.obs_buf)
        obs_norm = o.get("obs_norm", self.obs_norm)
        rew, raw = o.get("reward", self.rew_buf), o.get("reward_raw", self.reward_raw)
        reset, term = o.get("reset", self.reset_buf), o.get("terminated", self.terminate_buf)
        sin = _ffi.StepIn(
            bs.data_ptr(), bs.stride(0), progress_buf.data_ptr(), motion_start_times.data_ptr(), motion_start_times_offset.data_ptr(),
            sampled_motion_ids.data_ptr(), global_offset.data_ptr(),
            dof_force.data_ptr() if use_power else None, dof_vel.data_ptr() if use_power else None,
            self.termination_distances.data_ptr(),
            self.rms.running_mean.data_ptr() if self.normalize else None, self.rms.running_var.data_ptr() if self.normalize else None, N)
        sout = _ffi.StepOut(
            obs.data_ptr(), obs.stride(0), obs_norm.data_ptr() if self.normalize else None, rew.data_ptr(), raw.data_ptr(),
            raw.stride(0), reset.data_ptr(), term.data_ptr(), self.partials.data_ptr() if self.accumulate_moments else None,
            1 if self.defer_moments else 0,
            self.ref_t[row0:].data_ptr() if self.ref_t is not None else None,
            self.ref_t1[row0:].data_ptr() if self.ref_t1 is not None else None,
            self.metric_partials.data_ptr() if self.metrics else None)
        with _ffi.on_device(self.device):
            _ffi.check(self.lib.phc_step_fused(C.byref(self.motion_lib.ctables_for(self._ccfg.ref_device)), C.byref(sin), C.byref(self._ccfg), C.byref(sout),
                                               _ffi.stream_ptr()), "phc_step_fused")
            if self.defer_moments:
                if not torch.cuda.is_current_stream_capturing():      # a capture runs no kernel: replays are counted by count_replayed()
                    self._pending_rows += N
            elif self.accumulate_moments or self.metrics:
                self._reduce(N, zero=False)
        res = {"obs": obs, "reward": rew, "reward_raw": raw, "reset": reset, "terminated": term}
        if self.normalize:
            res["obs_norm"] = obs_norm
        return res

    def count_replayed(self, steps: int = 1) -> None:
        """``defer_moments`` with CUDA-graph replay: a replay runs the kernel (which adds its sums to the slots) but not the Python
        bookkeeping, so tell the object how many steps were replayed before ``flush_moments()``."""
        if self.defer_moments:
            self._pending_rows += int(steps) * self.N

    def _reduce(self, rows: int, zero: bool) -> None:
        """phc_stats_reduce: per-CTA moment / metric partials (+ the auto-reset tail's corrections) -> ``self.stats``."""
        mp = self.partials if self.accumulate_moments else None
        _ffi.check(self.lib.phc_stats_reduce(_ffi.ptr(mp), 0 if mp is None else self.active_partials(),
                                             OBS_DIM, int(rows) if mp is not None else 0, _ffi.ptr(self.row_adjust) if mp is not None else None,
                                             _ffi.ptr(self.metric_partials), self.num_partials if self.metrics else 0,
                                             _ffi.ptr(self.stats), 1 if zero else 0, _ffi.stream_ptr()), "phc_stats_reduce")

    def active_partials(self) -> int:
        """Partial slots a fold has to visit: the step kernel's, plus the auto-reset tail's once an AutoReset is attached."""
        return self.partials.shape[0] if (self.defer_moments and self.tail_slots_used) else self.num_partials

    def flush_moments(self) -> None:
        """``defer_moments`` mode: fold the sums the kernel has accumulated since the last flush (moments, the auto-reset tail's
        corrections and the episode metrics) into ``self.stats`` and clear the per-CTA slots -- ONE launch."""
        if self.defer_moments and self._pending_rows:
            with _ffi.on_device(self.device):
                self._reduce(self._pending_rows, zero=True)
            self._pending_rows = 0

    def metric_values(self, reset: bool = False) -> Dict[str, float]:
        """The accumulated episode metrics as a dict (one device->host read of 16 doubles); after ``rms.finalize()`` on several
        ranks they are the global sums.  ``reset`` clears them."""
        m = self.stats[1 + 2 * OBS_DIM:]
        v = m.cpu().tolist()
        if reset:
            m.zero_()
        return {k: v[i] for i, k in enumerate(_ffi.METRIC_NAMES)}

    # ---- host-buffer entry (end-to-end path): H2D of the per-env inputs, the kernel, D2H of reward / flags ------------
    _HOST_KEYS = ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")

    def step_host(self, host: Dict[str, torch.Tensor], chunks: int = 4, h2d_streams: int = 1, wait: bool = True):
        """Same step with HOST (ideally pinned) input tensors, as a simulator living on the host would hand them over.
        Returns host tensors ``reward, reward_raw, reset, terminated`` (valid on return); the observation buffers stay on
        the device for the policy (``self.obs_buf`` / ``self.obs_norm``).

        The envs are processed in ``chunks`` ranges: the H2D copy of range c+1 (copy stream) runs under the kernel of range c
        (compute stream) and the D2H of range c-1 (second copy stream), so only the PCIe transfer of the inputs is exposed.

        ``wait=False`` queues the step and returns a ``HostStepHandle`` at once; ``handle.result()`` blocks until this step's
        results have landed in host memory and returns them.  Staging buffers (device inputs, pinned results) are double-buffered,
        so a caller that alternates two env groups -- submit group B, then read group A's results -- keeps the H2D copy of one
        step running under the kernels and the result read-back of the other; at most TWO steps may be in flight (the third
        submission reuses the first one's buffers: call ``result()`` on a handle before submitting the step after next)."""
        keys = [k for k in self._HOST_KEYS if k in host and (self.cfg.use_power_reward or not k.startswith("dof_"))]
        if not hasattr(self, "_dev_in"):
            from . import hostmem                    # result buffers pinned on the NUMA node this GPU hangs off
            self._dev_in = [{k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.device) for k in keys} for _ in range(2)]
            self._host_out = [{"reward": hostmem.pinned_empty(self.N, torch.float32, self.device),
                               "reward_raw": hostmem.pinned_empty((self.N, self.raw_dim), torch.float32, self.device),
                               "reset": hostmem.pinned_empty(self.N, torch.bool, self.device),
                               "terminated": hostmem.pinned_empty(self.N, torch.bool, self.device)} for _ in range(2)]
            self._slot_read = [None, None]           # event: the kernels that read staging slot s have finished
            self._host_seq = 0
            self.host_h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in keys)
            self.host_d2h_bytes = sum(v.numel() * v.element_size() for v in self._host_out[0].values())
        for k in keys:
            if host[k].is_cuda:
                raise RuntimeError("step_host expects host tensors; use __call__ for device-resident inputs")
        if not hasattr(self, "_h2d_streams") or len(self._h2d_streams) != max(1, int(h2d_streams)):
            # several copy streams: the fixed set-up time of one DMA transfer hides under the payload of another
            self._h2d_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, int(h2d_streams)))]
            self._d2h_stream = torch.cuda.Stream(device=self.device)
        slot = self._host_seq & 1
        self._host_seq += 1
        d, ho, N = self._dev_in[slot], self._host_out[slot], self.N
        chunks = max(1, min(int(chunks), N // 8)) if N >= 8 else 1
        per = -(-N // chunks)
        per += (-per) % 8
        main = torch.cuda.current_stream(self.device)
        hs = self._h2d_streams
        # this slot's staging buffers were last read by the kernels of the step before last: the copies wait for THOSE kernels
        # only (not for the previous step's, which use the other slot and may still be queued behind their own copies)
        if self._slot_read[slot] is not None:
            for st in hs:
                st.wait_event(self._slot_read[slot])
        else:
            for st in hs:
                st.wait_stream(main)                    # first use: earlier work on the compute stream may still own the memory
        # the kernels below overwrite the device-side result buffers the previous step's read-back copies from
        main.wait_stream(self._d2h_stream)
        small = [k for k in keys if host[k].numel() * host[k].element_size() < (4 << 20)]     # per-env scalars: one copy each,
        with torch.cuda.stream(hs[-1]):                                                         # not one per range
            for k in small:
                d[k].copy_(host[k], non_blocking=True)
        big = [k for k in keys if k not in small]
        for lo in range(0, N, per):
            hi = min(lo + per, N)
            for i, k in enumerate(big):
                with torch.cuda.stream(hs[i % len(hs)]):
                    d[k][lo:hi].copy_(host[k][lo:hi], non_blocking=True)
            for st in hs:
                main.wait_stream(st)
            out = self(d["body_state"], d["progress"], d["start_time"], d["start_offset"], d["motion_ids"], d["global_offset"],
                       d.get("dof_force"), d.get("dof_vel"), env_range=(lo, hi))
            self._d2h_stream.wait_stream(main)
            with torch.cuda.stream(self._d2h_stream):
                for k, v in ho.items():
                    v[lo:hi].copy_(out[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        self._slot_read[slot] = ev
        done = torch.cuda.Event()
        done.record(self._d2h_stream)
        handle = HostStepHandle(ho, done)
        if not wait:
            return handle
        handle.result()
        main.synchronize()
        return ho

    # ---- CUDA-graph replay: at small batch sizes the step is launch-bound (4096 envs = a few microseconds of GPU work) ------
    def capture(self, body_state, progress_buf, motion_start_times, motion_start_times_offset, sampled_motion_ids, global_offset,
                dof_force=None, dof_vel=None, out: Optional[Dict[str, torch.Tensor]] = None, extra=None):
        """Capture one step (plus an optional ``extra()`` callable that enqueues further kernels, e.g. the GAE pass) on the
        given, fixed input/output buffers into a CUDA graph.  Returns ``(graph, outputs)``; call ``graph.replay()`` each
        step after the simulator has refreshed the input buffers in place (as Isaac Gym does)."""
        args = (body_state, progress_buf, motion_start_times, motion_start_times_offset, sampled_motion_ids, global_offset, dof_force, dof_vel)
        # the two warm-up calls below must leave no trace in the normaliser statistics: fold what is pending first, then discard
        # the warm-up's contribution (per-CTA slots / pending moments) after it
        self.flush_moments()
        saved_moments = self.rms.moments_buffer().clone() if (self.accumulate_moments and self.rms is not None) else None
        saved_metrics = self.stats[1 + 2 * OBS_DIM:].clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm-up outside capture (lazy initialisation, allocator)
                res = self(*args, out=out)
                if extra is not None:
                    extra()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if self.accumulate_moments:
            if self.defer_moments:
                self.partials.zero_()
                self._pending_rows = 0
            if saved_moments is not None:
                self.rms.moments_buffer().copy_(saved_moments)
        if self.metrics:
            self.metric_partials.zero_()
            self.stats[1 + 2 * OBS_DIM:].copy_(saved_metrics)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            res = self(*args, out=out)
            if extra is not None:
                extra()
        return graph, res
