/*
 * phc_b200.h -- C ABI of libphc_b200.so: hand-written sm_100a CUDA kernels for the per-step,
 * data-parallel rollout hot path of howird/puffer-phc.
 *
 * The reference has no FFI/plugin registry for this path; its boundary is a Python call surface
 * (SURVEY.md section 8b).  Each entry point below is what a binding for that surface calls, and
 * cites the reference function it replaces (paths relative to the reference checkout).  The
 * ctypes binding that ships with this repo is puffer_phc_b200/_ffi.py; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only: pointers, sizes, strides; no torch / CUDA types in the signatures
 *    (a stream is passed as void* holding a cudaStream_t; NULL = legacy default stream).
 *  - every data pointer is a DEVICE pointer unless its name ends in _h.  The library never
 *    allocates, frees or retains caller memory; table pointers are borrowed for one call.
 *  - every function only ENQUEUES work on `stream` and returns; there is no host synchronisation
 *    and no device->host read inside the library.
 *  - return value: 0 = PHC_OK; < 0 = argument error (enum below); > 0 = a cudaError_t from the
 *    launch.  phc_last_error() returns a thread-local description of the last failure.
 *  - all floating point is IEEE fp32 (kernels are compiled with -fmad=false, precise div/sqrt and
 *    the precise libm), quaternions are xyzw, index types are int64 like torch.long.
 *  - J (bodies per env) is 24 for the SMPL humanoid (puffer_phc/body_sets.py:11-36); the stand-alone
 *    entry points accept 1 <= J <= 32 (body subsets), the fused step requires 24.
 */
#ifndef PHC_B200_H
#define PHC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHC_B200_VERSION 121 /* 0.1.3 */

enum {
    PHC_OK = 0,
    PHC_EINVAL = -1,       /* NULL where a pointer is required, negative size, bad flag       */
    PHC_EALIGN = -2,       /* pointer / stride alignment the kernel needs is not met          */
    PHC_ESHAPE = -3,       /* J, C or L outside the supported range                           */
    PHC_EUNSUPPORTED = -4  /* semantically valid in the reference but not implemented (e.g. time_steps != 1) */
};

/* REFERENCE DEVICE FLAVOUR.  The reference runs this path with torch on CUDA (puffer_phc/envs/humanoid_phc.py) but can also be
 * run on the CPU (where the committed golden vectors were made).  torch rounds three small reductions differently on the two
 * devices -- torch.sum(q0*q1,-1) inside slerp (torch_utils.py:113), torch.norm(.., dim=-1) over xyz in the termination test
 * (envs/common.py:343) and .mean(-1) of the eval variant (:344) -- plus tensor / python-scalar in sample_time_interval
 * (motion_lib.py:533).  All four were fitted bit-exactly on the B200 box (profiles/r2_torch_device_flavours.md).  Every entry
 * point that contains one takes `ref_device` and reproduces that device's rounding, so flags are bit-exact against either. */
enum { PHC_REF_DEVICE_CPU = 0, PHC_REF_DEVICE_CUDA = 1 };

typedef void *phc_stream_t; /* cudaStream_t */

int phc_version(void);
const char *phc_last_error(void);

/* A [N, J, C] fp32 tensor whose innermost (component) stride is 1: element (n, j, c) is at
 * ptr[n * stride_env + j * stride_body + c].  Strides are in floats.  This covers both the
 * contiguous copies and the strided views of the PhysX AoS rigid-body buffer that the reference
 * passes (puffer_phc/envs/humanoid_phc.py:542-549).  For a [N, C] tensor stride_body is ignored. */
typedef struct phc_view {
    const float *ptr;
    int64_t stride_env;
    int64_t stride_body;
} phc_view;

/* ------------------------------------------------------------------------------------------- */
/* Motion tables built by MotionLibBase.load_motions (puffer_phc/motion_lib.py:396-420).        */
/* All contiguous.  F = total frames, M = loaded motions.                                       */
/* ------------------------------------------------------------------------------------------- */
typedef struct phc_motion_tables {
    const float *gts;            /* [F,24,3] global translation                      */
    const float *grs;            /* [F,24,4] global rotation                         */
    const float *lrs;            /* [F,24,4] local rotation                          */
    const float *gvs;            /* [F,24,3] global linear velocity                  */
    const float *gavs;           /* [F,24,3] global angular velocity                 */
    const float *dvs;            /* [F,23,3] dof velocity                            */
    const float *motion_aa;      /* [F,72]   _motion_aa                              */
    const float *motion_len;     /* [M] _motion_lengths                              */
    const float *motion_dt;      /* [M] _motion_dt                                   */
    const int64_t *num_frames;   /* [M] _motion_num_frames                           */
    const int64_t *length_starts;/* [M] exclusive cumsum of num_frames (:416-419)    */
    const float *motion_bodies;  /* [M,17] _motion_bodies                            */
    const float *limb_weights;   /* [M,10] _motion_limb_weights                      */
    const float *packed;         /* optional [F,312]: per frame gts|grs|gvs|gavs rows back to back, written by
                                    phc_pack_frames(); NULL = gather from the four separate tables */
    const float *pair_aux;       /* optional [F,24,2], 16-byte aligned, written by phc_build_pair_aux(): what slerp needs to know
                                    about the rotation pair (grs[f], grs[min(f+1, last frame of the clip)]) of every body --
                                    h = acos|q0.q1| and +-1/sqrt(1-(q0.q1)^2), or the fall-back codes -- computed once per pair with
                                    the very operations the kernels use at query time (bit-identical results)          */
    const uint8_t *pair_flags;   /* optional [F]: bit 0 = some body of the pair takes slerp's un-normalised midpoint fall-back
                                    (|sin_half| < 0.001), i.e. frame f+1 is needed even when blend == 0                */
    int pair_device;             /* PHC_REF_DEVICE_* the pair tables were built for (the dot product's summation order) */
    int64_t F, M;
} phc_motion_tables;

/* Concatenate the four per-frame rows the step needs into one 1248-byte record per frame
 * (B200 layout: one contiguous 1248 B gather per frame instead of four 288/384 B ones).
 * packed: caller-allocated [F,312] fp32, 16-byte aligned. */
int phc_pack_frames(const phc_motion_tables *t, float *packed, phc_stream_t stream);

/* Pair tables of the fused step (see phc_motion_tables.pair_aux): with them a query whose time sits exactly on a table frame
 * (blend == 0: ~84 % of the queries when control and motion run at the same rate) gathers ONE frame record + 192 B instead of two
 * records, and no query evaluates acos / sqrt / reciprocal of the pair again.  Needs grs, num_frames, length_starts. */
int phc_build_pair_aux(const phc_motion_tables *t, int ref_device, float *pair_aux /* [F,24,2] */, uint8_t *pair_flags /* [F] */,
                       phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* Motion TABLE BUILD ("next" row f4): MotionLibSMPL.load_motion_with_skeleton                   */
/* (puffer_phc/motion_lib.py:744-825) = SkeletonState.from_rotation_and_root_translation          */
/* (is_local=False) + SkeletonMotion.from_skeleton_state (puffer_phc/poselib_skeleton.py:1167-   */
/* 1249: forward kinematics, np.gradient + gaussian-filtered velocities) + compute_motion_dof_   */
/* vels_jit (motion_lib.py:119-140), for all clips of a library in ONE launch, written straight  */
/* into the concatenated tables of load_motions (motion_lib.py:405-412).  The reference's        */
/* precision mix is reproduced: float64 rotations, float32 forward kinematics, float32 gradient, */
/* double-accumulated gaussian filter (see oracle/loader_oracle.c).                              */
/* Raw clips are in the pkl format of scripts/convert_amass_data.py:186-196, concatenated:       */
/* pose_quat_global [Fin,J,4] float64 (GLOBAL rotations, xyzw), root_trans_offset [Fin,3]        */
/* float64.  Clip m reads raw frames [in_start[m], in_start[m] + num_frames[m]) (the crop of     */
/* motion_lib.py:773-785 is folded into in_start) and writes table rows                          */
/* [out_start[m], out_start[m] + num_frames[m]).  num_frames[m] >= 2 (the reference raises on    */
/* one-frame clips; the kernel writes zero velocities for them).                                 */
/* ------------------------------------------------------------------------------------------- */
#define PHC_BUILD_TILE 32 /* frames per CTA; tile_prefix counts ceil(num_frames / PHC_BUILD_TILE) per clip */

typedef struct phc_build_in {
    const double *pose_quat_global;  /* [Fin,J,4] float64, 16-byte aligned                                       */
    const double *root_trans;        /* [Fin,3]   float64                                                        */
    const int64_t *in_start;         /* [M] first raw frame of clip m (crop start included)                      */
    const int64_t *num_frames;       /* [M] frames kept                                                          */
    const int64_t *out_start;        /* [M] first table row of clip m (= length_starts, motion_lib.py:416-419)   */
    const int32_t *fps;              /* [M] curr_file.get("fps", 30)                                             */
    const int64_t *tile_prefix;      /* [M+1] exclusive prefix sum of ceil(num_frames / PHC_BUILD_TILE)          */
    const int32_t *parents;          /* [J] SkeletonTree.parent_indices (-1 = root; parents precede children)    */
    const float *local_translation;  /* [M or 1, J, 3] SkeletonTree.local_translation of skeleton_trees[m]       */
    int64_t lt_clip_stride;          /* floats between the skeletons of consecutive clips; 0 = one shared tree   */
    const double *heading;           /* optional [M]: random heading angle (rad) applied to rotations and root
                                        translation before the build (motion_lib.py:789-799); NULL = none        */
    int64_t M;
    int64_t n_tiles;                 /* = tile_prefix[M] (the library never reads device memory on the host)     */
    int J;                           /* bodies, 2..32                                                            */
} phc_build_in;

typedef struct phc_build_out {
    float *gts;     /* [F,J,3] */
    float *grs;     /* [F,J,4] 16-byte aligned */
    float *lrs;     /* [F,J,4] 16-byte aligned */
    float *gvs;     /* [F,J,3] */
    float *gavs;    /* [F,J,3] */
    float *dvs;     /* [F,J-1,3] */
    float *packed;  /* optional [F,312] (J == 24 only): the frame records of phc_pack_frames, written in the same pass */
} phc_build_out;

int phc_build_motion_tables(const phc_build_in *in, const phc_build_out *out, phc_stream_t stream);

/* _motion_aa (motion_lib.py:381, 399): slot s contributes the float64 pose_aa rows of its WHOLE clip (the reference appends
 * the uncropped array): raw rows [seg_src[s], seg_src[s] + len_s) -> float32 table rows [seg_dst_prefix[s], seg_dst_prefix[s+1]).
 * heading (optional [S], with crop_lo / crop_hi [S] = the crop relative to the clip): the root rotation vector of the rows
 * inside the crop becomes (h * from_rotvec(rv)).as_rotvec() (motion_lib.py:793). */
int phc_build_motion_aa(const double *pose_aa, int row_len, const int64_t *seg_src, const int64_t *seg_dst_prefix, int64_t S,
                        int64_t n_rows, const double *heading, const int64_t *crop_lo, const int64_t *crop_hi, float *out,
                        phc_stream_t stream);

/* plain float64 -> float32 cast of n elements */
int phc_cast_f64_f32(const double *x, int64_t n, float *y, phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* MotionLibBase.get_motion_state(motion_ids, motion_times, offset=None)                        */
/* (puffer_phc/motion_lib.py:549-626) and get_root_pos_smpl (:628-653).                         */
/* Any output pointer may be NULL (that output is skipped); get_root_pos_smpl = only root_pos.  */
/* All outputs contiguous [B, ...] fp32.                                                        */
/* ------------------------------------------------------------------------------------------- */
typedef struct phc_motion_state_out {
    float *root_pos;            /* [B,3]    */
    float *root_rot;            /* [B,4]    */
    float *dof_pos;             /* [B,69]   */
    float *root_vel;            /* [B,3]    */
    float *root_ang_vel;        /* [B,3]    */
    float *dof_vel;             /* [B,69]   */
    float *motion_aa;           /* [B,72]   */
    float *rg_pos;              /* [B,24,3] */
    float *rb_rot;              /* [B,24,4] */
    float *body_vel;            /* [B,24,3] */
    float *body_ang_vel;        /* [B,24,3] */
    float *motion_bodies;       /* [B,17]   */
    float *motion_limb_weights; /* [B,10]   */
    int64_t *frame_idx0;        /* [B] optional: _calc_frame_blend outputs (:655-665) */
    int64_t *frame_idx1;        /* [B] optional */
    float *blend;               /* [B] optional */
} phc_motion_state_out;

int phc_motion_state(const phc_motion_tables *t, const int64_t *motion_ids, const float *motion_times,
                     const float *offset /* [B,3] or NULL */, int64_t B, const phc_motion_state_out *out,
                     int ref_device, phc_stream_t stream);

/* Reset path ("next" row f1): HumanoidPHC._sample_ref_state + _set_env_state for the envs listed in env_ids
 * (puffer_phc/envs/humanoid_phc.py:843-873, 899-929): one get_motion_state query per reset env
 * (motion id = sampled_motion_ids[e], time = motion_times[i], offset = global_offset[e]) whose result is written
 * straight into the env's state tensors at row e instead of being materialised and scattered by 12 index_put ops:
 *   root_states[e]      = root_pos | root_rot | root_vel | root_ang_vel           [N,13]
 *   dof_pos[e], dof_vel[e]                                                          [N,69]
 *   body_state[e, j]    = rg_pos | rb_rot | body_vel | body_ang_vel, j < 24         AoS [N, env_stride]
 * env_ids: [K] int64, ascending and unique (torch.nonzero order); motion_times: [K]. Any output may be NULL. */
int phc_reset_ref_state(const phc_motion_tables *t, const int64_t *env_ids, const int64_t *sampled_motion_ids /* [N] */,
                        const float *motion_times /* [K] */, const float *global_offset /* [N,3] or NULL */, int64_t K,
                        float *root_states, float *dof_pos, float *dof_vel, float *body_state, int64_t env_stride,
                        int ref_device, phc_stream_t stream);

/* MotionLibBase._calc_frame_blend(time, len, num_frames, dt) (puffer_phc/motion_lib.py:655-665), elementwise over n entries:
 * phase = clip(time/len, 0, 1); time < 0 -> 0; idx0 = int64(phase * (nf-1)); idx1 = min(idx0+1, nf-1);
 * blend = clip((time - idx0*dt)/dt, 0, 1).  Integer outputs are bit-exact against torch. */
int phc_frame_blend(const float *time, const float *len, const int64_t *num_frames, const float *dt, int64_t n,
                    int64_t *frame_idx0, int64_t *frame_idx1, float *blend, phc_stream_t stream);

/* MotionLibBase.sample_time_interval arithmetic (puffer_phc/motion_lib.py:526-535); the uniform
 * phase stays on the caller's torch generator.  out = float(int64((phase*len)/(1/30))) * (1/30).
 * div_mode = ref_device: 0 IEEE division by float32(1/30) (torch CPU); 1 multiplication by float(1.0 / (1/30)) = 30.0f, the
 * reciprocal formed in double (how torch CUDA divides a tensor by a Python scalar).  motion_len is already gathered per sample. */
int phc_sample_time_interval(const float *phase, const float *motion_len, int64_t n, int div_mode, float *out,
                             phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* puffer_phc/envs/common.py                                                                    */
/* ------------------------------------------------------------------------------------------- */

/* compute_imitation_observations_v6 (common.py:106-176).  obs: [N, obs_stride >= 24*J*time_steps]: per future step six
 * body-major blocks [3J,6J,3J,3J,3J,6J], the time_steps groups one after the other.  The ref_* views describe
 * N * time_steps rows (row b * time_steps + s = future step s of env b, i.e. the reference's .view(B, time_steps, J, .));
 * the reference itself only ever passes time_steps = 1 (humanoid_phc.py:1097). */
int phc_imitation_obs_v6(phc_view root_pos /* [N,3] */, phc_view root_rot /* [N,4] */,
                         phc_view body_pos, phc_view body_rot, phc_view body_vel, phc_view body_ang_vel,
                         phc_view ref_body_pos, phc_view ref_body_rot, phc_view ref_body_vel, phc_view ref_body_ang_vel,
                         int64_t N, int J, int time_steps, int upright, float *obs, int64_t obs_stride,
                         phc_stream_t stream);

/* compute_humanoid_observations_smpl_max (common.py:23-103) without the smpl / limb-weight
 * pass-through columns (those are plain concatenations done by the caller).
 * obs: [N, obs_stride >= (root_height_obs?1:0) + 3(J-1) + 12J]. */
int phc_self_obs_smpl_max(phc_view body_pos, phc_view body_rot, phc_view body_vel, phc_view body_ang_vel,
                          int64_t N, int J, int local_root_obs, int root_height_obs, int upright,
                          float *obs, int64_t obs_stride, phc_stream_t stream);

/* compute_imitation_reward (common.py:270-322).  k, w: HOST arrays of 4 floats in the order
 * pos, rot, vel, ang_vel (rwd_specs k_* / w_*).  reward [N]; reward_raw [N, raw_stride >= 4]. */
int phc_imitation_reward(phc_view body_pos, phc_view body_rot, phc_view body_vel, phc_view body_ang_vel,
                         phc_view ref_body_pos, phc_view ref_body_rot, phc_view ref_body_vel, phc_view ref_body_ang_vel,
                         int64_t N, int J, const float *k_h, const float *w_h, float *reward, float *reward_raw,
                         int64_t raw_stride, phc_stream_t stream);

/* compute_humanoid_im_reset (common.py:325-364).  progress: int16 (humanoid_phc.py:571);
 * pass_time / reset / terminated: 1 byte per env (torch.bool); termination_distance: device [J]
 * (only element 0 is read when use_mean). */
int phc_im_reset(const int16_t *progress, phc_view rigid_body_pos, phc_view ref_body_pos, const uint8_t *pass_time,
                 int enable_early_termination, const float *termination_distance, int use_mean,
                 int64_t N, int J, uint8_t *reset, uint8_t *terminated, int ref_device, phc_stream_t stream);

/* The evaluation metric HumanoidPHC.step adds when flag_im_eval is set (puffer_phc/envs/humanoid_phc.py:159-163, consumed by
 * EvalStats, scripts/train.py:139-166): mpjpe[n] = mean_j || body_pos[n,j] - ref_body_pos[n,j] ||, ref = rg_pos at t. */
int phc_mpjpe(phc_view body_pos, phc_view ref_body_pos, int64_t N, int J, float *mpjpe, int ref_device, phc_stream_t stream);

/* build_amp_observations_smpl + dof_to_obs_smpl (common.py:179-267; "next" row f3, only used with use_amp_obs) without the
 * shape / limb-weight pass-through columns.  Contiguous inputs: root_* [N,3|4], dof_pos / dof_vel [N,69], key_body_pos [N,K,3];
 * dof_subset: device [3*num_joints] int64 indices into the 69-dof vector, or NULL = all 23 joints in order.
 * obs: [N, obs_stride >= (root_height_obs?1:0) + 12 + 9*num_joints + 3*K]. */
int phc_amp_obs_smpl(const float *root_pos, const float *root_rot, const float *root_vel, const float *root_ang_vel,
                     const float *dof_pos, const float *dof_vel, const float *key_body_pos, const int64_t *dof_subset,
                     int num_joints, int K, int local_root_obs, int root_height_obs, int upright, int64_t N, float *obs,
                     int64_t obs_stride, int ref_device, phc_stream_t stream);

/* The AMP observation step of HumanoidPHC.step (humanoid_phc.py:154-157) on the history buffer _amp_obs_buf [N, num_steps, row_width]
 * (:600-606): _update_hist_amp_obs (:1339-1348: rows 1.. = old rows 0..num_steps-2) and _compute_amp_observations (:1123-1174: row 0 =
 * the current observation) in ONE pass -- the reference clones the buffer, copies it back shifted, computes the row and copies it in.
 * Same arguments as phc_amp_obs_smpl; 2 <= num_steps <= 16; row_width >= the kernel's row (pass-through columns stay the caller's). */
int phc_amp_obs_hist_step(const float *root_pos, const float *root_rot, const float *root_vel, const float *root_ang_vel,
                          const float *dof_pos, const float *dof_vel, const float *key_body_pos, const int64_t *dof_subset,
                          int num_joints, int K, int local_root_obs, int root_height_obs, int upright, int64_t N,
                          float *amp_obs_buf, int num_steps, int row_width, int ref_device, phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* The whole post-physics half of HumanoidPHC.step in ONE pass over HBM                          */
/* (puffer_phc/envs/humanoid_phc.py:136-149: _compute_reward :1228-1303, _compute_reset          */
/* :1311-1333, _compute_observations :935-959) with the two get_motion_state queries (t, t+1)    */
/* fused in, plus (optionally) RunningNorm.forward on the fresh observation and the per-column   */
/* moment partial sums RunningNorm.update needs (puffer_phc/policies/running_norm.py:15-34).     */
/* ------------------------------------------------------------------------------------------- */
typedef struct phc_step_in {
    const float *body_state;     /* PhysX rigid-body tensor, AoS [N, env_stride floats]; body j of env n at
                                    n*env_stride + 13*j: pos3, rot4 (xyzw), vel3, angvel3   (humanoid_phc.py:542-549) */
    int64_t env_stride;          /* >= 312 */
    const int16_t *progress;     /* [N] progress_buf, already incremented (humanoid_phc.py:138)            */
    const float *start_time;     /* [N] _motion_start_times                                                */
    const float *start_offset;   /* [N] _motion_start_times_offset                                         */
    const int64_t *motion_ids;   /* [N] _sampled_motion_ids                                                */
    const float *global_offset;  /* [N,3] _global_offset                                                   */
    const float *dof_force;      /* [N,69] or NULL: power reward off (humanoid_phc.py:1295-1303)           */
    const float *dof_vel;        /* [N,69] or NULL                                                         */
    const float *term_dist;      /* device [24] _termination_distances, indexed by body id                 */
    const float *rms_mean;       /* device [934] or NULL (needed iff out.obs_norm != NULL)                 */
    const float *rms_var;        /* device [934] or NULL                                                   */
    int64_t N;
} phc_step_in;

typedef struct phc_step_cfg {
    float dt;                    /* float32(isaac dt) = 1/30 at defaults (isaacgym_env.py:39-41)           */
    float k[4], w[4];            /* RewardConfig k_pos,k_rot,k_vel,k_ang_vel / w_* (config.py:25-32)       */
    float power_coef;            /* rew_power_coef (config.py:96)                                          */
    uint32_t reset_body_mask;    /* bit j = body j takes part in the termination test (_reset_bodies_id)   */
    int enable_early_termination;
    int use_mean;                /* flag_im_eval: mean distance against term_dist[first reset body]        */
    float rms_eps, rms_clip;     /* RunningNorm epsilon / clip (running_norm.py:6)                         */
    int ref_device;              /* PHC_REF_DEVICE_CPU / _CUDA: whose rounding the slerp dot product, the termination norm
                                    and the eval-mode mean reproduce                                       */
} phc_step_cfg;

typedef struct phc_step_out {
    float *obs;                  /* [N, obs_stride>=934] raw obs_buf: self 358 | task 576 (humanoid_phc.py:947) */
    int64_t obs_stride;
    float *obs_norm;             /* optional [N, obs_stride]: clamp((obs-mean)/sqrt(var+eps), +-clip)       */
    float *reward;               /* [N] rew_buf                                                            */
    float *reward_raw;           /* [N, raw_stride]: r_pos,r_rot,r_vel,r_ang_vel[,power]                   */
    int64_t raw_stride;          /* >= 4, >= 5 with the power reward                                       */
    uint8_t *reset;              /* [N] reset_buf                                                          */
    uint8_t *terminated;         /* [N] _terminate_buf                                                     */
    double *moment_partials;     /* optional [phc_step_num_partials(), 2, 934] fp64: per-CTA sum and sum of
                                    squares of the raw obs columns; reduce with phc_rms_reduce_partials()  */
    int accumulate_partials;     /* 0: the slots are overwritten with this step's sums; 1: this step's sums are ADDED to the
                                    slots (CTA -> slot and env -> CTA are fixed, so the result is deterministic): zero the
                                    buffer once, step a whole rollout, reduce once                         */
    float *ref_state_t;          /* optional debug [N,312]: blended reference pos72|rot96|vel72|ang72 at t */
    float *ref_state_t1;         /* optional debug [N,312] at t+1                                          */
    double *metric_partials;     /* optional [phc_step_num_partials(), PHC_NUM_METRICS] fp64: per-CTA sums of the step's episode
                                    metrics PHC_M_STEPS .. PHC_M_TERMINATIONS (same overwrite / accumulate rule as
                                    moment_partials); fold with phc_stats_reduce()                         */
} phc_step_out;

/* Episode metrics the reference logs (puffer_phc/clean_pufferl/env.py:102-110, 118-131, 139-164): sums, so that ranks can be
 * all-reduced and means formed afterwards.  phc_step_fused fills 0..8, phc_auto_reset 9..12 (and 0..8 when asked to). */
#define PHC_NUM_METRICS 16
enum {
    PHC_M_STEPS = 0,        /* env-steps counted                                                        */
    PHC_M_REWARD = 1,       /* sum rew_buf                                                              */
    PHC_M_RAW0 = 2,         /* sum reward_raw[:, 0..4]: r_pos, r_rot, r_vel, r_ang_vel, power  (2..6)   */
    PHC_M_RESETS = 7,       /* sum reset_buf                                                            */
    PHC_M_TERMINATIONS = 8, /* sum _terminate_buf                                                       */
    PHC_M_TRUNCATIONS = 9,  /* resets that are not terminations (env.py:128-131)                        */
    PHC_M_EP_RETURN = 10,   /* sum of the finished episodes' returns (env.py:118)                       */
    PHC_M_EP_LENGTH = 11,   /* sum of the finished episodes' lengths (env.py:119)                       */
    PHC_M_EPISODES = 12     /* finished episodes (env.py:117)                                           */
};

/* number of [2,934] fp64 partial slots phc_step_fused writes (a function of the device only) */
int phc_step_num_partials(void);

int phc_step_fused(const phc_motion_tables *t, const phc_step_in *in, const phc_step_cfg *cfg,
                   const phc_step_out *out, phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* Auto-reset after the step ("next" row f1), on the device, no host synchronisation:           */
/* PHCPufferEnv.step bookkeeping (puffer_phc/clean_pufferl/env.py:102-140) + HumanoidPHC.reset  */
/* of the flagged envs = _reset_envs (puffer_phc/envs/humanoid_phc.py:663-674):                 */
/* _sample_ref_state (:843-873) + _set_env_state (:899-929) + the per-env scalars (:721-727,    */
/* :774-777) + _compute_observations(env_ids) (:935-959).  Replaces torch.nonzero (a host sync),*/
/* ~20 indexed assignments, a subset get_motion_state x2 and the subset observation functions.  */
/* Two kernel launches, CUDA-graph capturable together with phc_step_fused.                     */
/* ------------------------------------------------------------------------------------------- */
typedef struct phc_reset_env {   /* the per-env tensors HumanoidPHC owns (humanoid_phc.py:523-598); flagged rows are rewritten in place */
    float *body_state;           /* PhysX rigid-body tensor AoS [N, env_stride]: _rigid_body_pos/rot/vel/ang_vel of the flagged envs */
    int64_t env_stride;          /* >= 312 */
    float *root_states;          /* optional [N,13] _humanoid_root_states                                  */
    float *dof_pos, *dof_vel;    /* optional [N,69] _dof_pos / _dof_vel                                    */
    int16_t *progress;           /* [N] progress_buf      -> 0                                             */
    float *start_time;           /* [N] _motion_start_times -> the sampled start time                      */
    float *start_offset;         /* [N] _motion_start_times_offset -> 0                                    */
    float *global_offset;        /* [N,3] _global_offset: READ by the state query, then -> 0 (:721)        */
    const int64_t *motion_ids;   /* [N] _sampled_motion_ids (unchanged by a reset)                         */
    uint8_t *reset, *terminated; /* [N] reset_buf / _terminate_buf: in = this step's flags, cleared for the flagged envs */
    float *obs;                  /* [N, obs_stride >= 934] obs_buf: rows of the flagged envs recomputed    */
    int64_t obs_stride;
    float *obs_norm;             /* optional [N, obs_stride]: the RunningNorm-normalised copy of those rows */
    const float *rms_mean, *rms_var; /* device [934], needed iff obs_norm                                   */
} phc_reset_env;

typedef struct phc_reset_book {  /* PHCPufferEnv.step bookkeeping (clean_pufferl/env.py:102-140); every pointer may be NULL */
    const float *rewards;        /* [N] rew_buf of this step                                               */
    const float *reward_raw;     /* [N, raw_stride] extras["reward_raw"]                                   */
    int64_t raw_stride;
    int raw_dim;                 /* 4 or 5                                                                 */
    uint8_t *terminals, *truncations, *masks;   /* [N] out (env.py:111-133)                                */
    float *episode_returns;      /* [N] in/out (env.py:118-120, 139)                                       */
    int32_t *episode_lengths;    /* [N] in/out (env.py:119-121, 140)                                       */
    double *metrics;             /* [PHC_NUM_METRICS] device accumulators, += : always PHC_M_TRUNCATIONS .. PHC_M_EPISODES;
                                    PHC_M_STEPS .. PHC_M_TERMINATIONS only when step_metrics (otherwise phc_step_fused sums them) */
    int step_metrics;
} phc_reset_book;

typedef struct phc_reset_cfg {
    float dt;                    /* float32(isaac dt)                                                      */
    int state_init;              /* 0 = StateInit.Random / Hybrid-ref: start = sample_time_interval (motion_lib.py:526-535);
                                    1 = StateInit.Start: start = 0 (humanoid_phc.py:846-851)               */
    int flag_test;               /* motion_times[:] = 0 (humanoid_phc.py:853-854)                          */
    int ref_device;              /* PHC_REF_DEVICE_*                                                       */
    float rms_eps, rms_clip;
} phc_reset_cfg;

/* phase: device [N] uniforms drawn by the caller with torch.rand (the RNG stays torch's): the k-th flagged env in ascending env
 * order consumes phase[k], as the reference's torch.rand(len(env_ids)) assigns them.  reset_ids (optional, [N] int64) receives the
 * flagged env ids in ascending order, reset_count (optional, device int32) their number -- nothing is read back by the library.
 * scratch: >= phc_auto_reset_scratch_bytes(N) bytes, 8-byte aligned.
 * moment_partials (optional, [phc_auto_reset_num_partials(), 2, 934] fp64, ADDED to) and row_adjust (optional, device double, -=
 * truncated rows): the correction that turns column moments accumulated by phc_step_fused over the PRE-reset observations into
 * moments over the rows the reference stores (post-reset row for a terminated env, no row for a truncated env: clean_pufferl/
 * env.py:132-133, structs.py:116); fold both with phc_stats_reduce(). */
int phc_auto_reset_num_partials(void);
int64_t phc_auto_reset_scratch_bytes(int64_t N);
int phc_auto_reset(const phc_motion_tables *t, const phc_reset_env *env, const phc_reset_book *book, const phc_reset_cfg *cfg,
                   const float *phase, int64_t N, int64_t *reset_ids, int32_t *reset_count, void *scratch,
                   double *moment_partials, double *row_adjust, phc_stream_t stream);

/* Fold per-CTA partials into the rank's statistics buffer stats = [n, sum x (C), sum x^2 (C), metrics (PHC_NUM_METRICS)] -- the ONE
 * buffer that is all-reduced across ranks (RunningNorm moments + episode metrics):
 *   stats[0] += rows (+ *row_adjust); stats[1 + i] += sum_p moment_partials[p][i]; stats[1 + 2C + k] += sum_p metric_partials[p][k]
 * in a fixed order (deterministic).  zero_partials: every partial read (and *row_adjust) is cleared in the same launch. */
int phc_stats_reduce(double *moment_partials, int num_partials, int C, int64_t rows, double *row_adjust,
                     double *metric_partials, int num_metric_partials, double *stats, int zero_partials, phc_stream_t stream);

/* The exchange step fused with what precedes and follows it (multi-GPU, one process per GPU, NVLink / NVSwitch peer memory):
 * phc_stats_reduce + all-reduce(SUM) of [n, sum x, sum x^2 | metrics] over the ranks + phc_rms_finalize as ONE kernel.
 * Every rank owns an exchange buffer of phc_stats_comm_bytes(world, C) bytes, zero-initialised once, that all ranks can address
 * (e.g. torch.distributed._symmetric_memory: peer_bufs[r] = the address of rank r's buffer in THIS process; with world == 1 any
 * device buffer).  Blocks push their folded columns into every rank's buffer with plain stores, publish per-block flags
 * (st.release.sys) and add the contributions in rank order, so all ranks end with bit-identical running_mean / running_var /
 * count / metric sums.  epoch: 1, 2, 3, ... the same on every rank, incremented per call (buffers are double-buffered by its
 * parity).  ticket: a zero-initialised device uint32 owned by the caller.  running_mean / running_var / count may be NULL
 * (metrics only).  The partial slots and *row_adjust are cleared; stats[1 + 2C ..] += the GLOBAL metric sums. */
typedef struct phc_stats_comm {
    int rank, world;               /* world <= 32 */
    void *peer_bufs[32];           /* [world] device-addressable exchange buffers, index = rank */
    uint64_t epoch;
    uint32_t *ticket;
} phc_stats_comm;

int64_t phc_stats_comm_bytes(int world, int C);
int phc_stats_allreduce_finalize(double *moment_partials, int num_partials, int C, int64_t rows, double *row_adjust,
                                 double *metric_partials, int num_metric_partials, double *stats, const phc_stats_comm *comm,
                                 float *running_mean, float *running_var, float *count, phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* RunningNorm (puffer_phc/policies/running_norm.py:5-53)                                       */
/* ------------------------------------------------------------------------------------------- */

/* forward (:15-20): y = clamp((x - mean) / sqrt(var + eps), -clip, clip).  x, y: [B, C] with row
 * strides in floats; mean, var: device [C]. */
int phc_rms_forward(const float *x, int64_t x_stride, const float *mean, const float *var, float eps, float clip,
                    int64_t B, int C, float *y, int64_t y_stride, phc_stream_t stream);

/* update (:23-34), split so that per-rank moments can be all-reduced in between:
 *   phc_rms_moments         moments[0] += B; moments[1+c] += sum_b x[b,c]; moments[1+C+c] += sum_b x[b,c]^2
 *                           (fp64, deterministic two-stage reduction; scratch: >= phc_rms_scratch_doubles(C) doubles)
 *   phc_rms_reduce_partials same accumulation from the per-CTA partials phc_step_fused wrote
 *   phc_rms_finalize        mean_b = S/n; var_b = SS/n - mean_b^2 (fp64, biased); w = 1/count;
 *                           running = running*(1-w) + batch*w (fp32, reference op order); count += 1 */
int64_t phc_rms_scratch_doubles(int C);
int phc_rms_moments(const float *x, int64_t x_stride, int64_t B, int C, double *moments /* [1+2C] */,
                    double *scratch, phc_stream_t stream);
int phc_rms_reduce_partials(const double *partials, int num_partials, int64_t rows, int C, double *moments,
                            phc_stream_t stream);
int phc_rms_finalize(const double *moments, int C, float *running_mean, float *running_var, float *count,
                     phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* c_gae.compute_gae(dones, values, rewards, gamma, gae_lambda) (puffer_phc/c_gae.pyx:11-32)    */
/* Flat serial-scan semantics over the whole array of length L (the carry crosses env           */
/* boundaries; adv[L-1] = 0; reward/done indexed at t+1).  All arrays device fp32 [L].          */
/* mode 0 = auto; 1 = blocked scan with a warm-up window (bit-exact whenever the carry decays   */
/* below fp32 resolution inside the window, which auto verifies from gamma*lambda);             */
/* 2 = single-thread serial scan (always bit-exact, slow).                                      */
/* ------------------------------------------------------------------------------------------- */
int phc_gae(const float *dones, const float *values, const float *rewards, int64_t L, float gamma, float gae_lambda,
            float *advantages, int mode, phc_stream_t stream);

/* ------------------------------------------------------------------------------------------- */
/* Device-resident rollout buffer around compute_gae: the reference's */
/* Experience.store / sort_training_data (puffer_phc/clean_pufferl/structs.py:108-145) and the GAE */
/* call site (clean_pufferl/core.py:213-259), which keep host numpy arrays in arrival order, sort  */
/* batch_size (env_id, step) tuples in Python and copy four arrays between host and device.        */
/* ------------------------------------------------------------------------------------------- */

/* One env step of all N envs (env id = index) into row t of the [T, N] arrays: values_row .. mask_row are the row pointers
 * (&values[t * N] ...); done / trunc are float32 (flags_are_float != 0) or bool / uint8 vectors (trunc may be NULL); mask is
 * bool / uint8 (truncated envs are masked out, clean_pufferl/env.py:133).  Also written: subrank_row[e] = non-masked envs before e
 * in e's 32-env group, group_counts_row[g] = non-masked envs of group g (ceil(N / 32) ints per row); *row_count += the row's
 * non-masked envs (zero it when the rollout starts); *stored += the same (running total, structs.py:104-106). */
int phc_rollout_store(const float *value, const float *reward, const void *done, const void *trunc, int flags_are_float,
                      const uint8_t *mask, int64_t N, float *values_row, float *rewards_row, float *dones_row,
                      float *truncateds_row, uint8_t *mask_row, uint8_t *subrank_row, int32_t *group_counts_row,
                      int32_t *row_count, int64_t *stored, phc_stream_t stream);

/* sort_training_data + the gathers in front of compute_gae, for the T rows stored so far: the rows the reference would have stored
 * (per step the non-masked envs in env order, until batch_size rows) ordered by (env, step).  Outputs (capacity >= batch_size
 * elements each): sorted_dones / sorted_values / sorted_rewards (what core.py:249 hands to compute_gae), idxs (the reference's
 * arrival row numbers in sorted order, structs.py:133-145) and pos_em (e * T + t of every kept element: index into the env-major
 * flattening of any [T, N] array).  meta (device int64[4]) = {first step that passes batch_size (T if none), rows that step still
 * stores, rows kept in total, T}: read meta[2] to size the results.  scratch: phc_rollout_scratch_bytes(N, T) bytes, 8-byte aligned.
 * Four launches, no host synchronisation. */
int64_t phc_rollout_scratch_bytes(int64_t N, int T);
int phc_rollout_sort(const float *dones, const float *values, const float *rewards, const uint8_t *mask, const uint8_t *subrank,
                     const int32_t *group_counts, const int32_t *row_counts, int64_t N, int T, int64_t batch_size, void *scratch,
                     int64_t *meta, float *sorted_dones, float *sorted_values, float *sorted_rewards, int64_t *idxs,
                     int64_t *pos_em, phc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PHC_B200_H */
