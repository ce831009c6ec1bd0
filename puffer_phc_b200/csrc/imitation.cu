// imitation.cu -- stand-alone drop-ins for the four free functions of reference puffer_phc/envs/common.py:
// compute_imitation_observations_v6 (:106-176), compute_humanoid_observations_smpl_max (:23-103),
// compute_imitation_reward (:270-322), compute_humanoid_im_reset (:325-364).
//
// One warp per env, lane j = body j (J <= 32).  Inputs are strided views (phc_view) so the PhysX AoS
// buffer slices the reference passes are consumed in place; the heading quaternion is computed once per
// env; body reductions are warp shuffles.  When the four simulated-body views are the pos / rot / vel / ang-vel
// slices of ONE 13-float AoS record per body (exactly what humanoid_phc.py:546-549 hands over) the kernels run
// an AOS instantiation: the warp stages the env's whole record with coalesced (16-byte) loads into shared
// memory and the lanes pick their body out of it, instead of 13 strided scalar loads per lane that each touch
// ten cache lines.  (The fused step kernel in step_fused.cu shares the same
// per-body math through phc_body.cuh.)
#include "phc_body.cuh"

namespace phc {

constexpr int IM_WARPS = 4;

// kernel-side copy of a phc_view with the body stride as a 32-bit int: a lane address is one wide multiply-add on top of the env's base
// instead of a 64-bit multiplication per view (the stand-alone kernels are instruction-issue-bound, not memory-bound)
struct KView { const float* ptr; int64_t stride_env; int stride_body; };
__device__ __forceinline__ const float* at(const KView& v, int64_t n, int j) { return v.ptr + n * v.stride_env + j * v.stride_body; }

// .mean(dim=-1) over the J <= 32 lane values d in torch's summation order for the chosen device (phc_math.cuh mean_ordered); the
// result is valid in every lane.  Only the eval-mode paths (use_mean, mpjpe) come here.
__device__ __forceinline__ float warp_mean(float d, int J, int lane, int dev) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __shfl_sync(FULL, d, i);
    return mean_ordered(v, J, dev);
}

constexpr int IM_REC = 32 * REC;          // floats of staging per warp (J <= 32 bodies x 13)

// coalesced copy of one env's AoS record (nfl floats) into the warp's shared-memory buffer
__device__ __forceinline__ void stage_record(const float* rec, int nfl, float* s, int lane) {
    if ((reinterpret_cast<uintptr_t>(rec) & 15u) == 0 && (nfl & 3) == 0) {
        for (int i = lane; i < (nfl >> 2); i += 32) reinterpret_cast<float4*>(s)[i] = __ldg(reinterpret_cast<const float4*>(rec) + i);
    } else {
        for (int i = lane; i < nfl; i += 32) s[i] = __ldg(rec + i);
    }
    __syncwarp();
}
__device__ __forceinline__ BodyState body_from_record(const float* s, int j) {
    const float* b = s + REC * j;
    return BodyState{ld3(b), ld4(b + 3), ld3(b + 7), ld3(b + 10)};
}

struct ObsArgs {
    KView root_pos, root_rot, pos, rot, vel, ang, rpos, rrot, rvel, rang;
    int64_t N; int J; int upright; float* obs; int64_t obs_stride; int ts;
};

// One warp per (env, future step): with time_steps > 1 the reference views its reference tensors as [B, time_steps, J, .] and
// evaluates every future step against the SAME simulated bodies (common.py:137-173); row b of the output is the time_steps blocks of
// 24 J values one after the other.
template <bool AOS>
__global__ void __launch_bounds__(IM_WARPS * 32) imitation_obs_kernel(const ObsArgs a) {
    __shared__ __align__(16) float s_rec[AOS ? IM_WARPS * IM_REC : 4];
    const int lane = threadIdx.x & 31;
    const int64_t nv = (int64_t)blockIdx.x * IM_WARPS + (threadIdx.x >> 5);      // (env, step) index: the reference tensors' row
    if (nv >= a.N * a.ts) return;
    const int64_t n = a.ts == 1 ? nv : nv / a.ts;                                // env: the simulated tensors' row
    const int step = (int)(nv - n * a.ts);
    const float* rec = s_rec + (AOS ? (threadIdx.x >> 5) * IM_REC : 0);
    BodyState r{};                                   // reference loads first: they fly while the record is staged
    if (lane < a.J) r = BodyState{ld3(at(a.rpos, nv, lane)), ld4(at(a.rrot, nv, lane)), ld3(at(a.rvel, nv, lane)), ld3(at(a.rang, nv, lane))};
    if (AOS) stage_record(a.pos.ptr + n * a.pos.stride_env, a.J * REC, const_cast<float*>(rec), lane);
    Q4 rr = ld4(a.root_rot.ptr + n * a.root_rot.stride_env);
    if (!a.upright) rr = remove_base_rot(rr);
    float hz, hw;
    heading_quat(calc_heading(rr), hz, hw);                      // h = (0,0,hz,hw), h^-1 = (0,0,-hz,hw)
    const V3 rp = ld3(a.root_pos.ptr + n * a.root_pos.stride_env);
    if (lane >= a.J) return;
    const int j = lane, J = a.J;
    const BodyState b = AOS ? body_from_record(rec, j)
                            : BodyState{ld3(at(a.pos, n, j)), ld4(at(a.rot, n, j)), ld3(at(a.vel, n, j)), ld3(at(a.ang, n, j))};
    float* o = a.obs + n * a.obs_stride + (int64_t)step * 24 * J;
    task_obs_body_fma(b, r, rp, hz, hw, zrot_make(hz, hw), o + 3 * j, o + 3 * J + 6 * j, o + 9 * J + 3 * j, o + 12 * J + 3 * j,
                      o + 15 * J + 3 * j, o + 18 * J + 6 * j);
}

struct SelfArgs {
    KView pos, rot, vel, ang;
    int64_t N; int J; int local_root_obs, root_height_obs, upright; float* obs; int64_t obs_stride;
};

template <bool AOS>
__global__ void __launch_bounds__(IM_WARPS * 32) self_obs_kernel(const SelfArgs a) {
    __shared__ __align__(16) float s_rec[AOS ? IM_WARPS * IM_REC : 4];
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * IM_WARPS + (threadIdx.x >> 5);
    if (n >= a.N) return;
    const float* rec = s_rec + (AOS ? (threadIdx.x >> 5) * IM_REC : 0);
    if (AOS) stage_record(a.pos.ptr + n * a.pos.stride_env, a.J * REC, const_cast<float*>(rec), lane);
    Q4 rr = ld4(at(a.rot, n, 0));
    if (!a.upright) rr = remove_base_rot(rr);                    // common.py:41-42
    float hz, hw;
    heading_quat(calc_heading(rr), hz, hw);
    const V3 rp = ld3(at(a.pos, n, 0));
    if (lane >= a.J) return;
    const int j = lane, J = a.J;
    float* o = a.obs + n * a.obs_stride;
    if (a.root_height_obs) { if (j == 0) o[0] = rp.z; o += 1; }  // common.py:40, 92-93
    const BodyState b = AOS ? body_from_record(rec, j)
                            : BodyState{ld3(at(a.pos, n, j)), ld4(at(a.rot, n, j)), ld3(at(a.vel, n, j)), ld3(at(a.ang, n, j))};
    float* rot_out = o + 3 * (J - 1) + 6 * j;
    const ZRot hrot = zrot_make(hz, hw);
    self_obs_pos_rot_fma(b, rp, hz, hw, hrot, j, o + 3 * (j - 1), rot_out);
    self_obs_vel_ang_fma(b, hrot, o + 3 * (J - 1) + 6 * J + 3 * j, o + 3 * (J - 1) + 9 * J + 3 * j);
    if (!a.local_root_obs && j == 0) tan_norm(rr, rot_out);      // common.py:77-79
}

struct RewardArgs {
    KView pos, rot, vel, ang, rpos, rrot, rvel, rang;
    int64_t N; int J; float k[4], w[4]; float inv3j, invj; float* reward; float* raw; int64_t raw_stride;
};

template <bool AOS>
__global__ void __launch_bounds__(IM_WARPS * 32) reward_kernel(const RewardArgs a) {
    __shared__ __align__(16) float s_rec[AOS ? IM_WARPS * IM_REC : 4];
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * IM_WARPS + (threadIdx.x >> 5);
    if (n >= a.N) return;
    const float* rec = s_rec + (AOS ? (threadIdx.x >> 5) * IM_REC : 0);
    BodyState r{};                                   // reference loads first: they fly while the record is staged
    if (lane < a.J) r = BodyState{ld3(at(a.rpos, n, lane)), ld4(at(a.rrot, n, lane)), ld3(at(a.rvel, n, lane)), ld3(at(a.rang, n, lane))};
    if (AOS) stage_record(a.pos.ptr + n * a.pos.stride_env, a.J * REC, const_cast<float*>(rec), lane);
    float sp = 0.0f, sr = 0.0f, sv = 0.0f, sa = 0.0f;
    if (lane < a.J) {
        const int j = lane;
        const BodyState b = AOS ? body_from_record(rec, j)
                                : BodyState{ld3(at(a.pos, n, j)), ld4(at(a.rot, n, j)), ld3(at(a.vel, n, j)), ld3(at(a.ang, n, j))};
        reward_terms_body_fma(b, r, sp, sr, sv, sa);           // sums of squares + closed-form squared angle, as in the fused step
    }
    sp = warp_sum(sp); sr = warp_sum(sr); sv = warp_sum(sv); sa = warp_sum(sa);
    // env-level tail (common.py:300-320): lane v < 4 evaluates exponential kernel v, lane 0 collects -- four expf side by side
    float val = lane == 0 ? sp * a.inv3j : lane == 1 ? sr * a.invj : lane == 2 ? sv * a.inv3j : sa * a.inv3j;
    val = expf(-a.k[lane & 3] * val);
    const float r0 = __shfl_sync(FULL, val, 0), r1 = __shfl_sync(FULL, val, 1), r2 = __shfl_sync(FULL, val, 2), r3 = __shfl_sync(FULL, val, 3);
    if (lane == 0) {
        a.reward[n] = ((a.w[0] * r0 + a.w[1] * r1) + a.w[2] * r2) + a.w[3] * r3;
        float* o = a.raw + n * a.raw_stride;
        o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
    }
}

struct ResetArgs {
    const int16_t* progress; KView pos, rpos; const uint8_t* pass_time; int early; const float* term_dist; int use_mean;
    int64_t N; int J; uint8_t* reset; uint8_t* terminated; int dev;
};

template <bool AOS>      // AOS: rigid_body_pos is the pos slice of the 13-float records (stride_body 13): stage the span coalesced
__global__ void __launch_bounds__(IM_WARPS * 32) reset_kernel(const ResetArgs a) {
    __shared__ __align__(16) float s_rec[AOS ? IM_WARPS * IM_REC : 4];
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * IM_WARPS + (threadIdx.x >> 5);
    if (n >= a.N) return;
    bool fallen = false;
    if (a.early) {
        const float* rec = s_rec + (AOS ? (threadIdx.x >> 5) * IM_REC : 0);
        if (AOS) stage_record(a.pos.ptr + n * a.pos.stride_env, (a.J - 1) * REC + 3, const_cast<float*>(rec), lane);
        float d = 0.0f;
        bool over = false;
        if (lane < a.J) {
            d = norm3((AOS ? ld3(rec + REC * lane) : ld3(at(a.pos, n, lane))) - ld3(at(a.rpos, n, lane)), a.dev);
            over = d > __ldg(a.term_dist + (a.use_mean ? 0 : lane));
        }
        if (a.use_mean) fallen = warp_mean(d, a.J, lane, a.dev) > __ldg(a.term_dist);   // common.py:342-346
        else fallen = __any_sync(FULL, over);                                          // common.py:347-350
        fallen = fallen && (a.progress[n] > 1);                                        // common.py:354
    }
    if (lane == 0) {
        a.terminated[n] = fallen ? 1 : 0;                                              // common.py:356
        a.reset[n] = a.pass_time[n] ? 1 : (fallen ? 1 : 0);                            // common.py:362
    }
}

// Evaluation metric of HumanoidPHC.step (reference puffer_phc/envs/humanoid_phc.py:159-163):
// mpjpe = (body_pos - rg_pos).norm(dim=-1).mean(dim=-1).  One warp per env, lane = body.
struct MpjpeArgs { KView pos, rpos; int64_t N; int J; float* out; int dev; };

__global__ void __launch_bounds__(IM_WARPS * 32) mpjpe_kernel(const MpjpeArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * IM_WARPS + (threadIdx.x >> 5);
    if (n >= a.N) return;
    float d = 0.0f;
    if (lane < a.J) d = norm3(ld3(at(a.pos, n, lane)) - ld3(at(a.rpos, n, lane)), a.dev);
    d = warp_mean(d, a.J, lane, a.dev);
    if (lane == 0) a.out[n] = d;
}

// build_amp_observations_smpl + dof_to_obs_smpl (reference envs/common.py:179-267), "next" row f3.
// One warp per env: lane j = dof joint j of the subset (exp-map -> quaternion -> tan-norm, and the dof velocity copy);
// the root terms and the key-body positions are spread over the first lanes.
struct AmpArgs {
    const float *root_pos, *root_rot, *root_vel, *root_ang, *dof_pos, *dof_vel, *key_pos;
    const int64_t* subset; int nj, K, local_root_obs, root_height_obs, upright; int64_t N; float* obs; int64_t obs_stride; int dev;
    int hist, hist_w;      // history variant: rows per env and floats per row
};

// HIST > 0: obs is the history buffer _amp_obs_buf [N, HIST, W] (row stride obs_stride / HIST): the warp first shifts the env's rows
// one step back in place (_update_hist_amp_obs, humanoid_phc.py:1339-1348: hist[:, k] = buf[:, k-1]; a lane owns a column, so
// walking the rows from the oldest to the newest needs no synchronisation), then writes the current observation into row 0
// (_compute_amp_observations, :1123-1174) -- one pass instead of clone + copy + compute + copy.
template <bool HISTORY>
__global__ void __launch_bounds__(IM_WARPS * 32) amp_obs_kernel(const AmpArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * IM_WARPS + (threadIdx.x >> 5);
    if (n >= a.N) return;
    if (HISTORY) {
        float* buf = a.obs + n * a.obs_stride;
        const int W = a.hist_w;
        for (int c = lane; c < W; c += 32) {      // all loads of a column first (independent, in flight together), then the stores
            float v[15];
#pragma unroll
            for (int k = 0; k < 15; ++k)
                if (k < a.hist - 1) v[k] = buf[k * W + c];
#pragma unroll
            for (int k = 0; k < 15; ++k)
                if (k < a.hist - 1) buf[(k + 1) * W + c] = v[k];
        }
        __syncwarp();
    }
    Q4 rr = ld4(a.root_rot + n * 4);
    if (!a.upright) rr = remove_base_rot(rr);                                   // common.py:214-215
    float hz, hw;
    heading_quat(calc_heading(rr), hz, hw);                                     // h^-1 = (0,0,-hz,hw)  (:216)
    const V3 rp = ld3(a.root_pos + n * 3);
    float* o = a.obs + n * a.obs_stride;
    if (a.root_height_obs) { if (lane == 0) o[0] = rp.z; o += 1; }              // :213, 250-251
    if (lane == 0) tan_norm(a.local_root_obs ? quat_mul(Q4{0.0f, 0.0f, -hz, hw}, rr) : rr, o);   // :218-223
    if (lane == 1) put3(o + 6, rotate_z(-hz, hw, ld3(a.root_vel + n * 3)));     // :225
    if (lane == 2) put3(o + 9, rotate_z(-hz, hw, ld3(a.root_ang + n * 3)));     // :226
    float* dobs = o + 12;
    float* dvel = dobs + 6 * a.nj;
    float* kp = dvel + 3 * a.nj;
    for (int j = lane; j < a.nj; j += 32) {
        float e[3], v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t idx = a.subset ? __ldg(a.subset + 3 * j + c) : 3 * j + c;          // :244-246
            e[c] = __ldg(a.dof_pos + n * NDOF + idx);
            v[c] = __ldg(a.dof_vel + n * NDOF + idx);
        }
        tan_norm(exp_map_to_quat(V3{e[0], e[1], e[2]}, a.dev), dobs + 6 * j);                        // :186, 248
        dvel[3 * j] = v[0]; dvel[3 * j + 1] = v[1]; dvel[3 * j + 2] = v[2];
    }
    for (int k = lane; k < a.K; k += 32)                                                      // :228-242
        put3(kp + 3 * k, rotate_z(-hz, hw, ld3(a.key_pos + (n * a.K + k) * 3) - rp));
}

static int check_view(const char* fn, const char* name, const phc_view& v) {
    if (!v.ptr) return fail(PHC_EINVAL, "%s: %s is NULL", fn, name);
    if (v.stride_body >= (1 << 26) || v.stride_body <= -(1 << 26))
        return fail(PHC_EUNSUPPORTED, "%s: %s.stride_body=%lld (bodies of one env must lie within 2^26 floats of each other)", fn, name, (long long)v.stride_body);
    return PHC_OK;
}
static KView kv(const phc_view& v) { return KView{v.ptr, v.stride_env, (int)v.stride_body}; }

// the four views are the pos | rot | vel | ang-vel slices of one 13-float record per body (humanoid_phc.py:546-549)
static bool is_aos_record(const phc_view& p, const phc_view& r, const phc_view& v, const phc_view& w) {
    return p.stride_body == REC && r.stride_body == REC && v.stride_body == REC && w.stride_body == REC &&
           r.stride_env == p.stride_env && v.stride_env == p.stride_env && w.stride_env == p.stride_env &&
           r.ptr == p.ptr + 3 && v.ptr == p.ptr + 7 && w.ptr == p.ptr + 10;
}

}  // namespace phc

using namespace phc;

#define CHECK_VIEW(fn, v) do { int rc_ = check_view(fn, #v, v); if (rc_) return rc_; } while (0)

extern "C" int phc_imitation_obs_v6(phc_view root_pos, phc_view root_rot, phc_view body_pos, phc_view body_rot,
                                    phc_view body_vel, phc_view body_ang_vel, phc_view ref_body_pos, phc_view ref_body_rot,
                                    phc_view ref_body_vel, phc_view ref_body_ang_vel, int64_t N, int J, int time_steps,
                                    int upright, float* obs, int64_t obs_stride, phc_stream_t stream) {
    const char* fn = "phc_imitation_obs_v6";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(time_steps >= 1, PHC_EINVAL, "%s: time_steps=%d < 1", fn, time_steps);
    PHC_REQUIRE(J >= 1 && J <= 32, PHC_ESHAPE, "%s: J=%d outside [1,32]", fn, J);
    if (N == 0) return PHC_OK;
    CHECK_VIEW(fn, root_pos); CHECK_VIEW(fn, root_rot); CHECK_VIEW(fn, body_pos); CHECK_VIEW(fn, body_rot);
    CHECK_VIEW(fn, body_vel); CHECK_VIEW(fn, body_ang_vel); CHECK_VIEW(fn, ref_body_pos); CHECK_VIEW(fn, ref_body_rot);
    CHECK_VIEW(fn, ref_body_vel); CHECK_VIEW(fn, ref_body_ang_vel);
    PHC_REQUIRE(obs, PHC_EINVAL, "%s: obs is NULL", fn);
    PHC_REQUIRE(obs_stride >= (int64_t)24 * J * time_steps, PHC_ESHAPE, "%s: obs_stride=%lld < 24*J*time_steps", fn, (long long)obs_stride);
    ObsArgs a{kv(root_pos), kv(root_rot), kv(body_pos), kv(body_rot), kv(body_vel), kv(body_ang_vel), kv(ref_body_pos), kv(ref_body_rot),
              kv(ref_body_vel), kv(ref_body_ang_vel), N, J, upright, obs, obs_stride, time_steps};
    const unsigned grid = (unsigned)((N * time_steps + IM_WARPS - 1) / IM_WARPS);
    if (is_aos_record(body_pos, body_rot, body_vel, body_ang_vel)) imitation_obs_kernel<true><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    else imitation_obs_kernel<false><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_self_obs_smpl_max(phc_view body_pos, phc_view body_rot, phc_view body_vel, phc_view body_ang_vel,
                                     int64_t N, int J, int local_root_obs, int root_height_obs, int upright, float* obs,
                                     int64_t obs_stride, phc_stream_t stream) {
    const char* fn = "phc_self_obs_smpl_max";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(J >= 1 && J <= 32, PHC_ESHAPE, "%s: J=%d outside [1,32]", fn, J);
    if (N == 0) return PHC_OK;
    CHECK_VIEW(fn, body_pos); CHECK_VIEW(fn, body_rot); CHECK_VIEW(fn, body_vel); CHECK_VIEW(fn, body_ang_vel);
    PHC_REQUIRE(obs, PHC_EINVAL, "%s: obs is NULL", fn);
    PHC_REQUIRE(obs_stride >= (root_height_obs ? 1 : 0) + 3 * (J - 1) + 12 * J, PHC_ESHAPE, "%s: obs_stride too small", fn);
    SelfArgs a{kv(body_pos), kv(body_rot), kv(body_vel), kv(body_ang_vel), N, J, local_root_obs, root_height_obs, upright, obs, obs_stride};
    const unsigned grid = (unsigned)((N + IM_WARPS - 1) / IM_WARPS);
    if (is_aos_record(body_pos, body_rot, body_vel, body_ang_vel)) self_obs_kernel<true><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    else self_obs_kernel<false><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_amp_obs_smpl(const float* root_pos, const float* root_rot, const float* root_vel, const float* root_ang_vel,
                                const float* dof_pos, const float* dof_vel, const float* key_body_pos, const int64_t* dof_subset,
                                int num_joints, int K, int local_root_obs, int root_height_obs, int upright, int64_t N, float* obs,
                                int64_t obs_stride, int ref_device, phc_stream_t stream) {
    const char* fn = "phc_amp_obs_smpl";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(num_joints >= 0 && num_joints <= 23 && K >= 0, PHC_ESHAPE, "%s: num_joints=%d K=%d out of range", fn, num_joints, K);
    if (N == 0) return PHC_OK;
    PHC_REQUIRE(root_pos && root_rot && root_vel && root_ang_vel && dof_pos && dof_vel && (key_body_pos || K == 0) && obs, PHC_EINVAL,
                "%s: NULL pointer", fn);
    PHC_REQUIRE(obs_stride >= (root_height_obs ? 1 : 0) + 12 + 9 * num_joints + 3 * K, PHC_ESHAPE, "%s: obs_stride too small", fn);
    AmpArgs a{root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, dof_subset, num_joints, K, local_root_obs,
              root_height_obs, upright, N, obs, obs_stride, ref_device, 0, 0};
    amp_obs_kernel<false><<<(unsigned)((N + IM_WARPS - 1) / IM_WARPS), IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_amp_obs_hist_step(const float* root_pos, const float* root_rot, const float* root_vel, const float* root_ang_vel,
                                     const float* dof_pos, const float* dof_vel, const float* key_body_pos, const int64_t* dof_subset,
                                     int num_joints, int K, int local_root_obs, int root_height_obs, int upright, int64_t N,
                                     float* amp_obs_buf, int num_steps, int row_width, int ref_device, phc_stream_t stream) {
    const char* fn = "phc_amp_obs_hist_step";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(num_joints >= 0 && num_joints <= 23 && K >= 0, PHC_ESHAPE, "%s: num_joints=%d K=%d out of range", fn, num_joints, K);
    PHC_REQUIRE(num_steps >= 2 && num_steps <= 16, PHC_ESHAPE, "%s: num_steps=%d outside [2,16]", fn, num_steps);
    if (N == 0) return PHC_OK;
    PHC_REQUIRE(root_pos && root_rot && root_vel && root_ang_vel && dof_pos && dof_vel && (key_body_pos || K == 0) && amp_obs_buf, PHC_EINVAL,
                "%s: NULL pointer", fn);
    PHC_REQUIRE(row_width >= (root_height_obs ? 1 : 0) + 12 + 9 * num_joints + 3 * K, PHC_ESHAPE, "%s: row_width too small", fn);
    AmpArgs a{root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, dof_subset, num_joints, K, local_root_obs,
              root_height_obs, upright, N, amp_obs_buf, (int64_t)num_steps * row_width, ref_device, num_steps, row_width};
    amp_obs_kernel<true><<<(unsigned)((N + IM_WARPS - 1) / IM_WARPS), IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_imitation_reward(phc_view body_pos, phc_view body_rot, phc_view body_vel, phc_view body_ang_vel,
                                    phc_view ref_body_pos, phc_view ref_body_rot, phc_view ref_body_vel,
                                    phc_view ref_body_ang_vel, int64_t N, int J, const float* k_h, const float* w_h,
                                    float* reward, float* reward_raw, int64_t raw_stride, phc_stream_t stream) {
    const char* fn = "phc_imitation_reward";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(J >= 1 && J <= 32, PHC_ESHAPE, "%s: J=%d outside [1,32]", fn, J);
    if (N == 0) return PHC_OK;
    CHECK_VIEW(fn, body_pos); CHECK_VIEW(fn, body_rot); CHECK_VIEW(fn, body_vel); CHECK_VIEW(fn, body_ang_vel);
    CHECK_VIEW(fn, ref_body_pos); CHECK_VIEW(fn, ref_body_rot); CHECK_VIEW(fn, ref_body_vel); CHECK_VIEW(fn, ref_body_ang_vel);
    PHC_REQUIRE(k_h && w_h && reward && reward_raw, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(raw_stride >= 4, PHC_ESHAPE, "%s: raw_stride < 4", fn);
    RewardArgs a{kv(body_pos), kv(body_rot), kv(body_vel), kv(body_ang_vel), kv(ref_body_pos), kv(ref_body_rot), kv(ref_body_vel),
                 kv(ref_body_ang_vel), N, J,
                 {k_h[0], k_h[1], k_h[2], k_h[3]}, {w_h[0], w_h[1], w_h[2], w_h[3]}, 1.0f / (3.0f * (float)J), 1.0f / (float)J, reward,
                 reward_raw, raw_stride};
    const unsigned grid = (unsigned)((N + IM_WARPS - 1) / IM_WARPS);
#ifndef IM_REWARD_AOS
#define IM_REWARD_AOS 0          // A/B builds: 1 = stage the AoS record through shared memory as the observation kernels do
#endif
    if (IM_REWARD_AOS && is_aos_record(body_pos, body_rot, body_vel, body_ang_vel)) reward_kernel<true><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    else reward_kernel<false><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_im_reset(const int16_t* progress, phc_view rigid_body_pos, phc_view ref_body_pos, const uint8_t* pass_time,
                            int enable_early_termination, const float* termination_distance, int use_mean, int64_t N, int J,
                            uint8_t* reset, uint8_t* terminated, int ref_device, phc_stream_t stream) {
    const char* fn = "phc_im_reset";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(J >= 1 && J <= 32, PHC_ESHAPE, "%s: J=%d outside [1,32]", fn, J);
    if (N == 0) return PHC_OK;
    CHECK_VIEW(fn, rigid_body_pos); CHECK_VIEW(fn, ref_body_pos);
    PHC_REQUIRE(progress && pass_time && reset && terminated, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(!enable_early_termination || termination_distance, PHC_EINVAL, "%s: termination_distance is NULL", fn);
    PHC_REQUIRE(ref_device == PHC_REF_DEVICE_CPU || ref_device == PHC_REF_DEVICE_CUDA, PHC_EINVAL, "%s: ref_device must be 0 or 1", fn);
    ResetArgs a{progress, kv(rigid_body_pos), kv(ref_body_pos), pass_time, enable_early_termination, termination_distance, use_mean,
                N, J, reset, terminated, ref_device};
    const unsigned grid = (unsigned)((N + IM_WARPS - 1) / IM_WARPS);
    // (staging the whole record span measured slower -- 27.8 vs 17.2 us: the strided loads touch only the position sectors)
    reset_kernel<false><<<grid, IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_mpjpe(phc_view body_pos, phc_view ref_body_pos, int64_t N, int J, float* mpjpe, int ref_device, phc_stream_t stream) {
    const char* fn = "phc_mpjpe";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(J >= 1 && J <= 32, PHC_ESHAPE, "%s: J=%d outside [1,32]", fn, J);
    if (N == 0) return PHC_OK;
    CHECK_VIEW(fn, body_pos); CHECK_VIEW(fn, ref_body_pos);
    PHC_REQUIRE(mpjpe, PHC_EINVAL, "%s: NULL pointer", fn);
    MpjpeArgs a{kv(body_pos), kv(ref_body_pos), N, J, mpjpe, ref_device};
    mpjpe_kernel<<<(unsigned)((N + IM_WARPS - 1) / IM_WARPS), IM_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}
