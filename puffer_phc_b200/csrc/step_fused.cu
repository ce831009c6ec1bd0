// step_fused.cu -- the whole post-physics half of HumanoidPHC.step in one pass over HBM
// (reference puffer_phc/envs/humanoid_phc.py:136-149): two get_motion_state queries (t for reward /
// reset, t+1 for the task observation; motion_lib.py:549-626), compute_imitation_reward (+ power term,
// humanoid_phc.py:1295-1303), compute_humanoid_im_reset, compute_humanoid_observations_smpl_max,
// compute_imitation_observations_v6, and optionally RunningNorm.forward plus the column moments that
// RunningNorm.update needs (policies/running_norm.py:15-34).
//
// Mapping: persistent CTAs of 8 warps; each iteration a CTA owns 8 consecutive envs, one warp per env,
// lane j = body j.  The env's 1248-byte PhysX record is staged in shared memory with coalesced float4
// loads; the reference frames are gathered straight into registers (frames shared between the t and
// t+1 queries are loaded once); body reductions are warp shuffles; the 8 x 934-float observation tile
// is assembled in shared memory and leaves the SM as one contiguous, 16-byte aligned 29.9 KB block
// (float4 stores), normalised copy and fp64 column moments are produced from the same tile.
#include "phc_body.cuh"

namespace phc {

constexpr int ST_WARPS = 8;
constexpr int ST_THREADS = ST_WARPS * 32;
constexpr int ST_COLS_PER_THREAD = (OBS_W + ST_THREADS - 1) / ST_THREADS;   // 4

struct StepArgs {
    phc_motion_tables t;
    phc_step_in in;
    phc_step_cfg cfg;
    phc_step_out out;
    int sim_vec;          // body_state rows are 16-byte aligned -> float4 staging
    int obs_vec;          // obs (and obs_norm) tiles are contiguous and 16-byte aligned -> float4 tile stores
    int64_t num_blocks;   // ceil(N / 8)
};

template <bool PACKED>
__device__ __forceinline__ BodyState load_frame(const phc_motion_tables& T, int64_t f, int j) {
    BodyState s;
    if (PACKED) {
        const float* base = T.packed + f * FRAME_F;
        s.p = ldg3(base + 3 * j);
        s.q = ldg4a(base + 72 + 4 * j);
        s.v = ldg3(base + 168 + 3 * j);
        s.w = ldg3(base + 240 + 3 * j);
    } else {
        const int64_t r = f * NB + j;
        s.p = ldg3(T.gts + r * 3);
        s.q = ldg4a(T.grs + r * 4);
        s.v = ldg3(T.gvs + r * 3);
        s.w = ldg3(T.gavs + r * 3);
    }
    return s;
}

__device__ __forceinline__ void store_ref(float* dst, int j, const BodyState& r) {
    st3(dst + 3 * j, r.p);
    st4(dst + 72 + 4 * j, r.q);
    st3(dst + 168 + 3 * j, r.v);
    st3(dst + 240 + 3 * j, r.w);
}

template <bool PACKED>
__global__ void __launch_bounds__(ST_THREADS, 2) step_fused_kernel(const StepArgs a) {
    extern __shared__ float4 smem4[];
    float* tile = reinterpret_cast<float*>(smem4);                 // [8][934]
    float* sim = tile + ST_WARPS * OBS_W;                          // [8][312]
    float* s_mean = sim + ST_WARPS * SIM_F;                        // [934]   (only with obs_norm)
    float* s_den = s_mean + OBS_W;                                 // [934]   sqrt(var + eps)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const phc_motion_tables& T = a.t;
    const phc_step_in& in = a.in;
    const phc_step_cfg& cfg = a.cfg;
    const phc_step_out& out = a.out;
    const bool do_norm = out.obs_norm != nullptr;
    const bool do_mom = out.moment_partials != nullptr;

    if (do_norm) {
        for (int c = tid; c < OBS_W; c += ST_THREADS) {
            s_mean[c] = __ldg(in.rms_mean + c);
            s_den[c] = sqrtf(__ldg(in.rms_var + c) + cfg.rms_eps);          // running_norm.py:17
        }
    }
    double msum[ST_COLS_PER_THREAD], msq[ST_COLS_PER_THREAD];
#pragma unroll
    for (int u = 0; u < ST_COLS_PER_THREAD; ++u) { msum[u] = 0.0; msq[u] = 0.0; }

    float* my_tile = tile + warp * OBS_W;
    float* my_sim = sim + warp * SIM_F;

    for (int64_t blk = blockIdx.x; blk < a.num_blocks; blk += gridDim.x) {
        const int64_t e = blk * ST_WARPS + warp;
        if (e < in.N) {
            // ---- stage the env's PhysX record (24 x 13 floats) -------------------------------------
            const float* rec = in.body_state + e * in.env_stride;
            if (a.sim_vec) {
                const float4* src = reinterpret_cast<const float4*>(rec);
                float4* dst = reinterpret_cast<float4*>(my_sim);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = lane + 32 * k;
                    if (i < SIM_F / 4) dst[i] = __ldg(src + i);
                }
            } else {
                for (int i = lane; i < SIM_F; i += 32) my_sim[i] = __ldg(rec + i);
            }

            // ---- per-env scalars and the two frame-blend computations (all lanes, same addresses) ----
            const int64_t id = __ldg(in.motion_ids + e);
            const int16_t prog = __ldg(in.progress + e);
            const float st = __ldg(in.start_time + e), so = __ldg(in.start_offset + e);
            const float mlen = __ldg(T.motion_len + id), mdt = __ldg(T.motion_dt + id);
            const int64_t nf = __ldg(T.num_frames + id), ls = __ldg(T.length_starts + id);
            // humanoid_phc.py:1233-1235 and :1060-1064 ((progress_buf + 1) stays int16)
            const float t0 = ((float)prog * cfg.dt + st) + so;
            const float t1 = ((float)(int16_t)(prog + 1) * cfg.dt + st) + so;
            int64_t a0, a1, b0, b1;
            float bla, blb;
            frame_blend(t0, mlen, nf, mdt, a0, a1, bla);
            frame_blend(t1, mlen, nf, mdt, b0, b1, blb);
            const V3 off = ldg3(in.global_offset + e * 3);

            float sp = 0.0f, sr = 0.0f, sv = 0.0f, sa = 0.0f, dist = 0.0f;
            bool over = false;
            const bool in_mask = lane < NB && ((cfg.reset_body_mask >> lane) & 1u);
            __syncwarp();
            const Q4 root_q = ld4(my_sim + 3);
            const V3 root_p = ld3(my_sim);
            float hz, hw;
            heading_quat(calc_heading(root_q), hz, hw);          // upright start: no base-rot removal

            if (lane < NB) {
                const int j = lane;
                // ---- reference frames: gather, sharing frames between the t and t+1 queries ---------
                const BodyState A0 = load_frame<PACKED>(T, a0 + ls, j);
                const BodyState A1 = (a1 == a0) ? A0 : load_frame<PACKED>(T, a1 + ls, j);
                const BodyState B0 = (b0 == a1) ? A1 : ((b0 == a0) ? A0 : load_frame<PACKED>(T, b0 + ls, j));
                const BodyState B1 = (b1 == a1) ? A1 : ((b1 == b0) ? B0 : load_frame<PACKED>(T, b1 + ls, j));
                const float* sj = my_sim + REC * j;
                const BodyState body{ld3(sj), ld4(sj + 3), ld3(sj + 7), ld3(sj + 10)};

                // ---- reward and reset use the reference at t --------------------------------------
                const BodyState r0 = blend_frames(A0, A1, bla, off);
                reward_terms_body(body, r0, sp, sr, sv, sa);
                if (in_mask) {
                    dist = norm3(body.p - r0.p);
                    over = dist > __ldg(in.term_dist + j);
                }
                if (out.ref_state_t) store_ref(out.ref_state_t + e * FRAME_F, j, r0);

                // ---- observations use the reference at t+1 -----------------------------------------
                const BodyState r1 = blend_frames(B0, B1, blb, off);
                if (out.ref_state_t1) store_ref(out.ref_state_t1 + e * FRAME_F, j, r1);
                float* o = my_tile;
                if (j == 0) o[0] = root_p.z;                                              // common.py:40
                self_obs_body(body, root_p, hz, hw, j, o + 1 + 3 * (j - 1), o + 70 + 6 * j, o + 214 + 3 * j, o + 286 + 3 * j);
                float* q = my_tile + OBS_SELF;
                task_obs_body(body, r1, root_p, hz, hw, q + 3 * j, q + 72 + 6 * j, q + 216 + 3 * j, q + 288 + 3 * j,
                              q + 360 + 3 * j, q + 432 + 6 * j);
            } else {
                dist = 0.0f;
            }

            // ---- env-level reductions ---------------------------------------------------------------
            sp = warp_sum(sp); sr = warp_sum(sr); sv = warp_sum(sv); sa = warp_sum(sa);
            bool fallen = false;
            if (cfg.enable_early_termination) {
                if (cfg.use_mean) {
                    const float total = warp_sum(in_mask ? dist : 0.0f);
                    const int first = __ffs(cfg.reset_body_mask) - 1;
                    fallen = (total / (float)__popc(cfg.reset_body_mask & 0xffffffu)) > __ldg(in.term_dist + first);
                } else {
                    fallen = __any_sync(FULL, over);
                }
                fallen = fallen && (prog > 1);                                            // common.py:354
            }
            float power = 0.0f;
            if (in.dof_force) {                                                           // humanoid_phc.py:1295-1303
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = lane + 32 * k;
                    if (c < NDOF) power = power + fabsf(__ldg(in.dof_force + e * NDOF + c) * __ldg(in.dof_vel + e * NDOF + c));
                }
                power = warp_sum(power);
            }
            if (lane == 0) {
                float raw[4];
                float rew = reward_from_sums(sp, sr, sv, sa, (float)NB, cfg.k, cfg.w, raw);
                float* rr = out.reward_raw + e * out.raw_stride;
                rr[0] = raw[0]; rr[1] = raw[1]; rr[2] = raw[2]; rr[3] = raw[3];
                if (in.dof_force) {
                    float pr = -cfg.power_coef * power;
                    if (prog <= 3) pr = 0.0f;
                    rew = rew + pr;
                    rr[4] = pr;
                }
                out.reward[e] = rew;
                out.terminated[e] = fallen ? 1 : 0;
                out.reset[e] = (t0 >= mlen) ? 1 : (fallen ? 1 : 0);                       // humanoid_phc.py:1315, common.py:362
            }
        }
        __syncthreads();

        // ---- the CTA's 8 x 934 tile leaves as one contiguous block --------------------------------
        const int64_t e0 = blk * ST_WARPS;
        const int rows = (int)((in.N - e0 < ST_WARPS) ? (in.N - e0) : ST_WARPS);
        const int total = rows * OBS_W;
        if (a.obs_vec) {
            float4* dst = reinterpret_cast<float4*>(out.obs + e0 * OBS_W);
            const float4* src = reinterpret_cast<const float4*>(tile);
            const int n4 = total >> 2;
            for (int i = tid; i < n4; i += ST_THREADS) dst[i] = src[i];
            for (int i = (n4 << 2) + tid; i < total; i += ST_THREADS) out.obs[e0 * OBS_W + i] = tile[i];
            if (do_norm) {
                float4* dn = reinterpret_cast<float4*>(out.obs_norm + e0 * OBS_W);
                for (int i = tid; i < n4; i += ST_THREADS) {
                    float4 v = src[i];
                    int c = (i << 2) % OBS_W;
                    float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float y = (r[u] - s_mean[c]) / s_den[c];
                        r[u] = (y != y) ? y : fminf(fmaxf(y, -cfg.rms_clip), cfg.rms_clip);
                        c = (c + 1 == OBS_W) ? 0 : c + 1;
                    }
                    dn[i] = make_float4(r[0], r[1], r[2], r[3]);
                }
                for (int i = (n4 << 2) + tid; i < total; i += ST_THREADS) {
                    const int c = i % OBS_W;
                    float y = (tile[i] - s_mean[c]) / s_den[c];
                    out.obs_norm[e0 * OBS_W + i] = (y != y) ? y : fminf(fmaxf(y, -cfg.rms_clip), cfg.rms_clip);
                }
            }
        } else {
            for (int r = 0; r < rows; ++r)
                for (int c = tid; c < OBS_W; c += ST_THREADS) {
                    const float x = tile[r * OBS_W + c];
                    out.obs[(e0 + r) * out.obs_stride + c] = x;
                    if (do_norm) {
                        float y = (x - s_mean[c]) / s_den[c];
                        out.obs_norm[(e0 + r) * out.obs_stride + c] = (y != y) ? y : fminf(fmaxf(y, -cfg.rms_clip), cfg.rms_clip);
                    }
                }
        }
        if (do_mom) {
#pragma unroll
            for (int u = 0; u < ST_COLS_PER_THREAD; ++u) {
                const int c = tid + u * ST_THREADS;
                if (c < OBS_W)
                    for (int r = 0; r < rows; ++r) {
                        const double x = (double)tile[r * OBS_W + c];
                        msum[u] += x;
                        msq[u] += x * x;
                    }
            }
        }
        __syncthreads();
    }

    if (do_mom) {
        double* p = out.moment_partials + (int64_t)blockIdx.x * 2 * OBS_W;
#pragma unroll
        for (int u = 0; u < ST_COLS_PER_THREAD; ++u) {
            const int c = tid + u * ST_THREADS;
            if (c < OBS_W) { p[c] = msum[u]; p[OBS_W + c] = msq[u]; }
        }
    }
}

constexpr size_t ST_SMEM = (size_t)(ST_WARPS * OBS_W + ST_WARPS * SIM_F + 2 * OBS_W) * sizeof(float);

}  // namespace phc

using namespace phc;

extern "C" int phc_step_num_partials(void) { return 2 * sm_count(); }

extern "C" int phc_step_fused(const phc_motion_tables* t, const phc_step_in* in, const phc_step_cfg* cfg,
                              const phc_step_out* out, phc_stream_t stream) {
    const char* fn = "phc_step_fused";
    PHC_REQUIRE(t && in && cfg && out, PHC_EINVAL, "%s: NULL argument struct", fn);
    PHC_REQUIRE(in->N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    PHC_REQUIRE(in->body_state && in->progress && in->start_time && in->start_offset && in->motion_ids && in->global_offset &&
                    in->term_dist, PHC_EINVAL, "%s: a required input pointer is NULL", fn);
    PHC_REQUIRE(in->env_stride >= SIM_F, PHC_ESHAPE, "%s: env_stride=%lld < 312", fn, (long long)in->env_stride);
    PHC_REQUIRE((in->dof_force == nullptr) == (in->dof_vel == nullptr), PHC_EINVAL, "%s: dof_force and dof_vel must be given together", fn);
    PHC_REQUIRE(out->obs && out->reward && out->reward_raw && out->reset && out->terminated, PHC_EINVAL,
                "%s: a required output pointer is NULL", fn);
    PHC_REQUIRE(out->obs_stride >= OBS_W, PHC_ESHAPE, "%s: obs_stride=%lld < 934", fn, (long long)out->obs_stride);
    PHC_REQUIRE(out->raw_stride >= (in->dof_force ? 5 : 4), PHC_ESHAPE, "%s: raw_stride=%lld too small", fn, (long long)out->raw_stride);
    PHC_REQUIRE(!out->obs_norm || (in->rms_mean && in->rms_var), PHC_EINVAL, "%s: obs_norm needs rms_mean and rms_var", fn);
    PHC_REQUIRE((cfg->reset_body_mask & 0xffffffu) != 0 || !cfg->enable_early_termination, PHC_EINVAL, "%s: empty reset_body_mask", fn);
    PHC_REQUIRE(t->motion_len && t->motion_dt && t->num_frames && t->length_starts, PHC_EINVAL, "%s: per-motion tables missing", fn);
    const bool packed = t->packed != nullptr;
    if (packed) {
        PHC_REQUIRE(aligned16(t->packed), PHC_EALIGN, "%s: packed table must be 16-byte aligned", fn);
    } else {
        PHC_REQUIRE(t->gts && t->grs && t->gvs && t->gavs, PHC_EINVAL, "%s: gts/grs/gvs/gavs tables missing", fn);
        PHC_REQUIRE(aligned16(t->grs), PHC_EALIGN, "%s: grs table must be 16-byte aligned", fn);
    }
    // the grid is fixed (2 CTAs per SM) so that the number of moment partial slots does not depend on N
    const int grid = phc_step_num_partials();
    StepArgs a{*t, *in, *cfg, *out, 0, 0, (in->N + ST_WARPS - 1) / ST_WARPS};
    a.sim_vec = aligned16(in->body_state) && (in->env_stride % 4 == 0);
    a.obs_vec = out->obs_stride == OBS_W && aligned16(out->obs) && (!out->obs_norm || aligned16(out->obs_norm));
    if (in->N == 0 && !out->moment_partials) return PHC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (packed) step_fused_kernel<true><<<grid, ST_THREADS, ST_SMEM, s>>>(a);
    else step_fused_kernel<false><<<grid, ST_THREADS, ST_SMEM, s>>>(a);
    return check_launch(fn);
}
