// step_fused.cu -- the whole post-physics half of HumanoidPHC.step in one pass over HBM
// (reference puffer_phc/envs/humanoid_phc.py:136-149): two get_motion_state queries (t for reward /
// reset, t+1 for the task observation; motion_lib.py:549-626), compute_imitation_reward (+ power term,
// humanoid_phc.py:1295-1303), compute_humanoid_im_reset, compute_humanoid_observations_smpl_max,
// compute_imitation_observations_v6, and optionally RunningNorm.forward plus the column moments that
// RunningNorm.update needs (policies/running_norm.py:15-34).
//
// Mapping: persistent CTAs of 16 warps; each iteration a CTA owns 8 consecutive envs and TWO warps work on
// each env, lane j = body j:
//   role A (warps 0-7):  heading quaternion (handed to role B through shared memory + a named barrier),
//                        reference state at t  -> reward terms, termination test, power term, self observation
//   role B (warps 8-15): reference state at t+1 -> imitation (task) observation
// Splitting the env halves the per-warp dependency chain and the live register state (2 frames per warp
// instead of 4), which is what bounds this kernel (it is latency-, not bandwidth-limited at 1 warp/env).
// The 8 PhysX records (8 x 1248 B) of the NEXT iteration are prefetched with cp.async while the current
// tile is written out.  The 8 x 934-float observation tile is assembled in shared memory and leaves the SM
// as one contiguous, 16-byte aligned 29.9 KB block (float4 stores); the normalised copy and the fp64 column
// moments are produced from the same tile by column-owning threads (mean / 1/sqrt(var+eps) in registers).
#include "phc_body.cuh"

namespace phc {

#ifndef ST_ENVS_PER_CTA
#define ST_ENVS_PER_CTA 8
#endif
constexpr int ST_ENVS = ST_ENVS_PER_CTA;         // envs per CTA iteration (4 or 8; tiles of 4+ rows stay 16-byte aligned)
constexpr int ST_WARPS = 2 * ST_ENVS;            // two warps (roles A, B) per env
constexpr int ST_THREADS = ST_WARPS * 32;        // 512
#ifndef ST_MIN_CTAS
#define ST_MIN_CTAS 2
#endif
constexpr int ST_COLS_PER_THREAD = (OBS_W + ST_THREADS - 1) / ST_THREADS;   // 2

struct StepArgs {
    phc_motion_tables t;
    phc_step_in in;
    phc_step_cfg cfg;
    phc_step_out out;
    int sim_vec;          // body_state rows are 16-byte aligned -> 16-byte cp.async staging
    int obs_vec;          // obs tiles are contiguous and 16-byte aligned -> float4 tile stores
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// named barriers 1..8: one per env slot, 64 threads (the slot's role-A and role-B warps)
__device__ __forceinline__ void pair_arrive(int slot) { asm volatile("bar.arrive %0, 64;" ::"r"(slot + 1) : "memory"); }
__device__ __forceinline__ void pair_sync(int slot) { asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory"); }
// torch.clamp propagates NaN: min.NaN / max.NaN do too (fminf / fmaxf would drop it)
__device__ __forceinline__ float clamp_nan(float y, float lim) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(y), "f"(-lim));
    asm("min.NaN.f32 %0, %0, %1;" : "+f"(r) : "f"(lim));
    return r;
}

// stage the PhysX records (24 x 13 floats each) of the 8 envs of block `blk` into shared memory (async)
__device__ __forceinline__ void stage_sim(const StepArgs& a, int64_t blk, float* sim, int tid) {
    const int64_t e0 = blk * ST_ENVS;
    const int rows = (int)((a.in.N - e0 < ST_ENVS) ? (a.in.N - e0) : ST_ENVS);
    if (a.sim_vec) {
        for (int i = tid; i < rows * (SIM_F / 4); i += ST_THREADS) {
            const int r = i / (SIM_F / 4), c = i - r * (SIM_F / 4);
            cp_async16(sim + r * SIM_F + 4 * c, a.in.body_state + (e0 + r) * a.in.env_stride + 4 * c);
        }
    } else {
        for (int i = tid; i < rows * SIM_F; i += ST_THREADS) {
            const int r = i / SIM_F, c = i - r * SIM_F;
            cp_async4(sim + r * SIM_F + c, a.in.body_state + (e0 + r) * a.in.env_stride + c);
        }
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(ST_THREADS, ST_MIN_CTAS) step_fused_kernel(const StepArgs a) {
    extern __shared__ float4 smem4[];
    float* tile = reinterpret_cast<float*>(smem4);                 // [8][934]
    float* sim = tile + ST_ENVS * OBS_W;                           // [8][312]
    float* s_head = sim + ST_ENVS * SIM_F;                         // [8][2]  heading quaternion (z, w) per env slot

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int role = warp / ST_ENVS, slot = warp % ST_ENVS;                   // each SM sub-partition gets two A and two B warps
    const phc_motion_tables& T = a.t;
    const phc_step_in& in = a.in;
    const phc_step_cfg& cfg = a.cfg;
    const phc_step_out& out = a.out;
    const bool do_norm = out.obs_norm != nullptr;
    const bool do_mom = out.moment_partials != nullptr;

    // column-owning state for the tile phase: thread tid owns observation columns tid and tid + 512
    float c_mean[ST_COLS_PER_THREAD], c_inv[ST_COLS_PER_THREAD];
    double msum[ST_COLS_PER_THREAD], msq[ST_COLS_PER_THREAD];
#pragma unroll
    for (int u = 0; u < ST_COLS_PER_THREAD; ++u) {
        const int c = tid + u * ST_THREADS;
        msum[u] = 0.0; msq[u] = 0.0; c_mean[u] = 0.0f; c_inv[u] = 1.0f;
        if (do_norm && c < OBS_W) {
            c_mean[u] = __ldg(in.rms_mean + c);
            c_inv[u] = 1.0f / sqrtf(__ldg(in.rms_var + c) + cfg.rms_eps);   // running_norm.py:17, one IEEE reciprocal per column
        }
    }

    float* my_tile = tile + slot * OBS_W;
    const float* my_sim = sim + slot * SIM_F;

    if ((int64_t)blockIdx.x < a.num_blocks) stage_sim(a, blockIdx.x, sim, tid);
    cp_async_wait_all();
    __syncthreads();

    for (int64_t blk = blockIdx.x; blk < a.num_blocks; blk += gridDim.x) {
        const int64_t e = blk * ST_ENVS + slot;
        if (e < in.N) {
            // ---- per-env scalars (all lanes read the same addresses) ---------------------------------
            const int64_t id = __ldg(in.motion_ids + e);
            const int16_t prog = __ldg(in.progress + e);
            const float st = __ldg(in.start_time + e), so = __ldg(in.start_offset + e);
            const float mlen = __ldg(T.motion_len + id), mdt = __ldg(T.motion_dt + id);
            const int64_t nf = __ldg(T.num_frames + id), ls = __ldg(T.length_starts + id);
            const V3 off = ldg3(in.global_offset + e * 3);
            const V3 root_p = ld3(my_sim);
            const int j = lane;

            if (role == 0) {
                // ================= role A: reference at t -> reward, reset, power, self observation ===========
                const Q4 root_q = ld4(my_sim + 3);
                float hz, hw;
                heading_quat(calc_heading(root_q), hz, hw);      // upright start: no base-rot removal
                if (lane == 0) { s_head[2 * slot] = hz; s_head[2 * slot + 1] = hw; }
                __threadfence_block();
                pair_arrive(slot);                               // role B picks the heading up with pair_sync

                const float t0 = ((float)prog * cfg.dt + st) + so;                        // humanoid_phc.py:1233-1235
                int64_t a0, a1;
                float bla;
                frame_blend(t0, mlen, nf, mdt, a0, a1, bla);
                float sp = 0.0f, sr = 0.0f, sv = 0.0f, sa = 0.0f, dist = 0.0f;
                bool over = false;
                const bool in_mask = lane < NB && ((cfg.reset_body_mask >> lane) & 1u);
                if (lane < NB) {
                    const BodyState A0 = load_frame<PACKED>(T, a0 + ls, j);
                    const BodyState A1 = (a1 == a0) ? A0 : load_frame<PACKED>(T, a1 + ls, j);
                    const float* sj = my_sim + REC * j;
                    const BodyState body{ld3(sj), ld4(sj + 3), ld3(sj + 7), ld3(sj + 10)};
                    float* o = my_tile;
                    if (j == 0) o[0] = root_p.z;                                          // common.py:40
                    self_obs_body(body, root_p, hz, hw, j, o + 1 + 3 * (j - 1), o + 70 + 6 * j, o + 214 + 3 * j, o + 286 + 3 * j);
                    const BodyState r0 = blend_frames(A0, A1, bla, off);
                    reward_terms_body_fast(body, r0, sp, sr, sv, sa);
                    if (in_mask) {
                        dist = norm3(body.p - r0.p);
                        over = dist > __ldg(in.term_dist + j);
                    }
                    if (out.ref_state_t) store_ref(out.ref_state_t + e * FRAME_F, j, r0);
                }
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
                    fallen = fallen && (prog > 1);                                        // common.py:354
                }
                float power = 0.0f;
                if (in.dof_force) {                                                       // humanoid_phc.py:1295-1303
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int c = lane + 32 * k;
                        if (c < NDOF) power = power + fabsf(__ldg(in.dof_force + e * NDOF + c) * __ldg(in.dof_vel + e * NDOF + c));
                    }
                    power = warp_sum(power);
                }
                if (lane == 0) {
                    float raw[4];
                    float rew = reward_from_sq_sums(sp, sr, sv, sa, (float)NB, cfg.k, cfg.w, raw);
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
                    out.reset[e] = (t0 >= mlen) ? 1 : (fallen ? 1 : 0);                   // humanoid_phc.py:1315, common.py:362
                }
            } else {
                // ================= role B: reference at t+1 -> imitation observation ==========================
                const float t1 = ((float)(int16_t)(prog + 1) * cfg.dt + st) + so;         // humanoid_phc.py:1060-1064 (int16 + 1)
                int64_t b0, b1;
                float blb;
                frame_blend(t1, mlen, nf, mdt, b0, b1, blb);
                BodyState body, r1;
                if (lane < NB) {
                    const BodyState B0 = load_frame<PACKED>(T, b0 + ls, j);
                    const BodyState B1 = (b1 == b0) ? B0 : load_frame<PACKED>(T, b1 + ls, j);
                    const float* sj = my_sim + REC * j;
                    body = BodyState{ld3(sj), ld4(sj + 3), ld3(sj + 7), ld3(sj + 10)};
                    r1 = blend_frames(B0, B1, blb, off);
                    if (out.ref_state_t1) store_ref(out.ref_state_t1 + e * FRAME_F, j, r1);
                }
                pair_sync(slot);
                const float hz = s_head[2 * slot], hw = s_head[2 * slot + 1];
                if (lane < NB) {
                    float* q = my_tile + OBS_SELF;
                    task_obs_body(body, r1, root_p, hz, hw, q + 3 * j, q + 72 + 6 * j, q + 216 + 3 * j, q + 288 + 3 * j,
                                  q + 360 + 3 * j, q + 432 + 6 * j);
                }
            }
        }
        __syncthreads();     // the tile is complete and the sim records of this iteration are no longer needed

        // ---- prefetch the next iteration's PhysX records while the tile is written out ----------------
        const int64_t nxt = blk + gridDim.x;
        if (nxt < a.num_blocks) stage_sim(a, nxt, sim, tid);

        // ---- the CTA's 8 x 934 tile leaves as one contiguous block ------------------------------------
        const int64_t e0 = blk * ST_ENVS;
        const int rows = (int)((in.N - e0 < ST_ENVS) ? (in.N - e0) : ST_ENVS);
        if (a.obs_vec) {
            const int total = rows * OBS_W;
            float4* dst = reinterpret_cast<float4*>(out.obs + e0 * OBS_W);
            const float4* src = reinterpret_cast<const float4*>(tile);
            const int n4 = total >> 2;
            for (int i = tid; i < n4; i += ST_THREADS) dst[i] = src[i];
            for (int i = (n4 << 2) + tid; i < total; i += ST_THREADS) out.obs[e0 * OBS_W + i] = tile[i];
        }
        if (!a.obs_vec || do_norm || do_mom) {
#pragma unroll
            for (int u = 0; u < ST_COLS_PER_THREAD; ++u) {
                const int c = tid + u * ST_THREADS;
                if (c < OBS_W) {
                    for (int r = 0; r < rows; ++r) {
                        const float x = tile[r * OBS_W + c];
                        if (!a.obs_vec) out.obs[(e0 + r) * out.obs_stride + c] = x;
                        if (do_norm)   // (x - mean) / sqrt(var + eps) as a multiplication by the column's reciprocal (<= 1.5 ulp)
                            out.obs_norm[(e0 + r) * out.obs_stride + c] = clamp_nan((x - c_mean[u]) * c_inv[u], cfg.rms_clip);
                        if (do_mom) {
                            const double xd = (double)x;
                            msum[u] += xd;
                            msq[u] = fma(xd, xd, msq[u]);          // xd*xd is exact in fp64, so this equals msq + xd*xd
                        }
                    }
                }
            }
        }
        cp_async_wait_all();
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

constexpr size_t ST_SMEM = (size_t)(ST_ENVS * OBS_W + ST_ENVS * SIM_F + 2 * ST_ENVS) * sizeof(float);

}  // namespace phc

using namespace phc;

extern "C" int phc_step_num_partials(void) { return ST_MIN_CTAS * sm_count(); }

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
    // the grid is fixed (ST_MIN_CTAS CTAs per SM) so that the number of moment partial slots does not depend on N
    const int grid = phc_step_num_partials();
    StepArgs a{*t, *in, *cfg, *out, 0, 0, (in->N + ST_ENVS - 1) / ST_ENVS};
    a.sim_vec = aligned16(in->body_state) && (in->env_stride % 4 == 0);
    a.obs_vec = out->obs_stride == OBS_W && aligned16(out->obs);
    if (in->N == 0 && !out->moment_partials) return PHC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (packed) step_fused_kernel<true><<<grid, ST_THREADS, ST_SMEM, s>>>(a);
    else step_fused_kernel<false><<<grid, ST_THREADS, ST_SMEM, s>>>(a);
    return check_launch(fn);
}
