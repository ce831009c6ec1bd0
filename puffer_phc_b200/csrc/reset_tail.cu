// reset_tail.cu -- the auto-reset that follows every env step, on the device and without a host round trip
// ("next" row f1 of SURVEY.md section 8).  Reference flow, all of it eager torch with a torch.nonzero host sync in the middle:
//     PHCPufferEnv.step                  puffer_phc/clean_pufferl/env.py:102-140   terminals / truncations / masks, episode returns and
//                                                                                  lengths, the logged means
//       -> HumanoidPHC.reset(indices)    puffer_phc/envs/humanoid_phc.py:90-103
//         -> _reset_envs                 :663-674
//           -> _reset_ref_state_init     :692-727   _sample_ref_state (:843-873: sample_time_interval + get_motion_state with the env's
//                                                   CURRENT global offset) + _set_env_state (:899-929) + per-env scalars
//           -> _reset_env_tensors        :729-777   progress / reset / terminate cleared
//           -> _compute_observations(ids):935-959   self obs + task obs (reference at t+1, global offset now 0) into obs_buf[ids]
//
// Two launches, both captured with the step in one CUDA graph:
//   1. auto_reset_scan_kernel   CTA = 1024 consecutive envs, thread = 4 envs: reads reset / terminate / reward, writes terminals,
//      truncations, masks, updates episode returns / lengths, sums the episode metrics (block partials, fixed order), and leaves an
//      ORDERED list of the block's flagged envs plus the block's count -- so the k-th flagged env in ascending env order is known
//      without a global scan (the tail finds it by a binary search over the <= 4096 block counts).  torch.nonzero is not needed.
//   2. auto_reset_tail_kernel   persistent, one CTA of 8 warps per SM, warp = flagged env (rank k), lane = body: start time from the
//      pre-drawn uniform phase[k] (the k-th flagged env consumes the k-th random number, like the reference's
//      torch.rand(len(env_ids))), motion-state query written straight into root / dof / rigid-body state, per-env scalars cleared,
//      the 934-float observation row recomputed in shared memory and written with coalesced stores (+ the normalised copy), and --
//      when the step accumulates RunningNorm moments -- the CORRECTION of those moments: the reference's statistics are taken over
//      the observations it stores, i.e. the post-reset row for a terminated env and no row at all for a truncated (masked) env
//      (clean_pufferl/env.py:132-133, structs.py:116), so the tail adds (new - old) resp. subtracts old, per column, in fp64, into
//      its CTA's own partial slot, and reports the rows to subtract.
// Everything is deterministic: fixed env -> (CTA, warp) assignment for a given flag vector, fixed summation orders.
#include "phc_body.cuh"

namespace phc {

constexpr int AR_BLOCK = 1024;       // envs per scan CTA
constexpr int AR_THREADS = 256;      // 4 envs per thread
constexpr int AR_TWARPS = 8;         // warps per tail CTA
#ifndef AR_TCTAS
#define AR_TCTAS 2                    // tail CTAs per SM (68 KB of shared memory each; 3 would fit but need 80 registers: spills, 73.6 vs 68.3 us)
#endif
constexpr int AR_MAX_BLOCKS = 4096;  // N <= 4 Mi envs per call
constexpr int AR_ROW = 936;          // floats per row buffer (934 padded to a multiple of 4)
constexpr int AR_COLS = (934 + 31) / 32;   // observation columns per lane
constexpr int AR_NM = 12;            // metric sums the scan kernel forms: PHC_M_REWARD .. PHC_M_EPISODES

// What the tail needs to know about a flagged env, gathered by the (massively parallel) scan kernel so that the tail's per-env chain
// of dependent loads is: this record -> both frame gathers -> math, instead of id -> motion meta -> offset -> frames -> frames.
struct __align__(16) ARec {
    int32_t e_local;     // env = block * AR_BLOCK + e_local
    int32_t truncated;   // flagged but not terminated
    int64_t id;          // _sampled_motion_ids[e]
    int64_t nf, ls;      // _motion_num_frames[id], length_starts[id]
    float mlen, mdt;     // _motion_lengths[id], _motion_dt[id]
    float offx, offy, offz, pad;   // _global_offset[e] before it is cleared
};
static_assert(sizeof(ARec) == 64, "record layout");

struct ARArgs {
    phc_motion_tables t;
    phc_reset_env env;
    phc_reset_book book;
    phc_reset_cfg cfg;
    const float* phase;
    int64_t N;
    int64_t* reset_ids;
    int32_t* reset_count;
    int32_t* block_counts;      // [nb]
    double* block_metrics;      // [nb][PHC_NUM_METRICS]
    ARec* block_recs;           // [nb][AR_BLOCK] records of the block's flagged envs, ascending env order
    int nb;
    double* moment_partials;    // [grid][2][934] or NULL
    double* row_adjust;         // [1] or NULL: -= truncated rows
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(AR_THREADS) auto_reset_scan_kernel(const ARArgs a) {
    __shared__ int s_cnt[AR_THREADS / 32];
    __shared__ double s_m[AR_THREADS / 32][AR_NM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const phc_reset_env& env = a.env;
    const phc_reset_book& bk = a.book;
    const int64_t base = (int64_t)blockIdx.x * AR_BLOCK;
    const int64_t e0 = base + tid * 4;
    double m[AR_NM];
#pragma unroll
    for (int k = 0; k < AR_NM; ++k) m[k] = 0.0;
    unsigned flagged = 0;
    // phase 1: every input of the thread's four envs is requested before anything is stored (the stores below may alias the loads as far
    // as the compiler knows, so a single loop would pay one memory round trip per env: the kernel is pure latency, 64 CTAs at 65536 envs)
    bool ok[4], rs[4], tm[4];
    float rw[4], rt[4];
    int32_t ln[4];
    int64_t mid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t e = e0 + k;
        ok[k] = e < a.N;
        rs[k] = tm[k] = false; rw[k] = rt[k] = 0.0f; ln[k] = 0; mid[k] = 0;
        if (ok[k]) {
            rs[k] = env.reset[e] != 0;
            tm[k] = env.terminated[e] != 0;
            rw[k] = bk.rewards ? bk.rewards[e] : 0.0f;
            if (bk.episode_returns) {
                rt[k] = bk.episode_returns[e];
                ln[k] = bk.episode_lengths ? bk.episode_lengths[e] : 0;
            }
            mid[k] = __ldg(env.motion_ids + e);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t e = e0 + k;
        if (!ok[k]) continue;
        const bool r = rs[k], t = tm[k];
        const bool trunc = r && !t;                                        // env.py:128-129
        if (r) flagged |= 1u << k;
        if (bk.terminals) bk.terminals[e] = t ? 1 : 0;                     // env.py:124-126 (terminate is a subset of reset)
        if (bk.truncations) bk.truncations[e] = trunc ? 1 : 0;
        if (bk.masks) bk.masks[e] = trunc ? 0 : 1;                         // env.py:132-133
        const float rew = rw[k];
        if (bk.episode_returns) {
            float ret = rt[k];
            int32_t len = ln[k];
            if (r) {                                                       // env.py:116-120
                m[PHC_M_EP_RETURN - 1] += (double)ret;
                m[PHC_M_EP_LENGTH - 1] += (double)len;
                ret = 0.0f;
                len = 0;
            }
            ret = ret + rew;                                               // env.py:139-140: reset_buf is already cleared, so every
            len = len + 1;                                                 // env accumulates (the reset ones start their new episode)
            bk.episode_returns[e] = ret;
            if (bk.episode_lengths) bk.episode_lengths[e] = len;
        }
        if (r) m[PHC_M_EPISODES - 1] += 1.0;
        if (trunc) m[PHC_M_TRUNCATIONS - 1] += 1.0;
        if (bk.step_metrics) {                                             // env.py:102-110 (when the step kernel does not do it)
            m[PHC_M_REWARD - 1] += (double)rew;
            if (bk.reward_raw)
                for (int c = 0; c < bk.raw_dim && c < 5; ++c) m[PHC_M_RAW0 - 1 + c] += (double)bk.reward_raw[e * bk.raw_stride + c];
            if (r) m[PHC_M_RESETS - 1] += 1.0;
            if (t) m[PHC_M_TERMINATIONS - 1] += 1.0;
        }
    }
    // ---- ordered compaction of the block's flagged envs -----------------------------------------------------------
    const int cnt = __popc(flagged);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_cnt[warp] = incl;
#pragma unroll
    for (int k = 0; k < AR_NM; ++k) {
        const double s = warp_sum_d(m[k]);
        if (lane == 0) s_m[warp][k] = s;
    }
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < AR_THREADS / 32; ++w) {
        if (w < warp) woff += s_cnt[w];
        total += s_cnt[w];
    }
    int pos = woff + incl - cnt;
    ARec* recs = a.block_recs + (int64_t)blockIdx.x * AR_BLOCK;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((flagged >> k) & 1u) {
            const int64_t e = e0 + k;
            const int64_t id = mid[k];
            ARec rc;
            rc.e_local = tid * 4 + k;
            rc.truncated = !tm[k];
            rc.id = id;
            rc.nf = __ldg(a.t.num_frames + id);
            rc.ls = __ldg(a.t.length_starts + id);
            rc.mlen = __ldg(a.t.motion_len + id);
            rc.mdt = __ldg(a.t.motion_dt + id);
            rc.offx = env.global_offset[e * 3]; rc.offy = env.global_offset[e * 3 + 1]; rc.offz = env.global_offset[e * 3 + 2];
            rc.pad = 0.0f;
            recs[pos++] = rc;
        }
    if (tid == 0) a.block_counts[blockIdx.x] = total;
    if (tid < AR_NM) {
        double s = s_m[0][tid];
#pragma unroll
        for (int w = 1; w < AR_THREADS / 32; ++w) s += s_m[w][tid];
        a.block_metrics[(int64_t)blockIdx.x * PHC_NUM_METRICS + 1 + tid] = s;
    }
}

// get_motion_state for one body (motion_lib.py:596-610), reference operation order (every lerp two products and a sum), split into
// the gathers and the math so that the gathers of BOTH queries of a reset are in flight together.
struct RawPair { V3 p0, p1, v0, v1, w0, w1; Q4 q0, q1; };
__device__ __forceinline__ RawPair load_pair(const phc_motion_tables& T, int64_t f0, int64_t f1, int j) {
    RawPair r;
    r.p0 = ldg3(T.gts + (f0 * NB + j) * 3); r.p1 = ldg3(T.gts + (f1 * NB + j) * 3);
    r.v0 = ldg3(T.gvs + (f0 * NB + j) * 3); r.v1 = ldg3(T.gvs + (f1 * NB + j) * 3);
    r.w0 = ldg3(T.gavs + (f0 * NB + j) * 3); r.w1 = ldg3(T.gavs + (f1 * NB + j) * 3);
    r.q0 = ldg4a(T.grs + (f0 * NB + j) * 4); r.q1 = ldg4a(T.grs + (f1 * NB + j) * 4);
    return r;
}
__device__ __forceinline__ BodyState blend_pair(const RawPair& a, float blend, V3 off, int dev) {
    const float om = 1.0f - blend;
    BodyState r;
    r.p = V3{lerp(a.p0.x, a.p1.x, om, blend) + off.x, lerp(a.p0.y, a.p1.y, om, blend) + off.y, lerp(a.p0.z, a.p1.z, om, blend) + off.z};
    r.q = slerp_rcp(a.q0, a.q1, blend, dev);
    r.v = V3{lerp(a.v0.x, a.v1.x, om, blend), lerp(a.v0.y, a.v1.y, om, blend), lerp(a.v0.z, a.v1.z, om, blend)};
    r.w = V3{lerp(a.w0.x, a.w1.x, om, blend), lerp(a.w0.y, a.w1.y, om, blend), lerp(a.w0.z, a.w1.z, om, blend)};
    return r;
}

__global__ void __launch_bounds__(AR_TWARPS * 32, AR_TCTAS) auto_reset_tail_kernel(const ARArgs a) {
    extern __shared__ double smem_d[];
    const bool mom = a.moment_partials != nullptr;
    float* rows = reinterpret_cast<float*>(smem_d);                              // [AR_TWARPS][AR_ROW] the new observation rows
    float* olds = rows + AR_TWARPS * AR_ROW;                                     // [AR_TWARPS][AR_ROW] the rows they replace (only with moments)
    float* s_mean = olds + (mom ? AR_TWARPS * AR_ROW : 0);                       // [AR_ROW] RunningNorm mean (only with obs_norm)
    float* s_inv = s_mean + AR_ROW;                                              // [AR_ROW] 1 / sqrt(var + eps), one IEEE sqrt + division per column per CTA
    int* prefix = reinterpret_cast<int*>(s_inv + AR_ROW);                        // [nb + 1] exclusive prefix of the block counts
    __shared__ int s_part[AR_TWARPS * 32];
    __shared__ int s_flag[AR_TWARPS];             // this round's env of warp w: 0 none, 1 reset after a termination, 2 truncated
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const phc_motion_tables& T = a.t;
    const phc_reset_env& env = a.env;
    const phc_reset_cfg& cfg = a.cfg;

    // ---- exclusive prefix of the block counts (nb <= 4096): thread = a run of consecutive blocks ----------------------------
    const int per = (a.nb + AR_TWARPS * 32 - 1) / (AR_TWARPS * 32);
    int local = 0;
    for (int i = 0; i < per; ++i) {
        const int b = tid * per + i;
        if (b < a.nb) local += a.block_counts[b];
    }
    s_part[tid] = local;
    if (env.obs_norm)
        for (int c = tid; c < OBS_W; c += AR_TWARPS * 32) {
            s_mean[c] = __ldg(env.rms_mean + c);
            s_inv[c] = 1.0f / sqrtf(__ldg(env.rms_var + c) + cfg.rms_eps);
        }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < AR_TWARPS * 32; ++i) { const int v = s_part[i]; s_part[i] = run; run += v; }
    }
    __syncthreads();
    {
        int run = s_part[tid];
        for (int i = 0; i < per; ++i) {
            const int b = tid * per + i;
            if (b < a.nb) { prefix[b] = run; run += a.block_counts[b]; }
            if (b == a.nb - 1) prefix[a.nb] = run;
        }
    }
    __syncthreads();
    const int K = prefix[a.nb];

    if (blockIdx.x == 0) {      // the call's scalar results: count, metrics (block partials folded in block order), rows to subtract
        if (tid == 0 && a.reset_count) a.reset_count[0] = K;
        if (tid < AR_NM) {
            double s = 0.0;
            for (int b = 0; b < a.nb; ++b) s += a.block_metrics[(int64_t)b * PHC_NUM_METRICS + 1 + tid];
            const int idx = 1 + tid;
            const bool step_metric = idx <= PHC_M_TERMINATIONS;
            if (a.book.metrics && (!step_metric || a.book.step_metrics)) a.book.metrics[idx] += s;
            if (idx == PHC_M_TRUNCATIONS && a.row_adjust) a.row_adjust[0] -= s;
        }
        if (tid == 0 && a.book.metrics && a.book.step_metrics) a.book.metrics[PHC_M_STEPS] += (double)a.N;
    }

    float* row = rows + warp * AR_ROW;
    float* old = olds + warp * AR_ROW;
    // moment correction: thread t of the CTA owns columns t, t + 256, ... and adds the CTA's (up to) eight rows of a round in warp
    // order -- fp64 accumulators in REGISTERS, a fixed order, and no per-warp accumulator arrays in shared memory (those had held
    // the kernel at one CTA = 8 warps per SM; now three CTAs fit)
    // (thread t owns the column PAIRS t and t + 256: 8-byte shared-memory loads)
    constexpr int AR_TCOLS = 2 * ((OBS_W / 2 + AR_TWARPS * 32 - 1) / (AR_TWARPS * 32));
    static_assert(OBS_W % 2 == 0 && AR_ROW % 2 == 0, "column pairs");
    const bool vec2 = (reinterpret_cast<uintptr_t>(env.obs) & 7u) == 0 && (env.obs_stride & 1) == 0 &&
                      (!env.obs_norm || (reinterpret_cast<uintptr_t>(env.obs_norm) & 7u) == 0);
    double as[AR_TCOLS], aq[AR_TCOLS];
#pragma unroll
    for (int k = 0; k < AR_TCOLS; ++k) as[k] = aq[k] = 0.0;
    const int total_warps = gridDim.x * AR_TWARPS;
    const float fps = (float)(1.0 / 30.0);        // curr_fps = 1/30 as a Python double, cast at the op (motion_lib.py:532)
    for (int r0 = blockIdx.x * AR_TWARPS; r0 < K; r0 += total_warps) {
      const int r = r0 + warp;
      if (mom && lane == 0) s_flag[warp] = 0;
      if (r < K) {
        int lo = 0, hi = a.nb;                     // largest block with prefix <= r
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] <= r) lo = mid; else hi = mid;
        }
        const ARec rc = a.block_recs[(int64_t)lo * AR_BLOCK + (r - prefix[lo])];     // one 64-byte record, the same for every lane
        const int64_t e = (int64_t)lo * AR_BLOCK + rc.e_local;
        if (lane == 0 && a.reset_ids) a.reset_ids[r] = e;        // ascending env order = torch.nonzero order
        const bool truncated = rc.truncated != 0;               // flagged and not terminated
        float* orow = env.obs + e * env.obs_stride;
        float* nrow = env.obs_norm ? env.obs_norm + e * env.obs_stride : nullptr;
        // the pre-reset observation row (only needed for the moment correction): all 30 loads per lane are issued here, before the two
        // motion-state queries, so that their latency hides behind the gathers and the math instead of following them
        if (mom) {                                             // 934 floats = 467 8-byte pieces (rows are 8-byte aligned when the
            if ((reinterpret_cast<uintptr_t>(orow) & 7u) == 0) {   // stride is even), asynchronously: no registers, no stall here
                for (int i = lane; i < OBS_W / 2; i += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(old + 2 * i)), "l"(orow + 2 * i) : "memory");
            } else {
                for (int i = lane; i < OBS_W; i += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(old + i)), "l"(orow + i) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (lane == 0) s_flag[warp] = truncated ? 2 : 1;
        }
        const float mlen = rc.mlen, mdt = rc.mdt;
        const int64_t nf = rc.nf, ls = rc.ls;
        // _sample_ref_state (:843-857): StateInit.Random / Hybrid sample a frame-quantised start time, Start and flag_test use 0
        float start = 0.0f;
        if (cfg.state_init == 0 && !cfg.flag_test) {
            const float x = __ldg(a.phase + r) * mlen;                                           // motion_lib.py:527-533
            const float qv = cfg.ref_device == PHC_REF_CUDA ? x * (float)(1.0 / (1.0 / 30.0)) : x / fps;
            start = (float)(int64_t)qv * fps;
        }
        const V3 off_old{rc.offx, rc.offy, rc.offz};             // the query uses the env's CURRENT offset (:859-861) ...
        // both queries of the reset -- the new state at `start`, the reference of the observation at t+1 = (0 + 1) * dt + start + 0
        // (:935-959, offset now zero) -- depend on the record only: their gathers are issued back to back
        int64_t i0, i1, k0, k1;
        float bl, bl1;
        frame_blend(start, mlen, nf, mdt, i0, i1, bl);
        const float t1 = ((float)(int16_t)1 * cfg.dt + start) + 0.0f;
        frame_blend(t1, mlen, nf, mdt, k0, k1, bl1);
        const int j = lane;
        RawPair ra{}, rb{};
        Q4 l0{}, l1{};
        V3 d0{}, d1{};
        if (j < NB) {
            ra = load_pair(T, i0 + ls, i1 + ls, j);
            rb = load_pair(T, k0 + ls, k1 + ls, j);
            if (j >= 1) {
                if (env.dof_pos) { l0 = ldg4a(T.lrs + ((i0 + ls) * NB + j) * 4); l1 = ldg4a(T.lrs + ((i1 + ls) * NB + j) * 4); }
                if (env.dof_vel) { d0 = ldg3(T.dvs + ((i0 + ls) * 23 + (j - 1)) * 3); d1 = ldg3(T.dvs + ((i1 + ls) * 23 + (j - 1)) * 3); }
            }
        }
        BodyState b{};
        if (j < NB) b = blend_pair(ra, bl, off_old, cfg.ref_device);
        // ---- _set_env_state (:899-929) ----------------------------------------------------------------------------------
        if (j < NB) {
            float* o = env.body_state + e * env.env_stride + REC * j;
            st3(o, b.p); st4(o + 3, b.q); st3(o + 7, b.v); st3(o + 10, b.w);
            if (j == 0 && env.root_states) {
                float* rs = env.root_states + e * REC;
                st3(rs, b.p); st4(rs + 3, b.q); st3(rs + 7, b.v); st3(rs + 10, b.w);
            }
            if (j >= 1) {
                if (env.dof_pos) st3(env.dof_pos + e * NDOF + (j - 1) * 3, quat_exp_map_fast(slerp_rcp(l0, l1, bl, cfg.ref_device)));
                if (env.dof_vel) {
                    const float om = 1.0f - bl;
                    st3(env.dof_vel + e * NDOF + (j - 1) * 3, V3{lerp(d0.x, d1.x, om, bl), lerp(d0.y, d1.y, om, bl), lerp(d0.z, d1.z, om, bl)});
                }
            }
        }
        if (lane == 0) {                                         // :721-727, :774-777
            env.global_offset[e * 3] = 0.0f; env.global_offset[e * 3 + 1] = 0.0f; env.global_offset[e * 3 + 2] = 0.0f;
            env.start_time[e] = start;
            env.start_offset[e] = 0.0f;
            env.progress[e] = 0;
            env.reset[e] = 0;
            env.terminated[e] = 0;
        }
        // ---- _compute_observations(env_ids) (:935-959) against the reference at t+1 -------------------------------------------------
        BodyState ref{};
        if (j < NB) ref = blend_pair(rb, bl1, V3{0.0f, 0.0f, 0.0f}, cfg.ref_device);
        const V3 root_p{__shfl_sync(FULL, b.p.x, 0), __shfl_sync(FULL, b.p.y, 0), __shfl_sync(FULL, b.p.z, 0)};
        const Q4 root_q{__shfl_sync(FULL, b.q.x, 0), __shfl_sync(FULL, b.q.y, 0), __shfl_sync(FULL, b.q.z, 0), __shfl_sync(FULL, b.q.w, 0)};
        float hz, hw;
        heading_quat(calc_heading(root_q), hz, hw);
        if (j < NB) {
            if (j == 0) row[0] = root_p.z;                                                                 // common.py:40
            // the same per-body arithmetic as the stand-alone observation kernels (imitation.cu) and the fused step
            const ZRot hrot = zrot_make(hz, hw);
            self_obs_pos_rot_fma(b, root_p, hz, hw, hrot, j, row + 1 + 3 * (j - 1), row + 70 + 6 * j);
            self_obs_vel_ang_fma(b, hrot, row + 214 + 3 * j, row + 286 + 3 * j);
            float* q = row + OBS_SELF;
            task_obs_body_fma(b, ref, root_p, hz, hw, hrot, q + 3 * j, q + 72 + 6 * j, q + 216 + 3 * j, q + 288 + 3 * j, q + 360 + 3 * j,
                              q + 432 + 6 * j);
        }
        // the asynchronous copy of the pre-reset row must have READ the global row before the loop below overwrites it
        if (mom) asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (vec2) {                                              // 8-byte aligned rows: column pairs
#pragma unroll
            for (int k = 0; k < (OBS_W / 2 + 31) / 32; ++k) {
                const int c = 2 * (lane + 32 * k);
                if (c < OBS_W) {
                    const float2 nv = *reinterpret_cast<const float2*>(row + c);
                    *reinterpret_cast<float2*>(orow + c) = nv;
                    if (nrow) {
                        const float2 m = *reinterpret_cast<const float2*>(s_mean + c), iv = *reinterpret_cast<const float2*>(s_inv + c);
                        float y0 = (nv.x - m.x) * iv.x, y1 = (nv.y - m.y) * iv.y;
                        y0 = (y0 != y0) ? y0 : fminf(fmaxf(y0, -cfg.rms_clip), cfg.rms_clip);
                        y1 = (y1 != y1) ? y1 : fminf(fmaxf(y1, -cfg.rms_clip), cfg.rms_clip);
                        *reinterpret_cast<float2*>(nrow + c) = make_float2(y0, y1);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < AR_COLS; ++k) {
                const int c = lane + 32 * k;
                if (c < OBS_W) {
                    const float nv = row[c];
                    orow[c] = nv;
                    if (nrow) {
                        float y = (nv - s_mean[c]) * s_inv[c];
                        y = (y != y) ? y : fminf(fmaxf(y, -cfg.rms_clip), cfg.rms_clip);
                        nrow[c] = y;
                    }
                }
            }
        }
        __syncwarp();
      }
      if (mom) {
        __syncthreads();                                        // every warp's new row and (landed) old row are in shared memory
        for (int w = 0; w < AR_TWARPS; ++w) {
            const int f = s_flag[w];
            if (f == 0) continue;
            const float* nr = rows + w * AR_ROW;
            const float* orr = olds + w * AR_ROW;
#pragma unroll
            for (int k = 0; k < AR_TCOLS / 2; ++k) {
                const int c = 2 * (tid + k * AR_TWARPS * 32);
                if (c < OBS_W) {
                    const float2 o2 = *reinterpret_cast<const float2*>(orr + c), n2 = *reinterpret_cast<const float2*>(nr + c);
                    const double od0 = (double)o2.x, od1 = (double)o2.y, nd0 = (double)n2.x, nd1 = (double)n2.y;
                    if (f == 2) { as[2 * k] -= od0; aq[2 * k] -= od0 * od0; as[2 * k + 1] -= od1; aq[2 * k + 1] -= od1 * od1; }
                    else {
                        as[2 * k] += nd0 - od0; aq[2 * k] += nd0 * nd0 - od0 * od0;
                        as[2 * k + 1] += nd1 - od1; aq[2 * k + 1] += nd1 * nd1 - od1 * od1;
                    }
                }
            }
        }
        __syncthreads();                                        // the row buffers are free for the next round
      }
    }
    if (mom) {
        double* slot = a.moment_partials + (int64_t)blockIdx.x * 2 * OBS_W;
#pragma unroll
        for (int k = 0; k < AR_TCOLS / 2; ++k) {
            const int c = 2 * (tid + k * AR_TWARPS * 32);
            if (c < OBS_W) {
                slot[c] += as[2 * k]; slot[c + 1] += as[2 * k + 1];
                slot[OBS_W + c] += aq[2 * k]; slot[OBS_W + c + 1] += aq[2 * k + 1];
            }
        }
    }
}

static int ar_blocks(int64_t N) { return (int)((N + AR_BLOCK - 1) / AR_BLOCK); }

// moments[1 + i] += sum_p partial[p][i], metrics[k] += sum_p metric_partial[p][k], moments[0] += rows (+ *row_adjust); optionally
// every partial that was read is zeroed by the thread that read it (reduce + clear in one launch).  Same fixed order as
// rms_reduce_kernel: a block owns 32 columns, its 8 warps each sum a stride-8 subset of the partial rows.
__global__ void __launch_bounds__(256) stats_reduce_kernel(double* __restrict__ partial, int P, int C, double rows, double* row_adjust,
                                                           double* __restrict__ mpartial, int MP, double* __restrict__ stats, int zero) {
    __shared__ double sh[8][33];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int ncol = 2 * C;
    const int nblk_cols = (ncol + 31) / 32;
    if ((int)blockIdx.x < nblk_cols) {
        const int i = blockIdx.x * 32 + lane;
        double s = 0.0;
        if (i < ncol && partial) {              // four independent loads in flight per thread, then their clears (fixed order)
            int p = g;
            for (; p + 24 < P; p += 32) {
                double* q = partial + (int64_t)p * ncol + i;
                const double x0 = q[0], x1 = q[(int64_t)8 * ncol], x2 = q[(int64_t)16 * ncol], x3 = q[(int64_t)24 * ncol];
                if (zero) { q[0] = 0.0; q[(int64_t)8 * ncol] = 0.0; q[(int64_t)16 * ncol] = 0.0; q[(int64_t)24 * ncol] = 0.0; }
                s += (x0 + x1) + (x2 + x3);
            }
            for (; p < P; p += 8) {
                double* q = partial + (int64_t)p * ncol + i;
                s += *q;
                if (zero) *q = 0.0;
            }
        }
        sh[g][lane] = s;
        __syncthreads();
        if (g == 0 && i < ncol) {
            double t = sh[0][lane];
#pragma unroll
            for (int k = 1; k < 8; ++k) t += sh[k][lane];
            stats[1 + i] += t;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double n = rows;
            if (row_adjust) { n += row_adjust[0]; if (zero) row_adjust[0] = 0.0; }
            stats[0] += n;
        }
    } else if (mpartial) {          // the last block: the metric slots, thread = metric, partial rows in order
        const int k = threadIdx.x;
        if (k < PHC_NUM_METRICS) {
            double s = 0.0;
            for (int p = 0; p < MP; ++p) {
                double* q = mpartial + (int64_t)p * PHC_NUM_METRICS + k;
                s += *q;
                if (zero) *q = 0.0;
            }
            stats[1 + ncol + k] += s;
        }
    }
}

}  // namespace phc

using namespace phc;

extern "C" int phc_auto_reset_num_partials(void) { return AR_TCTAS * sm_count(); }

extern "C" int64_t phc_auto_reset_scratch_bytes(int64_t N) {
    if (N <= 0) return 0;
    const int64_t nb = ar_blocks(N);
    return nb * (int64_t)(sizeof(int32_t) + PHC_NUM_METRICS * sizeof(double) + AR_BLOCK * sizeof(ARec)) + 128;
}

extern "C" int phc_auto_reset(const phc_motion_tables* t, const phc_reset_env* env, const phc_reset_book* book,
                              const phc_reset_cfg* cfg, const float* phase, int64_t N, int64_t* reset_ids, int32_t* reset_count,
                              void* scratch, double* moment_partials, double* row_adjust, phc_stream_t stream) {
    const char* fn = "phc_auto_reset";
    PHC_REQUIRE(t && env && book && cfg, PHC_EINVAL, "%s: NULL argument struct", fn);
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    if (N == 0) return PHC_OK;
    PHC_REQUIRE(ar_blocks(N) <= AR_MAX_BLOCKS, PHC_EUNSUPPORTED, "%s: N=%lld > %d envs per call", fn, (long long)N, AR_MAX_BLOCKS * AR_BLOCK);
    PHC_REQUIRE(scratch && aligned16(scratch), PHC_EINVAL, "%s: scratch is NULL or not 16-byte aligned (phc_auto_reset_scratch_bytes)", fn);
    PHC_REQUIRE(env->body_state && env->progress && env->start_time && env->start_offset && env->global_offset && env->motion_ids &&
                    env->reset && env->terminated && env->obs, PHC_EINVAL, "%s: a required env tensor is NULL", fn);
    PHC_REQUIRE(env->env_stride >= SIM_F && env->obs_stride >= OBS_W, PHC_ESHAPE, "%s: env_stride / obs_stride too small", fn);
    PHC_REQUIRE(!env->obs_norm || (env->rms_mean && env->rms_var), PHC_EINVAL, "%s: obs_norm needs rms_mean and rms_var", fn);
    PHC_REQUIRE(cfg->state_init == 0 || cfg->state_init == 1, PHC_EINVAL, "%s: state_init must be 0 (random) or 1 (start)", fn);
    PHC_REQUIRE(cfg->state_init == 1 || cfg->flag_test || phase, PHC_EINVAL, "%s: phase is NULL", fn);
    PHC_REQUIRE(cfg->ref_device == PHC_REF_DEVICE_CPU || cfg->ref_device == PHC_REF_DEVICE_CUDA, PHC_EINVAL, "%s: ref_device must be 0 or 1", fn);
    PHC_REQUIRE(t->gts && t->grs && t->gvs && t->gavs && t->motion_len && t->motion_dt && t->num_frames && t->length_starts, PHC_EINVAL,
                "%s: motion tables missing", fn);
    PHC_REQUIRE(aligned16(t->grs) && (!env->dof_pos || (t->lrs && aligned16(t->lrs))) && (!env->dof_vel || t->dvs), PHC_EINVAL,
                "%s: grs / lrs / dvs tables missing or misaligned", fn);
    PHC_REQUIRE(!book->episode_lengths || book->episode_returns, PHC_EINVAL, "%s: episode_lengths needs episode_returns", fn);
    PHC_REQUIRE(!book->reward_raw || (book->raw_dim >= 1 && book->raw_stride >= book->raw_dim), PHC_ESHAPE, "%s: reward_raw shape", fn);
    const int nb = ar_blocks(N);
    char* sp = static_cast<char*>(scratch);
    ARArgs a{*t, *env, *book, *cfg, phase, N, reset_ids, reset_count, nullptr, nullptr, nullptr, nb, moment_partials, row_adjust};
    a.block_recs = reinterpret_cast<ARec*>(sp);
    a.block_metrics = reinterpret_cast<double*>(sp + (size_t)nb * AR_BLOCK * sizeof(ARec));
    a.block_counts = reinterpret_cast<int32_t*>(a.block_metrics + (size_t)nb * PHC_NUM_METRICS);
    cudaStream_t s = (cudaStream_t)stream;
    auto_reset_scan_kernel<<<nb, AR_THREADS, 0, s>>>(a);
    int rc = check_launch(fn);
    if (rc) return rc;
    const size_t smem = (size_t)((moment_partials ? 2 : 1) * AR_TWARPS + 2) * AR_ROW * sizeof(float) + (size_t)(nb + 1) * sizeof(int);
    cudaError_t e = cudaFuncSetAttribute(auto_reset_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute(%zu B smem): %s", fn, smem, cudaGetErrorString(e));
    auto_reset_tail_kernel<<<phc_auto_reset_num_partials(), AR_TWARPS * 32, smem, s>>>(a);
    return check_launch(fn);
}

extern "C" int phc_stats_reduce(double* moment_partials, int num_partials, int C, int64_t rows, double* row_adjust,
                                double* metric_partials, int num_metric_partials, double* stats, int zero_partials, phc_stream_t stream) {
    const char* fn = "phc_stats_reduce";
    PHC_REQUIRE(stats, PHC_EINVAL, "%s: stats is NULL", fn);
    PHC_REQUIRE(C >= 1 && num_partials >= 0 && num_metric_partials >= 0 && rows >= 0, PHC_EINVAL, "%s: bad size", fn);
    PHC_REQUIRE(moment_partials || num_partials == 0, PHC_EINVAL, "%s: moment_partials is NULL", fn);
    PHC_REQUIRE(metric_partials || num_metric_partials == 0, PHC_EINVAL, "%s: metric_partials is NULL", fn);
    const int blocks = (2 * C + 31) / 32 + (metric_partials ? 1 : 0);
    stats_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(moment_partials, num_partials, C, (double)rows, row_adjust, metric_partials,
                                                                 num_metric_partials, stats, zero_partials);
    return check_launch(fn);
}
