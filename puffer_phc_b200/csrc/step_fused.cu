// step_fused.cu -- the whole post-physics half of HumanoidPHC.step in one pass over HBM
// (reference puffer_phc/envs/humanoid_phc.py:136-149): two get_motion_state queries (t for reward /
// reset, t+1 for the task observation; motion_lib.py:549-626), compute_imitation_reward (+ power term,
// humanoid_phc.py:1295-1303), compute_humanoid_im_reset, compute_humanoid_observations_smpl_max,
// compute_imitation_observations_v6, and optionally RunningNorm.forward plus the column moments that
// RunningNorm.update needs (policies/running_norm.py:15-34).
//
// Design (persistent, warp-specialised, one CTA per SM):
//   * compute groups: each iteration the CTA owns S consecutive envs.  A group of THREE warps works on FOUR envs
//     (4 x 24 bodies = 96 lanes: every lane carries one body, no idle lanes) and there are two groups per four envs:
//         role A: reference at t   -> reward terms, termination test, power term and their per-env reductions / exponentials
//         role B: reference at t+1 -> imitation (task) observation and the whole self observation (the only heading-frame math)
//     Every (env, role) owns a private shared-memory buffer (PhysX record | frame 0 | frame 1 | dof force/vel).  The
//     NEXT env's record and frames are fetched with cp.async right after the current ones have been read into
//     registers, so the gathers fly behind ~700 instructions of math and never occupy registers (without this the
//     kernel is latency-bound, not bandwidth-bound).  Per-env reductions (24 bodies spread over two warps) go
//     through shared memory in a fixed order (deterministic).
//   * 4 writer warps: the S x 934-float observation tile is double-buffered in shared memory.  When a tile is
//     full (mbarrier) the writers send it to HBM as ONE contiguous 16-byte aligned block with a TMA bulk
//     store (cp.async.bulk), write the RunningNorm-normalised copy and accumulate the fp64 column moments
//     from the same tile (column-pair-owning threads keep mean, 1/sqrt(var+eps) and the accumulators in registers),
//     then release the tile (mbarrier).  Compute warps never wait for stores.
//   * 1 planner warp (runs up to three iterations ahead): lane = (env slot, role) reads the per-env
//     scalars (coalesced across envs), does the id -> motion-meta lookups and the frame-index / blend arithmetic
//     (bit-exact op order) and leaves a 48-byte plan per (env, role) in shared memory, so compute warps never
//     execute (32x redundantly) or wait on that dependent load chain.
//   * Register budget: each SM sub-partition holds 16384 registers = 5 warps x 96 or 6 warps x 80.  Default S = 12: 18 compute + 5 writer
//     + 1 planner warps = 24 warps at 80 registers (no spills in the production instantiation).  The flag-critical chain keeps the
//     reference's fp32 op order.
#include <type_traits>

#include "phc_body.cuh"

namespace phc {

#ifndef ST_SLOTS
#define ST_SLOTS 12                                  // envs per CTA iteration (multiple of 4: one compute group = 4 envs).  Round 2: 12 envs
#endif                                               // (18 compute + 5 writer + 1 planner warps at 80 registers, 0 spills, 214 KB smem)
                                                     // beat 8 envs (17 warps at 96 registers) by 5.4 % once the pair tables had
                                                     // trimmed the compute warps: 0.1574 vs 0.1663 ms at 65536 envs
constexpr int ST_ENVS = ST_SLOTS;
constexpr int ST_GROUPS = ST_ENVS / 4;               // compute groups per role
constexpr int ST_CWARPS = 2 * 3 * ST_GROUPS;         // compute warps: 3 warps per group, 2 roles
constexpr int ST_NBUF = 2 * ST_ENVS;                 // staging buffers / plans per iteration: (role, slot)
static_assert(ST_ENVS % 4 == 0 && ST_NBUF <= 32, "a compute group handles 4 envs; the planner has one lane per buffer");
#ifndef ST_CHINT
#define ST_CHINT 0
#endif
#ifndef ST_WHINT
#define ST_WHINT 400
#endif
#ifndef ST_BHINT
#define ST_BHINT ST_CHINT                            // suspend-time hint of role B's waits (it idles behind role A through the plan ring)
#endif
#ifndef ST_TILES
#define ST_TILES 2                                   // observation tiles in flight between compute and writer warps
#endif
#ifndef ST_WUNROLL
#define ST_WUNROLL 1
#endif
#ifndef ST_WRITERS
#define ST_WRITERS 5                                 // 12 slots: 4 writers 0.1594 ms, 5 writers 0.1574 ms; 6 do not fit (25 warps x 80 registers:
#endif                                               // a seventh warp on one SM sub-partition exceeds its 16384 registers)
#ifndef ST_TMA_LOADS
#define ST_TMA_LOADS 0                               // 1: sim record + packed frames arrive by TMA bulk loads (one lane issues three
                                                     // cp.async.bulk with mbarrier byte counting instead of 12 cp.async per lane);
                                                     // bit-identical, measured 0.1728-0.1742 vs 0.1716-0.1719 ms: the staging
                                                     // instructions are not what the critical role waits for
#endif
#ifndef ST_A_WAITS_TILE
#define ST_A_WAITS_TILE 0
#endif
#ifndef ST_STRESS_DELAY
#define ST_STRESS_DELAY 0                            // test builds only: 1 = role B sleeps in every iteration, 2 = role A does, 3 = the writers do
#endif                                               // (tests/test_fused_stress.py: outputs must stay bit-identical under any role skew)
#ifndef ST_PLAN_EARLY
#define ST_PLAN_EARLY 0                              // 1: the planner computes plan p before it waits for the ring slot -- no gain
#endif                                               // (0.1745 vs 0.1729 ms): the plans are not what the critical role waits for
#ifndef ST_PREFETCH
#define ST_PREFETCH 0                                // 1: planner lanes prefetch the coming frame / sim records into L2 -- measured
                                                     // SLOWER (0.179 -> 0.220 ms): the prefetches queue ahead of the demand gathers
#endif
constexpr int ST_WWARPS = ST_WRITERS;                // writer warps
// register budget per SM sub-partition (16384 registers): ceil(warps / 4) x 32 x ST_MAXREG must fit
#ifndef ST_MAXREG
#define ST_NWARPS (6 * (ST_SLOTS / 4) + ST_WRITERS + 1)
#define ST_MAXREG ((ST_NWARPS <= 20) ? 96 : ((ST_NWARPS <= 24) ? 80 : ((ST_NWARPS <= 28) ? 72 : 64)))
#endif
constexpr int ST_WTHREADS = ST_WWARPS * 32;
constexpr int ST_THREADS = (ST_CWARPS + ST_WWARPS + 1) * 32;     // + the planner warp
constexpr int ST_PLANS = 4;                          // plan ring: plans are produced three iterations ahead (a ring of 8 with its own
                                                     // "free" barriers, seven ahead, measured slower: 0.186 vs 0.176 ms)
// Ring of per-iteration metric records handed from role A's leader lanes to the writers.  Role A is never more than 5 iterations
// ahead of the writers (its plan it+1 needs full[it-2], role B's tile it-2 needed empty[] of tile it-4), so a ring of 8 needs no wait.
constexpr int ST_META = 8;
constexpr int ST_WPAIRS = (OBS_W / 2 + ST_WTHREADS - 1) / ST_WTHREADS;  // column pairs owned by a writer thread
constexpr int ST_DOF_F = 144;                        // dof_force (69, padded to 72) | dof_vel (69, padded to 72)
constexpr int ST_AUX_F = 2 * NB;                     // slerp pair quantities of frame 0's pair: (h, +-1/sin_half) per body
constexpr int ST_AUX_OFF = 3 * FRAME_F + ST_DOF_F;
constexpr int ST_WBUF_F = ST_AUX_OFF + ST_AUX_F;     // per (env, role): sim record | frame 0 | frame 1 | dof force/vel | pair aux
// diagnosis builds (profiles/tools/ab_variants.py): compile single round-2 additions out to price them
#ifndef ST_DIAG_NOMETRICS
#define ST_DIAG_NOMETRICS 0
#endif
#ifndef ST_DIAG_NOEVAL
#define ST_DIAG_NOEVAL 0
#endif
#ifndef ST_DIAG_DEV0
#define ST_DIAG_DEV0 0
#endif
#ifndef ST_DIAG_NOTAIL
#define ST_DIAG_NOTAIL 0                              // timing experiment only: role A skips its env-level tail (outputs are wrong)
#endif
#ifndef ST_DIAG_NOFULLWAIT
#define ST_DIAG_NOFULLWAIT 0
#endif
#if ST_DIAG_DEV0
#define ST_DEV(cfg) 0
#else
#define ST_DEV(cfg) (cfg).ref_device
#endif
#ifndef ST_SPLIT_DONE
#define ST_SPLIT_DONE 0                              // 1: the writers wait for role B only (role A signals on its own barrier ring);
                                                     // bit-identical, measured 0.1683 vs 0.1677 ms: role A's lateness is not what
                                                     // the writers wait for
#endif
#ifndef ST_BALANCE
#define ST_BALANCE 1                                 // 0: always fill all ST_SLOTS slots of a block (A/B builds)
#endif
#ifndef ST_USE_AUX
#define ST_USE_AUX 1                                 // 0: ignore the motion library's pair tables (A/B builds)
#endif
constexpr unsigned SPIN_LIMIT = 1u << 22;            // a stuck mbarrier traps (after a few seconds) instead of hanging the GPU

#ifndef ST_WARP_PERM
#define ST_WARP_PERM ((6 * (ST_SLOTS / 4) + ST_WRITERS + 1) == 24)      // the table below is for the 24-warp configuration
#endif
#if ST_WARP_PERM
// physical warp -> logical warp; logical 0..8 role A, 9..17 role B, 18..22 writers, 23 planner
__constant__ unsigned char c_warp_perm[24] = {0, 1, 2, 7, 3, 4, 5, 8, 6, 9, 10, 11, 18, 12, 13, 14, 19, 15, 16, 17, 23, 20, 21, 22};
#endif
// -DST_PROFILE=1: lane 0 of every compute warp accumulates the clock cycles it spends in each wait of its loop
// (0 cp.async landing + group barrier, 1 plan barrier, 2 tile-release barrier, 3 group barriers: buffers free + the two of the
// reduction, 4 whole loop) into
// g_prof[SM][warp][5]; read with phc_debug_profile().  Tuning builds only.
#ifndef ST_PROFILE
#define ST_PROFILE 0
#endif
#if ST_PROFILE
__device__ unsigned long long g_prof[160][32][5];
#define PROF_DECL unsigned long long prof_t[5] = {0, 0, 0, 0, 0}; long long prof_c = 0, prof_l = clock64();
#define PROF_BEGIN prof_c = clock64();
#define PROF_END(k) prof_t[k] += (unsigned long long)(clock64() - prof_c);
#else
#define PROF_DECL
#define PROF_BEGIN
#define PROF_END(k)
#endif

struct StepArgs {
    phc_motion_tables t;
    phc_step_in in;
    phc_step_cfg cfg;
    phc_step_out out;
    int sim_vec;          // body_state rows are 16-byte aligned -> 16-byte cp.async staging
    int obs_vec;          // obs tiles are contiguous and 16-byte aligned -> TMA bulk tile stores
    int64_t num_blocks;   // ceil(N / epb)
    int use_aux;          // the motion library's pair tables exist and were built for cfg.ref_device
    int epb;              // envs per block (even, <= ST_ENVS): small batches are spread evenly over the CTAs' iterations instead of
                          // filling 12 slots on some SMs and none on others; slots >= epb idle
};

// ---- async-copy / barrier primitives ------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Compute warps poll (they are the critical path and rarely wait); writer / planner warps pass a suspend-time hint
// so that the hardware parks them instead of letting their polls steal issue slots from the math warps.
template <int HINT_NS>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done = 0, spins = 0;
    while (!done) {
        if (HINT_NS > 0)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "n"(HINT_NS) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();
    }
}
// TMA bulk LOAD global -> shared, completion counted in bytes on an mbarrier (one instruction per 16-byte aligned block instead
// of 78 per-lane 16-byte cp.async for a 1248-byte record).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// L2 prefetch of a 16-byte aligned block (no destination, no register, no shared memory): issued by the planner lanes up to
// three iterations ahead, so that the compute warps' cp.async gathers of the frame records hit L2 instead of waiting on DRAM.
__device__ __forceinline__ void prefetch_l2(const void* gsrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// named barriers 2.. : one per compute group (3 warps = 96 threads)
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void writers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ST_WTHREADS) : "memory"); }
// torch.clamp propagates NaN: min.NaN / max.NaN do too (fminf / fmaxf would drop it)
__device__ __forceinline__ float clamp_nan(float y, float lim) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(y), "f"(-lim));
    asm("min.NaN.f32 %0, %0, %1;" : "+f"(r) : "f"(lim));
    return r;
}

// ---- what a compute warp needs to know about its env for one role (written by the planner warp) -----------------
struct EnvPlan {
    int64_t f0, f1;      // global frame rows (frame index + length_starts)
    float blend, t, mlen;
    float offx, offy, offz;
    int prog;
    int valid;           // bit 0: env < N; bit 1: blend == 0 fast path (frame 1 is neither fetched nor read)
};

// Planner lane: per-env scalars, motion meta and the role's frame-blend (reference op order, bit-exact).
__device__ __forceinline__ EnvPlan make_plan(const StepArgs& a, int64_t e, int role) {
    const phc_motion_tables& T = a.t;
    const phc_step_in& in = a.in;
    EnvPlan p;
    p.valid = e < in.N;
    if (!p.valid) { p.f0 = p.f1 = 0; p.blend = p.t = p.mlen = p.offx = p.offy = p.offz = 0.0f; p.prog = 0; return p; }
    const int64_t id = __ldg(in.motion_ids + e);
    const int16_t prog = __ldg(in.progress + e);
    const float st = __ldg(in.start_time + e), so = __ldg(in.start_offset + e);
    const float mdt = __ldg(T.motion_dt + id);
    const int64_t nf = __ldg(T.num_frames + id), ls = __ldg(T.length_starts + id);
    p.mlen = __ldg(T.motion_len + id);
    p.offx = __ldg(in.global_offset + e * 3);
    p.offy = __ldg(in.global_offset + e * 3 + 1);
    p.offz = __ldg(in.global_offset + e * 3 + 2);
    p.prog = prog;
    // humanoid_phc.py:1233-1235 (t) and :1060-1064 (t+1: progress_buf + 1 stays int16)
    const int16_t step = role == 0 ? prog : (int16_t)(prog + 1);
    p.t = ((float)step * a.cfg.dt + st) + so;
    int64_t i0, i1;
    frame_blend(p.t, p.mlen, nf, mdt, i0, i1, p.blend);
    p.f0 = i0 + ls;
    p.f1 = i1 + ls;
    // the query sits exactly on a table frame: frame 1 would only contribute 0 * x, so it is not fetched (pair tables of the motion
    // library).  The one exception, a body whose pair takes slerp's un-normalised midpoint fall-back, is rare (frozen poses) and
    // fetches its second rotation on demand in the compute warp -- deciding it here would cost the planner a third dependent load.
    if (a.use_aux && p.f1 != p.f0 && p.blend == 0.0f) p.valid |= 2;
    return p;
}

// Eval variant of the termination test (common.py:342-346): mean of the per-body distances of the body subset, in torch's own
// summation order for the chosen device.  Kept out of line: it runs once per env in eval mode only, and its local arrays and loops
// must not sit in the instruction stream the three warp roles share.
__device__ __noinline__ float eval_mean_distance(const float* dist8 /* stride 8 floats per body */, unsigned mask, int dev) {
    float dsub[NB];
    int n = 0;
    for (int b2 = 0; b2 < NB; ++b2)
        if ((mask >> b2) & 1u) dsub[n++] = dist8[b2 * 8];
    return mean_ordered(dsub, n, dev);
}

// Stage one reference frame (pos 72 | rot 96 | vel 72 | ang 72 floats) into shared memory: 78 x 16-byte chunks,
// issued by the 24 lanes (j = body index) that work on this env.
template <bool PACKED>
__device__ __forceinline__ void stage_frame(const phc_motion_tables& T, int64_t f, float* dst, int j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = j + NB * k;
        if (i < FRAME_F / 4) {
            const float* src;
            if (PACKED) src = T.packed + f * FRAME_F + 4 * i;
            else if (i < 18) src = T.gts + f * 72 + 4 * i;
            else if (i < 42) src = T.grs + f * 96 + 4 * (i - 18);
            else if (i < 60) src = T.gvs + f * 72 + 4 * (i - 42);
            else src = T.gavs + f * 72 + 4 * (i - 60);
            cp_async16(dst + 4 * i, src);
        }
    }
}

// Async fetch of everything the 24 lanes of (env e, role) read: PhysX record, two frames, (role A) dof force / vel.
// BULK (packed frame records and 16-byte aligned sim rows): lane j == 0 announces the byte count on the buffer's mbarrier and
// issues one TMA bulk load per record; the other lanes issue nothing for them.  Otherwise every lane stages its share with
// cp.async (completion through cp.async groups + the group barrier).
template <bool PACKED>
__device__ __forceinline__ void issue_env(const StepArgs& a, const EnvPlan& p, int64_t e, int role, float* wbuf, int j, uint64_t* lbar) {
    const phc_step_in& in = a.in;
    const float* rec = in.body_state + e * in.env_stride;
    const bool two = p.f1 != p.f0 && !(p.valid & 2);
    if (a.use_aux && j < ST_AUX_F / 4) cp_async16(wbuf + ST_AUX_OFF + 4 * j, a.t.pair_aux + p.f0 * ST_AUX_F + 4 * j);
    if (PACKED && ST_TMA_LOADS) {
        if (j == 0) {
            const unsigned fbytes = FRAME_F * 4;
            mbar_expect_tx(lbar, (a.sim_vec ? (unsigned)SIM_F * 4 : 0u) + (two ? 2 * fbytes : fbytes));
            if (a.sim_vec) bulk_load(wbuf, rec, SIM_F * 4, lbar);
            bulk_load(wbuf + FRAME_F, a.t.packed + p.f0 * FRAME_F, fbytes, lbar);
            if (two) bulk_load(wbuf + 2 * FRAME_F, a.t.packed + p.f1 * FRAME_F, fbytes, lbar);
        }
        if (!a.sim_vec) {
#pragma unroll
            for (int k = 0; k < REC; ++k) cp_async4(wbuf + j + NB * k, rec + j + NB * k);
        }
    } else {
        if (a.sim_vec) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = j + NB * k;
                if (i < SIM_F / 4) cp_async16(wbuf + 4 * i, rec + 4 * i);
            }
        } else {
#pragma unroll
            for (int k = 0; k < REC; ++k) cp_async4(wbuf + j + NB * k, rec + j + NB * k);
        }
        stage_frame<PACKED>(a.t, p.f0, wbuf + FRAME_F, j);
        if (two) stage_frame<PACKED>(a.t, p.f1, wbuf + 2 * FRAME_F, j);
    }
    if (role == 0 && in.dof_force) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int c = j + NB * k;
            if (c < NDOF) {
                cp_async4(wbuf + 3 * FRAME_F + c, in.dof_force + e * NDOF + c);
                cp_async4(wbuf + 3 * FRAME_F + 72 + c, in.dof_vel + e * NDOF + c);
            }
        }
    }
}

// The rotation block of a staged frame starts 288 bytes into a 16-byte aligned record: one 16-byte shared-memory load per lane
// (four scalar loads at a stride of four floats would put lanes j, j + 8 and j + 16 on the same banks).
__device__ __forceinline__ BodyState read_frame(const float* f, int j) {
    const float4 q = *reinterpret_cast<const float4*>(f + 72 + 4 * j);
    return BodyState{ld3(f + 3 * j), Q4{q.x, q.y, q.z, q.w}, ld3(f + 168 + 3 * j), ld3(f + 240 + 3 * j)};
}

__device__ __forceinline__ void store_ref(float* dst, int j, const BodyState& r) {
    st3(dst + 3 * j, r.p);
    st4(dst + 72 + 4 * j, r.q);
    st3(dst + 168 + 3 * j, r.v);
    st3(dst + 240 + 3 * j, r.w);
}

// DEBUG_REF instantiates the optional ref_state_t / ref_state_t1 outputs (otherwise their predicated-off stores would still
// take load/store issue slots in every lane).
template <bool PACKED, bool DEBUG_REF, bool AUX>
__global__ void __maxnreg__(ST_MAXREG) step_fused_kernel(const StepArgs a) {   // one persistent CTA per SM
    static_assert(!AUX || (PACKED && !ST_TMA_LOADS), "the pair tables ride with the packed records and cp.async staging");
    extern __shared__ float4 smem4[];
    float* tiles = reinterpret_cast<float*>(smem4);                          // [ST_TILES][S][934]
    float* wbufs = tiles + ST_TILES * ST_ENVS * OBS_W;                       // [2][S][record | frame 0 | frame 1 | dof]
    float* red = wbufs + ST_NBUF * ST_WBUF_F;                                // [S][24][8]  per-body partials (role A)
    float* red2 = red + ST_ENVS * NB * 8;                                    // [S][6][4]   second-stage partial sums
    float* meta = red2 + ST_ENVS * 24;                                       // [ST_META][S][12] per-env metric values, role A -> writers
    EnvPlan* plans = reinterpret_cast<EnvPlan*>(meta + ST_META * ST_ENVS * 12);   // [ST_PLANS][2S]
    uint64_t* bars = reinterpret_cast<uint64_t*>(plans + ST_PLANS * ST_NBUF);
    uint64_t* full = bars;                    // [ST_TILES] tile b written by all compute warps
    uint64_t* empty = bars + ST_TILES;        // [ST_TILES] tile b drained by the writers
    uint64_t* pfull = bars + 2 * ST_TILES;    // [ST_PLANS] plan set d written by the planning writer warp
    uint64_t* lfull = pfull + ST_PLANS;       // [2S] staging buffer (role, slot): bytes of the TMA bulk loads have landed
    uint64_t* adone = lfull + ST_NBUF;        // [ST_META] role A has finished iteration it (ring slot it % ST_META): its metric record is
                                              // written and its plan can be recycled.  Role A writes no tile rows, so the writers do
                                              // NOT wait for it before shipping a tile (full[] counts role B only).

    const int tid = threadIdx.x, lane = tid & 31;
    // logical warp (role assignment) of this physical warp.  Warp p issues from SM sub-partition p % 4; with the identity mapping
    // every role-A group (the critical role) has a warp on sub-partition 0, which then carries 3 role-A + 2 role-B + 1 writer warps,
    // the heaviest mix of the four.  ST_WARP_PERM deals the roles so that the sub-partition with three role-A warps gets the lightest
    // companions (two writers and the planner) and the others 2 A + 3 B + 1 writer each.
#if ST_WARP_PERM
    const int warp = c_warp_perm[tid >> 5];
#else
    const int warp = tid >> 5;
#endif
    const phc_step_in& in = a.in;
    const phc_step_cfg& cfg = a.cfg;
    const phc_step_out& out = a.out;

    if (tid == 0) {
        for (int i = 0; i < ST_TILES; ++i) { mbar_init(&full[i], ST_SPLIT_DONE ? ST_CWARPS / 2 : ST_CWARPS); mbar_init(&empty[i], 1); }
        for (int i = 0; i < ST_META; ++i) mbar_init(&adone[i], ST_CWARPS / 2);
        for (int i = 0; i < ST_PLANS; ++i) mbar_init(&pfull[i], 1);
        for (int i = 0; i < ST_NBUF; ++i) mbar_init(&lfull[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp < ST_CWARPS) {
        // ====================================== compute warps ======================================
        // group = 3 warps = 96 lanes = 4 envs x 24 bodies; lane t of the group works on body t % 24 of env slot 4g + t / 24
        const int role = warp / (3 * ST_GROUPS), gw = warp % (3 * ST_GROUPS);
        const int g = gw / 3, t = (gw % 3) * 32 + lane;
        const int el = t / NB, j = t % NB;
        const int slot = 4 * g + el, buf = role * ST_ENVS + slot;
        const int bar_id = 2 + role * ST_GROUPS + g;
        float* wbuf = wbufs + buf * ST_WBUF_F;
        EnvPlan cur{};
        if ((int64_t)blockIdx.x < a.num_blocks) {
            mbar_wait<ST_CHINT>(&pfull[0], 0);
            cur = plans[buf];
            if (cur.valid & 1) issue_env<PACKED>(a, cur, (int64_t)blockIdx.x * a.epb + slot, role, wbuf, j, &lfull[buf]);
        }
        cp_async_commit();
        int it = 0;
        const float my_term_dist = __ldg(in.term_dist + j);       // this lane's body never changes
        PROF_DECL
        for (int64_t blk = blockIdx.x; blk < a.num_blocks; blk += gridDim.x, ++it) {
            const int64_t e = blk * a.epb + slot;
            const bool valid = slot < a.epb && e < in.N;
            const int b = it % ST_TILES, use = it / ST_TILES;                  // use-th time tile buffer b is filled
            float* my_tile = tiles + (b * ST_ENVS + slot) * OBS_W;

            // ---- operands of this body: shared memory -> registers, blend the two frames ----------------
            PROF_BEGIN
            cp_async_wait_all();
            if (PACKED && ST_TMA_LOADS && (cur.valid & 1)) mbar_wait<ST_CHINT>(&lfull[buf], it & 1);   // the bulk loads of this env have landed
            group_sync(bar_id);                                                // the cp.async copies of all 96 lanes have landed
            PROF_END(0)
            BodyState body{}, ref{};
            V3 root_p{};
            Q4 root_q{0.0f, 0.0f, 0.0f, 1.0f};
            float power = 0.0f;
            if (valid) {
                if (role == 1) {                 // only the observation role works in the heading frame of the simulated root
                    root_p = ld3(wbuf);
                    root_q = ld4(wbuf + 3);
                }
                const float* sj = wbuf + REC * j;
                body = BodyState{ld3(sj), ld4(sj + 3), ld3(sj + 7), ld3(sj + 10)};
                const BodyState F0 = read_frame(wbuf + FRAME_F, j);
                const V3 off{cur.offx, cur.offy, cur.offz};
                if (AUX) {
                    const float2 ax = *reinterpret_cast<const float2*>(wbuf + ST_AUX_OFF + 2 * j);
                    const SlerpPair sp{ax.x, ax.y};
                    if (cur.valid & 2) {
                        ref = blend_frames_t0(F0, off, sp);
                        if (sp.h == -2.0f) {       // midpoint fall-back 0.5 q0 + 0.5 (+-q1): the only case that needs frame 1's rotation
                            const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.t.packed + cur.f1 * FRAME_F + 72 + 4 * j));
                            ref.q = slerp_pair(F0.q, Q4{q1.x, q1.y, q1.z, q1.w}, 0.0f, sp);
                        }
                    } else {
                        const BodyState F1 = (cur.f1 == cur.f0) ? F0 : read_frame(wbuf + 2 * FRAME_F, j);
                        ref = blend_frames_pair(F0, F1, cur.blend, off, sp);
                    }
                } else {
                    const BodyState F1 = (cur.f1 == cur.f0) ? F0 : read_frame(wbuf + 2 * FRAME_F, j);
                    ref = blend_frames(F0, F1, cur.blend, off, ST_DEV(cfg));
                }
                if (role == 0 && in.dof_force && j < 23) {                         // humanoid_phc.py:1295-1303
                    const float* df = wbuf + 3 * FRAME_F + 3 * j;
                    power = (fabsf(df[0] * df[72]) + fabsf(df[1] * df[73])) + fabsf(df[2] * df[74]);
                }
            }
            PROF_BEGIN
            group_sync(bar_id);                                                // everyone has read: the buffers are free again
            PROF_END(3)
            // ---- fetch the next env's record and frames behind the math -----------------------------------
            EnvPlan nxt{};
            {
                const int64_t nblk = blk + gridDim.x;
                if (nblk < a.num_blocks) {
                    const int d = (it + 1) % ST_PLANS;
                    PROF_BEGIN
                    if (role == 1) mbar_wait<ST_BHINT>(&pfull[d], ((it + 1) / ST_PLANS) & 1);
                    else mbar_wait<ST_CHINT>(&pfull[d], ((it + 1) / ST_PLANS) & 1);
                    PROF_END(1)
                    nxt = plans[d * ST_NBUF + buf];
                    if (nxt.valid & 1) issue_env<PACKED>(a, nxt, nblk * a.epb + slot, role, wbuf, j, &lfull[buf]);
                }
                cp_async_commit();
            }
            PROF_BEGIN
            // tile buffer b released by the writers -- only role B writes tile rows, so only role B waits for the writers.
            // Role A still must not ARRIVE on full[b] for this use before the previous phase of full[b] has completed: with
            // ST_TILES = 2 its six arrivals of iteration it could otherwise pair up with its own six of iteration it - 2 while role B
            // is still writing that older tile (the plan wait below holds role A back everywhere except in a CTA's last
            // iteration, which has no next plan).  The wait is one try_wait that succeeds at once whenever role A is the slower role.
            if ((ST_A_WAITS_TILE || role == 1) && use >= 1) mbar_wait<ST_BHINT>(&empty[b], (use - 1) & 1);
            else if (ST_SPLIT_DONE) { if (it >= ST_META) mbar_wait<ST_CHINT>(&adone[it & (ST_META - 1)], ((it / ST_META) - 1) & 1); }
            else if (!ST_DIAG_NOFULLWAIT && use >= 1) mbar_wait<ST_CHINT>(&full[b], (use - 1) & 1);
            PROF_END(2)

#if ST_STRESS_DELAY == 1 || ST_STRESS_DELAY == 2
            if (role == 2 - ST_STRESS_DELAY) __nanosleep(3000 + 997 * ((it * 7 + warp) % 5));
#endif
            if (role == 0) {
                // ============ role A: reward, reset, power (reference at t) ============
                float* rj = red + (slot * NB + j) * 8;
                if (valid) {
                    float sp, sr, sv, sa, dist = 0.0f;
                    reward_terms_body_fma(body, ref, sp, sr, sv, sa);
                    if ((cfg.reset_body_mask >> j) & 1u) {
                        dist = norm3(body.p - ref.p, ST_DEV(cfg));
                        if (!cfg.use_mean) dist = (dist > my_term_dist) ? 1.0f : 0.0f;                // common.py:347-350 (any)
                    }
                    *reinterpret_cast<float4*>(rj) = make_float4(sp, sr, sv, sa);
                    *reinterpret_cast<float2*>(rj + 4) = make_float2(dist, power);
                    if (DEBUG_REF && out.ref_state_t) store_ref(out.ref_state_t + e * FRAME_F, j, ref);
                }
                PROF_BEGIN
                group_sync(bar_id);
                PROF_END(3)
                {   // second stage: lane (value v, part p) of the env adds bodies 6p .. 6p+5, in order
                    const int v = j >> 2, p4 = j & 3;
                    const float* src = red + (slot * NB + 6 * p4) * 8 + v;
                    float s = src[0];
#pragma unroll
                    for (int k = 1; k < 6; ++k) s = s + src[8 * k];
                    red2[(slot * 6 + v) * 4 + p4] = s;
                }
#if ST_DIAG_NOTAIL
                if (false) {
#else
                {
#endif
                PROF_BEGIN
                group_sync(bar_id);
                PROF_END(3)
                // env-level tail.  The six lanes j < 6 of an env sit in ONE warp (the env's 24 lanes start at lane 0, 24, 16 or 8 of a
                // warp): lane v totals value v in the fixed order and, for the four reward terms, evaluates its exponential kernel
                // (common.py:300-316) -- four expf side by side instead of one after the other in a single leader lane; the leader
                // (j == 0) collects them with shuffles.
                float val = 0.0f;
                if (j < 6) {
                    const float4 q4 = *reinterpret_cast<const float4*>(red2 + (slot * 6 + j) * 4);
                    val = ((q4.x + q4.y) + q4.z) + q4.w;
                    if (j < 4) {
                        const float inv = (j == 1) ? (1.0f / (float)NB) : (1.0f / (3.0f * (float)NB));
                        val = expf(-cfg.k[j] * (val * inv));
                    }
                }
                const int base = lane - j;                                                // lane of this env's j == 0 (same warp for j < 6)
                const float r0 = __shfl_sync(FULL, val, base & 31), r1 = __shfl_sync(FULL, val, (base + 1) & 31);
                const float r2 = __shfl_sync(FULL, val, (base + 2) & 31), r3 = __shfl_sync(FULL, val, (base + 3) & 31);
                const float t4 = __shfl_sync(FULL, val, (base + 4) & 31), t5 = __shfl_sync(FULL, val, (base + 5) & 31);
                if (valid && j == 0) {
                    bool fallen = false;
                    if (cfg.enable_early_termination) {
                        if (cfg.use_mean) {                                               // common.py:342-346
                            const int first = __ffs(cfg.reset_body_mask) - 1;
#if ST_DIAG_NOEVAL
                            const float mean = t4 / (float)__popc(cfg.reset_body_mask & 0xffffffu);
#else
                            const float mean = eval_mean_distance(red + slot * NB * 8 + 4, cfg.reset_body_mask, cfg.ref_device);
#endif
                            fallen = mean > __ldg(in.term_dist + first);
                        } else {
                            fallen = t4 > 0.0f;
                        }
                        fallen = fallen && (cur.prog > 1);                                // common.py:354
                    }
                    float rew = ((cfg.w[0] * r0 + cfg.w[1] * r1) + cfg.w[2] * r2) + cfg.w[3] * r3;    // common.py:318-320
                    float* rr = out.reward_raw + e * out.raw_stride;
                    rr[0] = r0; rr[1] = r1; rr[2] = r2; rr[3] = r3;
                    float pr = 0.0f;
                    if (in.dof_force) {
                        pr = -cfg.power_coef * t5;
                        if (cur.prog <= 3) pr = 0.0f;
                        rew = rew + pr;
                        rr[4] = pr;
                    }
                    const bool rst = (cur.t >= cur.mlen) || fallen;                       // humanoid_phc.py:1315, common.py:362
                    out.reward[e] = rew;
                    out.terminated[e] = fallen ? 1 : 0;
                    out.reset[e] = rst ? 1 : 0;
                    if (!ST_DIAG_NOMETRICS && out.metric_partials) {       // episode metrics (clean_pufferl/env.py:102-110): plain stores into
                        // this iteration's record; the WRITER warps do the summing (role A's serial tail is the kernel's critical path)
                        float4* mm = reinterpret_cast<float4*>(meta + ((it & (ST_META - 1)) * ST_ENVS + slot) * 12);
                        mm[0] = make_float4(1.0f, rew, r0, r1);
                        mm[1] = make_float4(r2, r3, pr, rst ? 1.0f : 0.0f);
                        mm[2] = make_float4(fallen ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
                    }
                }
                }
            } else if (valid) {
                // ============ role B: every observation block (task obs against the reference at t+1, self obs) ===========
                float hz, hw;
                heading_quat_direct(root_q, hz, hw);                           // upright start: no base-rot removal; role A needs no heading
                const ZRot hrot = zrot_make(hz, hw);
                if (DEBUG_REF && out.ref_state_t1) store_ref(out.ref_state_t1 + e * FRAME_F, j, ref);
                if (j == 0) my_tile[0] = root_p.z;                                        // common.py:40
                self_obs_pos_rot_fma(body, root_p, hz, hw, hrot, j, my_tile + 1 + 3 * (j - 1), my_tile + 70 + 6 * j);
                self_obs_vel_ang_fma(body, hrot, my_tile + 214 + 3 * j, my_tile + 286 + 3 * j);
                float* q = my_tile + OBS_SELF;
                task_obs_body_fma(body, ref, root_p, hz, hw, hrot, q + 3 * j, q + 72 + 6 * j, q + 216 + 3 * j, q + 288 + 3 * j,
                                  q + 360 + 3 * j, q + 432 + 6 * j);
            }
            if (ST_SPLIT_DONE && role == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&adone[it & (ST_META - 1)]);
            } else {
                fence_proxy_async();        // tile rows were written by ordinary stores; the writers read them through TMA
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[b]);
            }
            cur = nxt;
        }
#if ST_PROFILE
        prof_t[4] = (unsigned long long)(clock64() - prof_l);
        if (lane == 0) for (int k = 0; k < 5; ++k) g_prof[blockIdx.x][warp][k] = prof_t[k];
#endif
        cp_async_wait_all();
    } else if (warp < ST_CWARPS + ST_WWARPS) {
        // ====================================== writer warps =======================================
        const int wtid = (warp - ST_CWARPS) * 32 + lane;
        const bool do_norm = out.obs_norm != nullptr;
        const bool do_mom = out.moment_partials != nullptr;
        const bool do_met = !ST_DIAG_NOMETRICS && out.metric_partials != nullptr && wtid < 12;
        double macc = 0.0;                  // writer thread k < 12 sums metric k over the envs of every tile, in tile / row order
        // each writer thread owns ST_WPAIRS pairs of adjacent observation columns (float2 granularity: rows are 8-byte aligned)
        float2 c_mean[ST_WPAIRS], c_inv[ST_WPAIRS];
        double msum[ST_WPAIRS][2], msq[ST_WPAIRS][2];
#pragma unroll
        for (int u = 0; u < ST_WPAIRS; ++u) {
            const int c = 2 * (wtid + u * ST_WTHREADS);
            msum[u][0] = msum[u][1] = msq[u][0] = msq[u][1] = 0.0;
            c_mean[u] = make_float2(0.0f, 0.0f);
            c_inv[u] = make_float2(1.0f, 1.0f);
            if (do_norm && c < OBS_W) {      // running_norm.py:17, one IEEE reciprocal per column
                c_mean[u] = make_float2(__ldg(in.rms_mean + c), __ldg(in.rms_mean + c + 1));
                c_inv[u] = make_float2(1.0f / sqrtf(__ldg(in.rms_var + c) + cfg.rms_eps), 1.0f / sqrtf(__ldg(in.rms_var + c + 1) + cfg.rms_eps));
            }
        }
        // VEC = contiguous, aligned output rows (the normal case): one float2 store per pair, no pitched-row code in the loop
        // (predicated-off stores would still occupy load/store issue slots).
        // (row pointers are advanced by the caller: no 64-bit index arithmetic per pair in the loop)
        auto column_pass = [&](const float* trow, float* orow, float* nrow, auto vec) {
            constexpr bool VEC = decltype(vec)::value;
#pragma unroll
            for (int u = 0; u < ST_WPAIRS; ++u) {
                const int c = 2 * (wtid + u * ST_WTHREADS);
                if (c < OBS_W) {
                    const float2 x = *reinterpret_cast<const float2*>(trow + c);
                    if (!VEC) { orow[c] = x.x; orow[c + 1] = x.y; }
                    if (do_norm) {   // (x - mean) / sqrt(var + eps) as a multiplication by the column's reciprocal (<= 1.5 ulp)
                        const float y0 = clamp_nan((x.x - c_mean[u].x) * c_inv[u].x, cfg.rms_clip);
                        const float y1 = clamp_nan((x.y - c_mean[u].y) * c_inv[u].y, cfg.rms_clip);
                        if (VEC) *reinterpret_cast<float2*>(nrow + c) = make_float2(y0, y1);
                        else { nrow[c] = y0; nrow[c + 1] = y1; }
                    }
                    if (do_mom) {    // xd*xd is exact in fp64, so fma(xd, xd, q) equals q + xd*xd
                        const double x0 = (double)x.x, x1 = (double)x.y;
                        msum[u][0] += x0; msq[u][0] = fma(x0, x0, msq[u][0]);
                        msum[u][1] += x1; msq[u][1] = fma(x1, x1, msq[u][1]);
                    }
                }
            }
        };
        int it = 0;
        for (int64_t blk = blockIdx.x; blk < a.num_blocks; blk += gridDim.x, ++it) {
            const int b = it % ST_TILES;
            const float* tile = tiles + b * ST_ENVS * OBS_W;
            const int64_t e0 = blk * a.epb;
            const int rows = (int)((in.N - e0 < a.epb) ? (in.N - e0) : a.epb);
            mbar_wait<ST_WHINT>(&full[b], (it / ST_TILES) & 1);
#if !ST_SPLIT_DONE
            if (do_met) {
                const float* mrec = meta + (it & (ST_META - 1)) * ST_ENVS * 12 + wtid;
                for (int r = 0; r < rows; ++r) macc += (double)mrec[r * 12];
            }
#endif
#if ST_STRESS_DELAY == 3
            __nanosleep(4000 + 1000 * (it % 3));
#endif
            // ---- raw observations: one TMA bulk store of the whole tile (manual copy for odd tails / pitched rows) ----
            const bool bulk = a.obs_vec && ((rows & 1) == 0);     // rows * 3736 B is a multiple of 16 for even rows
            if (bulk) {
                if (wtid == 0) bulk_store(out.obs + e0 * OBS_W, tile, (unsigned)(rows * OBS_W * sizeof(float)));
            } else if (a.obs_vec) {
                for (int i = wtid; i < rows * OBS_W; i += ST_WTHREADS) out.obs[e0 * OBS_W + i] = tile[i];
            }
            // ---- normalised copy + fp64 column moments from the same tile ----------------------------------------
            if (!a.obs_vec || do_norm || do_mom) {
                // rolled on purpose: three warp roles share the instruction cache, and the pairs give the ILP
                const float* trow = tile;
                if (a.obs_vec) {
                    float* nrow = do_norm ? out.obs_norm + e0 * OBS_W : nullptr;
#pragma unroll 1
                    for (int r = 0; r < rows; ++r, trow += OBS_W, nrow += OBS_W) column_pass(trow, nullptr, nrow, std::true_type{});
                } else {
                    float* orow = out.obs + e0 * out.obs_stride;
                    float* nrow = do_norm ? out.obs_norm + e0 * out.obs_stride : nullptr;
#pragma unroll 1
                    for (int r = 0; r < rows; ++r, trow += OBS_W, orow += out.obs_stride, nrow += out.obs_stride)
                        column_pass(trow, orow, nrow, std::false_type{});
                }
            }
            if (bulk && wtid == 0) bulk_wait_read();     // the TMA engine has finished reading the tile from shared memory
            writers_sync();
            if (wtid == 0) mbar_arrive(&empty[b]);
#if ST_SPLIT_DONE
            if (do_met) {                                // role A's record of this iteration (it is rarely still missing by now)
                mbar_wait<ST_WHINT>(&adone[it & (ST_META - 1)], (it / ST_META) & 1);
                const float* mrec = meta + (it & (ST_META - 1)) * ST_ENVS * 12 + wtid;
                for (int r = 0; r < rows; ++r) macc += (double)mrec[r * 12];
            }
#endif
        }
        if (wtid == 0) bulk_wait_all();
        if (do_met) {                       // this CTA owns its slot: overwrite, or read-modify-write in accumulate mode (deterministic)
            double* mp = out.metric_partials + (int64_t)blockIdx.x * PHC_NUM_METRICS + wtid;
            *mp = (out.accumulate_partials ? *mp : 0.0) + macc;
        }
        if (do_mom) {
            double* p = out.moment_partials + (int64_t)blockIdx.x * 2 * OBS_W;
#pragma unroll
            for (int u = 0; u < ST_WPAIRS; ++u) {
                const int c = 2 * (wtid + u * ST_WTHREADS);
                if (c < OBS_W) {
                    if (out.accumulate_partials) {       // this CTA owns its slot: read-modify-write, deterministic
                        msum[u][0] += p[c]; msum[u][1] += p[c + 1]; msq[u][0] += p[OBS_W + c]; msq[u][1] += p[OBS_W + c + 1];
                    }
                    p[c] = msum[u][0]; p[c + 1] = msum[u][1]; p[OBS_W + c] = msq[u][0]; p[OBS_W + c + 1] = msq[u][1];
                }
            }
        }
    } else {
        // ====================================== planner warp =======================================
        // Plan p may overwrite ring slot p % 4 (last used by plan p - 4) once full[p - 3] has completed: every compute warp
        // that finished iteration p - 3 has consumed plan p - 2 already.  The compute warps cannot complete iteration p - 1
        // without plan p, so the tile barrier polled here never runs a full phase ahead of this warp.
        int p = 0;
        for (int64_t blk = blockIdx.x; blk < a.num_blocks; blk += gridDim.x, ++p) {
#if ST_PLAN_EARLY
            // the plan is computed (its two dependent load latencies paid) BEFORE the wait for the ring slot
            const int64_t e = blk * a.epb + lane % ST_ENVS;
            EnvPlan pl{};
            if (lane < ST_NBUF && lane % ST_ENVS < a.epb) pl = make_plan(a, e, lane / ST_ENVS);
#endif
            if (p >= ST_PLANS - 1) {
                const int q = p - (ST_PLANS - 1);
                mbar_wait<ST_WHINT>(&full[q % ST_TILES], (q / ST_TILES) & 1);
                if (ST_SPLIT_DONE) mbar_wait<ST_WHINT>(&adone[q & (ST_META - 1)], (q / ST_META) & 1);
            }
            const int d = p % ST_PLANS;
            if (lane < ST_NBUF) {
#if !ST_PLAN_EARLY
                const int64_t e = blk * a.epb + lane % ST_ENVS;
                EnvPlan pl{};
                if (lane % ST_ENVS < a.epb) pl = make_plan(a, e, lane / ST_ENVS);
#endif
                plans[d * ST_NBUF + lane] = pl;
#if ST_PREFETCH
                if (pl.valid & 1) {
                    if (PACKED) {
                        prefetch_l2(a.t.packed + pl.f0 * FRAME_F, FRAME_F * 4);
                        if (pl.f1 != pl.f0) prefetch_l2(a.t.packed + pl.f1 * FRAME_F, FRAME_F * 4);
                    }
                    if (lane < ST_ENVS && a.sim_vec) prefetch_l2(in.body_state + e * in.env_stride, SIM_F * 4);
                }
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[d]);
        }
    }
}

constexpr size_t ST_SMEM = (size_t)(ST_TILES * ST_ENVS * OBS_W + ST_NBUF * ST_WBUF_F + ST_ENVS * NB * 8 + ST_ENVS * 24 + ST_META * ST_ENVS * 12) * sizeof(float) +
                           ST_PLANS * ST_NBUF * sizeof(EnvPlan) +
                           (2 * ST_TILES + ST_PLANS + ST_NBUF + ST_META) * sizeof(uint64_t);
static_assert(sizeof(EnvPlan) == 48, "plan record layout");
static_assert(ST_SMEM <= 227 * 1024, "shared memory budget");

}  // namespace phc

using namespace phc;

extern "C" int phc_step_num_partials(void) { return sm_count(); }

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
    PHC_REQUIRE(cfg->ref_device == PHC_REF_DEVICE_CPU || cfg->ref_device == PHC_REF_DEVICE_CUDA, PHC_EINVAL, "%s: ref_device must be 0 (torch CPU) or 1 (torch CUDA)", fn);
    PHC_REQUIRE(t->motion_len && t->motion_dt && t->num_frames && t->length_starts, PHC_EINVAL, "%s: per-motion tables missing", fn);
    const bool packed = t->packed != nullptr;
    if (packed) {
        PHC_REQUIRE(aligned16(t->packed), PHC_EALIGN, "%s: packed table must be 16-byte aligned", fn);
    } else {
        PHC_REQUIRE(t->gts && t->grs && t->gvs && t->gavs, PHC_EINVAL, "%s: gts/grs/gvs/gavs tables missing", fn);
        PHC_REQUIRE(aligned16(t->gts) && aligned16(t->grs) && aligned16(t->gvs) && aligned16(t->gavs), PHC_EALIGN,
                    "%s: gts/grs/gvs/gavs tables must be 16-byte aligned", fn);
    }
    // the grid is fixed (one persistent CTA per SM) so that the number of moment partial slots does not depend on N
    const int grid = phc_step_num_partials();
    // envs per block: the CTAs' iteration count is that of full 12-env blocks; the envs are then spread evenly over grid x iterations
    // block slots (4096 envs: 3 iterations of 10 on every SM instead of 3 x 12 on 46 SMs and 2 x 12 on the rest)
    int epb = ST_ENVS;
    if (ST_BALANCE && in->N > 0) {
        const int64_t blocks_full = (in->N + ST_ENVS - 1) / ST_ENVS;
        const int64_t iters = (blocks_full + grid - 1) / grid;
        int64_t per = (in->N + grid * iters - 1) / (grid * iters);
        per = (per + 1) & ~(int64_t)1;
        epb = (int)(per < 2 ? 2 : per > ST_ENVS ? ST_ENVS : per);
    }
    StepArgs a{*t, *in, *cfg, *out, 0, 0, (in->N + epb - 1) / epb, 0, epb};
    a.sim_vec = aligned16(in->body_state) && (in->env_stride % 4 == 0);
    a.obs_vec = out->obs_stride == OBS_W && aligned16(out->obs) && (!out->obs_norm || aligned8(out->obs_norm));
    if (in->N == 0 && !out->moment_partials) return PHC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const bool dbg = out->ref_state_t || out->ref_state_t1;
    a.use_aux = ST_USE_AUX && packed && !dbg && !ST_TMA_LOADS && t->pair_aux && t->pair_device == cfg->ref_device &&
                aligned16(t->pair_aux);
    using kernel_t = void (*)(const StepArgs);
    kernel_t k = a.use_aux ? (kernel_t)step_fused_kernel<true, false, ST_USE_AUX && !ST_TMA_LOADS>
                 : packed ? (dbg ? (kernel_t)step_fused_kernel<true, true, false> : (kernel_t)step_fused_kernel<true, false, false>)
                          : (dbg ? (kernel_t)step_fused_kernel<false, true, false> : (kernel_t)step_fused_kernel<false, false, false>);
    // opt in to > 48 KB of dynamic shared memory (a per-device, per-function attribute; the call is cheap, so no cache)
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute(%zu B smem): %s", fn, ST_SMEM, cudaGetErrorString(e));
    k<<<grid, ST_THREADS, ST_SMEM, s>>>(a);
    return check_launch(fn);
}

#if ST_PROFILE
extern "C" int phc_debug_profile(unsigned long long* out_h /* [160*32*5] host */) {
    return (int)cudaMemcpyFromSymbol(out_h, phc::g_prof, sizeof(phc::g_prof));
}
#endif
