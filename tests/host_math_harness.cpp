// TEST-ONLY.  Compiles the product's host/device math headers (puffer_phc_b200/csrc/phc_math.cuh,
// phc_body.cuh) for the HOST with g++ -ffp-contract=off and replays the fused step kernel's per-env flow
// sequentially (lane loop + butterfly-sum emulation), so the kernel math can be checked against the golden
// vectors on a CPU-only box.  The memory/indexing side of the CUDA kernels is covered by the -m gpu tests.
#include <string.h>

#include "../puffer_phc_b200/csrc/phc_body.cuh"

using namespace phc;

static float butterfly_sum(const float* lanes) {
    float v[32], w[32];
    memcpy(v, lanes, sizeof(v));
    for (int o = 16; o > 0; o >>= 1) {
        for (int i = 0; i < 32; ++i) w[i] = v[i] + v[i ^ o];
        memcpy(v, w, sizeof(v));
    }
    return v[0];
}

static BodyState frame(const phc_motion_tables* T, int64_t f, int j) {
    const int64_t r = f * NB + j;
    BodyState s;
    s.p = V3{T->gts[r * 3], T->gts[r * 3 + 1], T->gts[r * 3 + 2]};
    s.q = Q4{T->grs[r * 4], T->grs[r * 4 + 1], T->grs[r * 4 + 2], T->grs[r * 4 + 3]};
    s.v = V3{T->gvs[r * 3], T->gvs[r * 3 + 1], T->gvs[r * 3 + 2]};
    s.w = V3{T->gavs[r * 3], T->gavs[r * 3 + 1], T->gavs[r * 3 + 2]};
    return s;
}

extern "C" int harness_step(const phc_motion_tables* T, const phc_step_in* in, const phc_step_cfg* cfg, const phc_step_out* out) {
    for (int64_t e = 0; e < in->N; ++e) {
        const float* sim = in->body_state + e * in->env_stride;
        const int64_t id = in->motion_ids[e];
        const int16_t prog = in->progress[e];
        const float st = in->start_time[e], so = in->start_offset[e];
        const float mlen = T->motion_len[id], mdt = T->motion_dt[id];
        const int64_t nf = T->num_frames[id], ls = T->length_starts[id];
        const float t0 = ((float)prog * cfg->dt + st) + so;
        const float t1 = ((float)(int16_t)(prog + 1) * cfg->dt + st) + so;
        int64_t a0, a1, b0, b1;
        float bla, blb;
        frame_blend(t0, mlen, nf, mdt, a0, a1, bla);
        frame_blend(t1, mlen, nf, mdt, b0, b1, blb);
        const V3 off{in->global_offset[e * 3], in->global_offset[e * 3 + 1], in->global_offset[e * 3 + 2]};
        const Q4 root_q{sim[3], sim[4], sim[5], sim[6]};
        const V3 root_p{sim[0], sim[1], sim[2]};
        float hz, hw;
        heading_quat_direct(root_q, hz, hw);
        const ZRot hrot = zrot_make(hz, hw);
        float sp[32] = {0}, sr[32] = {0}, sv[32] = {0}, sa[32] = {0}, dist[32] = {0};
        bool over = false;
        float* tile = out->obs + e * out->obs_stride;
        for (int j = 0; j < NB; ++j) {
            const float* sj = sim + REC * j;
            const BodyState body{V3{sj[0], sj[1], sj[2]}, Q4{sj[3], sj[4], sj[5], sj[6]}, V3{sj[7], sj[8], sj[9]}, V3{sj[10], sj[11], sj[12]}};
            const BodyState r0 = blend_frames(frame(T, a0 + ls, j), frame(T, a1 + ls, j), bla, off);
            reward_terms_body_fma(body, r0, sp[j], sr[j], sv[j], sa[j]);
            if ((cfg->reset_body_mask >> j) & 1u) {
                dist[j] = norm3(body.p - r0.p);
                over = over || (dist[j] > in->term_dist[j]);
            }
            const BodyState r1 = blend_frames(frame(T, b0 + ls, j), frame(T, b1 + ls, j), blb, off);
            if (j == 0) tile[0] = root_p.z;
            self_obs_pos_rot_fma(body, root_p, hz, hw, hrot, j, tile + 1 + 3 * (j - 1), tile + 70 + 6 * j);
            self_obs_vel_ang_fma(body, hrot, tile + 214 + 3 * j, tile + 286 + 3 * j);
            float* q = tile + OBS_SELF;
            task_obs_body_fma(body, r1, root_p, hz, hw, hrot, q + 3 * j, q + 72 + 6 * j, q + 216 + 3 * j, q + 288 + 3 * j, q + 360 + 3 * j, q + 432 + 6 * j);
        }
        bool fallen = false;
        if (cfg->enable_early_termination) {
            if (cfg->use_mean) {
                const int first = __builtin_ffs((int)cfg->reset_body_mask) - 1;
                fallen = (butterfly_sum(dist) / (float)__builtin_popcount(cfg->reset_body_mask & 0xffffffu)) > in->term_dist[first];
            } else {
                fallen = over;
            }
            fallen = fallen && (prog > 1);
        }
        float raw[4];
        float rew = reward_from_sq_sums(butterfly_sum(sp), butterfly_sum(sr), butterfly_sum(sv), butterfly_sum(sa), (float)NB, cfg->k, cfg->w, raw);
        float* rr = out->reward_raw + e * out->raw_stride;
        memcpy(rr, raw, sizeof(raw));
        if (in->dof_force) {
            float pl[32] = {0};
            for (int lane = 0; lane < 32; ++lane)
                for (int k = 0; k < 3; ++k) {
                    const int c = lane + 32 * k;
                    if (c < NDOF) pl[lane] = pl[lane] + fabsf(in->dof_force[e * NDOF + c] * in->dof_vel[e * NDOF + c]);
                }
            float pr = -cfg->power_coef * butterfly_sum(pl);
            if (prog <= 3) pr = 0.0f;
            rew = rew + pr;
            rr[4] = pr;
        }
        out->reward[e] = rew;
        out->terminated[e] = fallen ? 1 : 0;
        out->reset[e] = (t0 >= mlen) ? 1 : (fallen ? 1 : 0);
    }
    return 0;
}

// slerp / exp-map path of phc_motion_state for one query batch (dof_pos only needs lrs).
extern "C" int harness_dof_pos(const phc_motion_tables* T, const int64_t* ids, const float* times, int64_t B, float* dof_pos,
                               float* rb_rot, int64_t* idx0, int64_t* idx1, float* blend) {
    for (int64_t q = 0; q < B; ++q) {
        const int64_t id = ids[q];
        int64_t i0, i1;
        float bl;
        frame_blend(times[q], T->motion_len[id], T->num_frames[id], T->motion_dt[id], i0, i1, bl);
        idx0[q] = i0; idx1[q] = i1; blend[q] = bl;
        const int64_t f0 = i0 + T->length_starts[id], f1 = i1 + T->length_starts[id];
        for (int j = 0; j < NB; ++j) {
            const float* a = T->grs + (f0 * NB + j) * 4;
            const float* b = T->grs + (f1 * NB + j) * 4;
            Q4 r = slerp_rcp(Q4{a[0], a[1], a[2], a[3]}, Q4{b[0], b[1], b[2], b[3]}, bl);
            float* o = rb_rot + (q * NB + j) * 4;
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
            if (j >= 1) {
                a = T->lrs + (f0 * NB + j) * 4;
                b = T->lrs + (f1 * NB + j) * 4;
                V3 e = quat_exp_map_fast(slerp_rcp(Q4{a[0], a[1], a[2], a[3]}, Q4{b[0], b[1], b[2], b[3]}, bl));
                put3(dof_pos + q * NDOF + (j - 1) * 3, e);
            }
        }
    }
    return 0;
}
