/*
 * phc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, scalar, CPU restatement of the per-step hot path of howird/puffer-phc,
 * used only as the parity checker for the CUDA kernels (tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg).  Nothing under puffer_phc_b200/ may import, link or
 * call this file.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py executes the reference's own
 * Python (torch CPU, fp32) functions in the build container and commits their inputs
 * and outputs as fixtures; tests/test_oracle_golden.py checks every function below
 * against those fixtures (bit-exact for frame indices / reset flags, 1e-5 rel for the
 * floating-point outputs).
 *
 * All arithmetic is fp32 with every operation individually rounded (torch eager
 * semantics: one kernel per op, no FMA contraction between ops).  Build with
 * -ffp-contract=off (see oracle/Makefile).  Quaternions are xyzw.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define NB 24          /* SMPL bodies                      (puffer_phc/body_sets.py:11-36) */
#define NDOF_J 23      /* actuated joints (all but root)   (puffer_phc/body_sets.py:39)    */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } q4;

/* ------------------------------------------------------------------------------------ */
/* puffer_phc/torch_utils.py                                                              */
/* ------------------------------------------------------------------------------------ */

/* torch_utils.py:55-75 quat_mul -- the 8-multiply form, expression order kept verbatim. */
static q4 o_quat_mul(q4 a, q4 b)
{
    float x1 = a.x, y1 = a.y, z1 = a.z, w1 = a.w;
    float x2 = b.x, y2 = b.y, z2 = b.z, w2 = b.w;
    float ww = (z1 + x1) * (x2 + y2);
    float yy = (w1 - y1) * (w2 + z2);
    float zz = (w1 + y1) * (w2 - z2);
    float xx = (ww + yy) + zz;
    float qq = 0.5f * (xx + (z1 - x1) * (x2 - y2));
    q4 r;
    r.w = (qq - ww) + (z1 - y1) * (y2 - z2);
    r.x = (qq - xx) + (x1 + w1) * (x2 + w2);
    r.y = (qq - yy) + (w1 - x1) * (y2 + z2);
    r.z = (qq - zz) + (z1 + y1) * (w2 - x2);
    return r;
}

/* torch_utils.py:79-82 quat_conjugate */
static q4 o_quat_conj(q4 a) { q4 r = { -a.x, -a.y, -a.z, a.w }; return r; }

/* torch_utils.py:274-281 my_quat_rotate: a + b + c with
 *   a = v * (2 w^2 - 1);  b = cross(q_vec, v) * w * 2;  c = q_vec * dot(q_vec, v) * 2 */
static v3 o_quat_rotate(q4 q, v3 v)
{
    float s = 2.0f * (q.w * q.w) - 1.0f;
    v3 a = { v.x * s, v.y * s, v.z * s };
    v3 cr = { q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x };
    v3 b = { (cr.x * q.w) * 2.0f, (cr.y * q.w) * 2.0f, (cr.z * q.w) * 2.0f };
    float d = (q.x * v.x + q.y * v.y) + q.z * v.z;
    v3 c = { (q.x * d) * 2.0f, (q.y * d) * 2.0f, (q.z * d) * 2.0f };
    v3 r = { (a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z };
    return r;
}

/* torch_utils.py:285-297 quat_to_tan_norm: rotate (1,0,0) then (0,0,1); tan first. */
static void o_tan_norm(q4 q, float out[6])
{
    v3 ex = { 1.0f, 0.0f, 0.0f }, ez = { 0.0f, 0.0f, 1.0f };
    v3 t = o_quat_rotate(q, ex), n = o_quat_rotate(q, ez);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = n.x; out[4] = n.y; out[5] = n.z;
}

/* torch_utils.py:369-380 calc_heading */
static float o_heading(q4 q)
{
    v3 ex = { 1.0f, 0.0f, 0.0f };
    v3 d = o_quat_rotate(q, ex);
    return atan2f(d.y, d.x);
}

/* torch_utils.py:354-358 quat_from_angle_axis with axis = (0,0,1) (:384-408):
 * normalize(axis) = (0,0,1)/max(1,1e-9); xyz = axis*sin(angle/2); w = cos(angle/2);
 * quat_unit divides by max(norm, 1e-9).  CPU torch.norm accumulates the squares with
 * an fma chain (hard part 1). */
static q4 o_quat_from_angle_z(float angle)
{
    float th = angle / 2.0f;
    float sn = sinf(th), cs = cosf(th);
    float x = 0.0f * sn, y = 0.0f * sn, z = 1.0f * sn, w = cs;
    float n = sqrtf(fmaf(w, w, fmaf(z, z, fmaf(y, y, x * x))));
    if (n < 1e-9f) n = 1e-9f;
    q4 r = { x / n, y / n, z / n, w / n };
    return r;
}

/* torch_utils.py:50-51 normalize_angle */
static float o_normalize_angle(float a) { return atan2f(sinf(a), cosf(a)); }

/* torch_utils.py:86-106 quat_to_angle_axis */
static void o_quat_to_angle_axis(q4 q, float *angle, v3 *axis)
{
    float s = sqrtf(1.0f - q.w * q.w);
    float ang = o_normalize_angle(2.0f * acosf(q.w));
    int mask = fabsf(s) > 1e-5f;            /* NaN -> false, like torch */
    if (mask) {
        axis->x = q.x / s; axis->y = q.y / s; axis->z = q.z / s;
        *angle = ang;
    } else {
        axis->x = 0.0f; axis->y = 0.0f; axis->z = 1.0f;
        *angle = 0.0f;
    }
}

/* torch_utils.py:144-150 quat_to_exp_map */
static v3 o_quat_to_exp_map(q4 q)
{
    float ang; v3 ax;
    o_quat_to_angle_axis(q, &ang, &ax);
    v3 r = { ang * ax.x, ang * ax.y, ang * ax.z };
    return r;
}

/* REFERENCE DEVICE FLAVOUR (test infrastructure state): torch rounds three small reductions of this path differently on CPU and on
 * CUDA; both orders were fitted bit-exactly against torch 2.11 on the B200 box (profiles/r2_torch_device_flavours.md):
 *   torch.sum(q0*q1,-1) over 4 (torch_utils.py:113)   CPU ((p0+p1)+p2)+p3              CUDA (p0+p2)+(p1+p3)
 *   torch.norm(d,dim=-1) over 3 (common.py:343)        CPU sqrt(fma(z,z,fma(y,y,x*x)))  CUDA sqrt((x*x+z*z)+y*y)
 *   .mean(-1) over J (common.py:344)                   CPU 8 lanes + tail first, / J    CUDA lane tree * float(1/J)
 *   tensor / python scalar (motion_lib.py:533)         CPU IEEE division                CUDA * float(1.0/scalar)
 * 0 = CPU (the device the golden vectors were made on), 1 = CUDA. */
static int g_ref_device = 0;
void phc_oracle_set_ref_device(int dev) { g_ref_device = dev ? 1 : 0; }
int phc_oracle_get_ref_device(void) { return g_ref_device; }

static float o_dot4(q4 a, q4 b)
{
    float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z, p3 = a.w * b.w;
    return g_ref_device ? (p0 + p2) + (p1 + p3) : ((p0 + p1) + p2) + p3;
}

/* .mean(-1) over n <= 32 values in the selected device's order */
static float o_mean(const float *v, int n)
{
    if (!g_ref_device) {   /* torch CPU: 8 vector lanes over the full rows of 8, scalar tail first, then the lanes one by one; / n */
        const int full = (n / 8) * 8;
        float s = 0.0f;
        for (int i = full; i < n; ++i) s = (i == full) ? v[i] : s + v[i];
        for (int k = 0; k < 8 && k < full; ++k) {
            float a = v[k];
            for (int i = k + 8; i < full; i += 8) a = a + v[i];
            s = (k == 0 && full == n) ? a : s + a;
        }
        return s / (float)n;
    }
    int bx = 1;
    while (bx * 2 <= n) bx *= 2;
    float t[32];
    for (int k = 0; k < bx; ++k) t[k] = (k + bx < n) ? v[k] + v[k + bx] : v[k];
    for (int off = bx / 2; off >= 1; off /= 2)
        for (int k = 0; k < off; ++k) t[k] = t[k] + t[k + off];
    return t[0] * (1.0f / (float)n);
}

/* torch_utils.py:110-131 slerp.  No renormalisation; the two torch.where fall-backs
 * are applied in the reference's order (lerp when |sin|<1e-3, then q0 when |cos|>=1). */
static q4 o_slerp(q4 q0, q4 q1, float t)
{
    float c = o_dot4(q0, q1);
    if (c < 0.0f) { q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z; q1.w = -q1.w; }
    c = fabsf(c);
    float h = acosf(c);
    float s = sqrtf(1.0f - c * c);
    float ra = sinf((1.0f - t) * h) / s;
    float rb = sinf(t * h) / s;
    q4 r = { ra * q0.x + rb * q1.x, ra * q0.y + rb * q1.y, ra * q0.z + rb * q1.z, ra * q0.w + rb * q1.w };
    if (fabsf(s) < 0.001f) {
        r.x = 0.5f * q0.x + 0.5f * q1.x; r.y = 0.5f * q0.y + 0.5f * q1.y;
        r.z = 0.5f * q0.z + 0.5f * q1.z; r.w = 0.5f * q0.w + 0.5f * q1.w;
    }
    if (fabsf(c) >= 1.0f) r = q0;
    return r;
}

static float o_clip01(float x) { /* torch.clip(x,0,1): NaN propagates */
    if (x != x) return x;
    return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
}

/* ------------------------------------------------------------------------------------ */
/* puffer_phc/motion_lib.py                                                               */
/* ------------------------------------------------------------------------------------ */

/* motion_lib.py:655-665 _calc_frame_blend (scalar). */
static void o_frame_blend(float time, float len, int64_t nf, float dt,
                          int64_t *i0, int64_t *i1, float *blend)
{
    float phase = o_clip01(time / len);
    if (time < 0.0f) time = 0.0f;
    int64_t a = (int64_t)(phase * (float)(nf - 1));
    int64_t b = a + 1 < nf - 1 ? a + 1 : nf - 1;
    *blend = o_clip01((time - (float)a * dt) / dt);
    *i0 = a; *i1 = b;
}

int phc_oracle_frame_blend(const float *time, const float *len, const int64_t *nf, const float *dt,
                           int64_t n, int64_t *i0, int64_t *i1, float *blend)
{
    for (int64_t i = 0; i < n; ++i) o_frame_blend(time[i], len[i], nf[i], dt[i], &i0[i], &i1[i], &blend[i]);
    return 0;
}

typedef struct {
    const float *gts, *grs, *lrs, *gvs, *gavs;  /* [F,24,3|4] */
    const float *dvs;                            /* [F,23,3]   */
    const float *motion_aa;                      /* [F,72]     */
    const float *motion_len, *motion_dt;         /* [M]        */
    const int64_t *num_frames, *length_starts;   /* [M]        */
    const float *motion_bodies;                  /* [M,17]     */
    const float *limb_weights;                   /* [M,10]     */
    int64_t F, M;
} phc_oracle_tables;

typedef struct {
    float *root_pos, *root_rot, *dof_pos, *root_vel, *root_ang_vel, *dof_vel, *motion_aa;
    float *rg_pos, *rb_rot, *body_vel, *body_ang_vel, *motion_bodies, *motion_limb_weights;
    int64_t *idx0, *idx1; float *blend;          /* optional debug outputs (may be NULL) */
} phc_oracle_state_out;

static float o_lerp(float a, float b, float one_m, float t) { return one_m * a + t * b; }

/* motion_lib.py:549-626 get_motion_state for one query; any output pointer may be NULL. */
static void o_motion_state_one(const phc_oracle_tables *T, int64_t id, float time, const float *off,
                               float *rg_pos /*72*/, float *rb_rot /*96*/, float *bvel /*72*/, float *bang /*72*/,
                               float *dof_pos /*69*/, float *dof_vel /*69*/, float *maa /*72*/,
                               int64_t *oi0, int64_t *oi1, float *oblend)
{
    int64_t i0, i1; float blend;
    o_frame_blend(time, T->motion_len[id], T->num_frames[id], T->motion_dt[id], &i0, &i1, &blend);
    int64_t f0 = i0 + T->length_starts[id], f1 = i1 + T->length_starts[id];
    float one_m = 1.0f - blend;
    if (oi0) *oi0 = i0;
    if (oi1) *oi1 = i1;
    if (oblend) *oblend = blend;
    for (int j = 0; j < NB; ++j) {
        for (int k = 0; k < 3; ++k) {
            if (rg_pos) {
                float p = o_lerp(T->gts[(f0 * NB + j) * 3 + k], T->gts[(f1 * NB + j) * 3 + k], one_m, blend);
                if (off) p = p + off[k];
                rg_pos[j * 3 + k] = p;
            }
            if (bvel) bvel[j * 3 + k] = o_lerp(T->gvs[(f0 * NB + j) * 3 + k], T->gvs[(f1 * NB + j) * 3 + k], one_m, blend);
            if (bang) bang[j * 3 + k] = o_lerp(T->gavs[(f0 * NB + j) * 3 + k], T->gavs[(f1 * NB + j) * 3 + k], one_m, blend);
        }
        if (rb_rot) {
            const float *a = &T->grs[(f0 * NB + j) * 4], *b = &T->grs[(f1 * NB + j) * 4];
            q4 qa = { a[0], a[1], a[2], a[3] }, qb = { b[0], b[1], b[2], b[3] };
            q4 r = o_slerp(qa, qb, blend);
            rb_rot[j * 4 + 0] = r.x; rb_rot[j * 4 + 1] = r.y; rb_rot[j * 4 + 2] = r.z; rb_rot[j * 4 + 3] = r.w;
        }
        if (dof_pos && j >= 1) {          /* motion_lib.py:605-606, 670-673 */
            const float *a = &T->lrs[(f0 * NB + j) * 4], *b = &T->lrs[(f1 * NB + j) * 4];
            q4 qa = { a[0], a[1], a[2], a[3] }, qb = { b[0], b[1], b[2], b[3] };
            v3 e = o_quat_to_exp_map(o_slerp(qa, qb, blend));
            dof_pos[(j - 1) * 3 + 0] = e.x; dof_pos[(j - 1) * 3 + 1] = e.y; dof_pos[(j - 1) * 3 + 2] = e.z;
        }
    }
    if (dof_vel)
        for (int k = 0; k < NDOF_J * 3; ++k)
            dof_vel[k] = o_lerp(T->dvs[f0 * NDOF_J * 3 + k], T->dvs[f1 * NDOF_J * 3 + k], one_m, blend);
    if (maa) memcpy(maa, &T->motion_aa[f0 * 72], 72 * sizeof(float));   /* :619 not blended */
}

int phc_oracle_motion_state(const phc_oracle_tables *T, const int64_t *ids, const float *times,
                            const float *offset /* [B,3] or NULL */, int64_t B, phc_oracle_state_out *o)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B; ++i) {
        int64_t id = ids[i];
        o_motion_state_one(T, id, times[i], offset ? &offset[i * 3] : NULL,
                           o->rg_pos + i * 72, o->rb_rot + i * 96, o->body_vel + i * 72, o->body_ang_vel + i * 72,
                           o->dof_pos + i * 69, o->dof_vel + i * 69, o->motion_aa + i * 72,
                           o->idx0 ? &o->idx0[i] : NULL, o->idx1 ? &o->idx1[i] : NULL, o->blend ? &o->blend[i] : NULL);
        memcpy(o->root_pos + i * 3, o->rg_pos + i * 72, 3 * sizeof(float));          /* :613 */
        memcpy(o->root_rot + i * 4, o->rb_rot + i * 96, 4 * sizeof(float));          /* :614 */
        memcpy(o->root_vel + i * 3, o->body_vel + i * 72, 3 * sizeof(float));        /* :616 */
        memcpy(o->root_ang_vel + i * 3, o->body_ang_vel + i * 72, 3 * sizeof(float));/* :617 */
        memcpy(o->motion_bodies + i * 17, T->motion_bodies + id * 17, 17 * sizeof(float));       /* :624 */
        memcpy(o->motion_limb_weights + i * 10, T->limb_weights + id * 10, 10 * sizeof(float)); /* :625 */
    }
    return 0;
}

/* motion_lib.py:526-535 sample_time_interval arithmetic (the torch.rand phase is an input).
 * div_mode 0: CPU torch true division by float32(1/30); 1: torch-CUDA's scalar fast path,
 * multiplication by float(1.0 / (1/30)) = 30.0f (the reciprocal is formed in double; fitted bit-exactly on the box). */
int phc_oracle_sample_time_interval(const float *phase, const float *motion_len, int64_t n, int div_mode, float *out)
{
    const float fps = (float)(1.0 / 30.0);
    const float inv = (float)(1.0 / (1.0 / 30.0));
    for (int64_t i = 0; i < n; ++i) {
        float x = phase[i] * motion_len[i];
        float q = div_mode ? x * inv : x / fps;
        out[i] = (float)(int64_t)q * fps;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* puffer_phc/envs/common.py                                                              */
/* ------------------------------------------------------------------------------------ */

/* common.py:15-19 remove_base_rot */
static q4 o_remove_base_rot(q4 q)
{
    q4 base = { 0.5f, 0.5f, 0.5f, 0.5f };
    return o_quat_mul(q, o_quat_conj(base));
}

static q4 ldq(const float *p) { q4 q = { p[0], p[1], p[2], p[3] }; return q; }
static v3 ldv(const float *p) { v3 v = { p[0], p[1], p[2] }; return v; }
static v3 v3sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }

/* common.py:106-176 compute_imitation_observations_v6, time_steps = 1.
 * Inputs are contiguous [N,J,3|4]; obs is [N, J*24] in six body-major blocks
 * [3J, 6J, 3J, 3J, 3J, 6J] (:168-175). */
int phc_oracle_imitation_obs_v6(const float *root_pos, const float *root_rot,
                                const float *body_pos, const float *body_rot, const float *body_vel, const float *body_ang,
                                const float *ref_pos, const float *ref_rot, const float *ref_vel, const float *ref_ang,
                                int64_t N, int J, int upright, float *obs)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        q4 rr = ldq(root_rot + i * 4);
        if (!upright) rr = o_remove_base_rot(rr);
        float hd = o_heading(rr);
        q4 hinv = o_quat_from_angle_z(-hd);       /* :130 */
        q4 h = o_quat_from_angle_z(hd);           /* :131 */
        v3 rp = ldv(root_pos + i * 3);
        float *o = obs + i * (int64_t)J * 24;
        float *b0 = o, *b1 = b0 + 3 * J, *b2 = b1 + 6 * J, *b3 = b2 + 3 * J, *b4 = b3 + 3 * J, *b5 = b4 + 3 * J;
        for (int j = 0; j < J; ++j) {
            int64_t e = i * J + j;
            v3 p = ldv(body_pos + e * 3), rpj = ldv(ref_pos + e * 3);
            v3 d = o_quat_rotate(hinv, v3sub(rpj, p));                       /* :138-139 */
            b0[j * 3] = d.x; b0[j * 3 + 1] = d.y; b0[j * 3 + 2] = d.z;
            q4 dq = o_quat_mul(ldq(ref_rot + e * 4), o_quat_conj(ldq(body_rot + e * 4)));   /* :142-145 */
            q4 lq = o_quat_mul(o_quat_mul(hinv, dq), h);                     /* :146-149 */
            o_tan_norm(lq, b1 + j * 6);                                      /* :169 */
            v3 dv = o_quat_rotate(hinv, v3sub(ldv(ref_vel + e * 3), ldv(body_vel + e * 3)));   /* :152-153 */
            b2[j * 3] = dv.x; b2[j * 3 + 1] = dv.y; b2[j * 3 + 2] = dv.z;
            v3 da = o_quat_rotate(hinv, v3sub(ldv(ref_ang + e * 3), ldv(body_ang + e * 3)));   /* :155-156 */
            b3[j * 3] = da.x; b3[j * 3 + 1] = da.y; b3[j * 3 + 2] = da.z;
            v3 lp = o_quat_rotate(hinv, v3sub(rpj, rp));                     /* :159-162 */
            b4[j * 3] = lp.x; b4[j * 3 + 1] = lp.y; b4[j * 3 + 2] = lp.z;
            o_tan_norm(o_quat_mul(hinv, ldq(ref_rot + e * 4)), b5 + j * 6);  /* :164-165 */
        }
    }
    return 0;
}

/* common.py:23-103 compute_humanoid_observations_smpl_max without smpl/limb params.
 * obs is [N, (root_height_obs?1:0) + 3(J-1) + 6J + 3J + 3J]. */
int phc_oracle_self_obs(const float *body_pos, const float *body_rot, const float *body_vel, const float *body_ang,
                        int64_t N, int J, int local_root_obs, int root_height_obs, int upright, float *obs)
{
    int W = (root_height_obs ? 1 : 0) + 3 * (J - 1) + 12 * J;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        q4 rr = ldq(body_rot + i * J * 4);
        v3 rp = ldv(body_pos + i * J * 3);
        if (!upright) rr = o_remove_base_rot(rr);                  /* :41-42 */
        q4 hinv = o_quat_from_angle_z(-o_heading(rr));             /* :43 */
        float *o = obs + i * (int64_t)W;
        if (root_height_obs) *o++ = rp.z;                          /* :40, :92-93 */
        float *bp = o, *br = bp + 3 * (J - 1), *bv = br + 6 * J, *ba = bv + 3 * J;
        for (int j = 0; j < J; ++j) {
            int64_t e = i * J + j;
            if (j >= 1) {                                          /* :57-66 */
                v3 d = o_quat_rotate(hinv, v3sub(ldv(body_pos + e * 3), rp));
                bp[(j - 1) * 3] = d.x; bp[(j - 1) * 3 + 1] = d.y; bp[(j - 1) * 3 + 2] = d.z;
            }
            o_tan_norm(o_quat_mul(hinv, ldq(body_rot + e * 4)), br + j * 6);   /* :68-75 */
            v3 v = o_quat_rotate(hinv, ldv(body_vel + e * 3));     /* :81-83 */
            bv[j * 3] = v.x; bv[j * 3 + 1] = v.y; bv[j * 3 + 2] = v.z;
            v3 a = o_quat_rotate(hinv, ldv(body_ang + e * 3));     /* :85-89 */
            ba[j * 3] = a.x; ba[j * 3 + 1] = a.y; ba[j * 3 + 2] = a.z;
        }
        if (!local_root_obs) o_tan_norm(rr, br);                   /* :77-79 */
    }
    return 0;
}

/* common.py:270-322 compute_imitation_reward.  k/w order: pos, rot, vel, ang_vel.
 * Means are sum/count with left-to-right sums. */
int phc_oracle_imitation_reward(const float *body_pos, const float *body_rot, const float *body_vel, const float *body_ang,
                                const float *ref_pos, const float *ref_rot, const float *ref_vel, const float *ref_ang,
                                int64_t N, int J, const float *k, const float *w, float *reward, float *reward_raw)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        float sp = 0.0f, sr = 0.0f, sv = 0.0f, sa = 0.0f;
        for (int j = 0; j < J; ++j) {
            int64_t e = i * J + j;
            v3 d = v3sub(ldv(ref_pos + e * 3), ldv(body_pos + e * 3));
            sp += ((d.x * d.x + d.y * d.y) + d.z * d.z) / 3.0f;                    /* :299-300 */
            float ang; v3 ax;
            o_quat_to_angle_axis(o_quat_mul(ldq(ref_rot + e * 4), o_quat_conj(ldq(body_rot + e * 4))), &ang, &ax);
            sr += ang * ang;                                                        /* :304-306 */
            v3 dv = v3sub(ldv(ref_vel + e * 3), ldv(body_vel + e * 3));
            sv += ((dv.x * dv.x + dv.y * dv.y) + dv.z * dv.z) / 3.0f;               /* :310-311 */
            v3 da = v3sub(ldv(ref_ang + e * 3), ldv(body_ang + e * 3));
            sa += ((da.x * da.x + da.y * da.y) + da.z * da.z) / 3.0f;               /* :315-316 */
        }
        float r0 = expf(-k[0] * (sp / (float)J));
        float r1 = expf(-k[1] * (sr / (float)J));
        float r2 = expf(-k[2] * (sv / (float)J));
        float r3 = expf(-k[3] * (sa / (float)J));
        reward[i] = ((w[0] * r0 + w[1] * r1) + w[2] * r2) + w[3] * r3;              /* :319 */
        reward_raw[i * 4 + 0] = r0; reward_raw[i * 4 + 1] = r1; reward_raw[i * 4 + 2] = r2; reward_raw[i * 4 + 3] = r3;
    }
    return 0;
}

static float o_norm3(v3 d)
{
    if (g_ref_device) return sqrtf((d.x * d.x + d.z * d.z) + d.y * d.y);
    return sqrtf(fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x)));
}

/* common.py:325-364 compute_humanoid_im_reset. progress is int16 (humanoid_phc.py:571). */
int phc_oracle_im_reset(const int16_t *progress, const float *body_pos, const float *ref_pos, const uint8_t *pass_time,
                        int enable_early_termination, const float *term_dist /* [J] */, int use_mean,
                        int64_t N, int J, uint8_t *reset, uint8_t *terminated)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        int fallen = 0;
        if (enable_early_termination) {
            if (use_mean) {                                                         /* :342-346 */
                float dj[32];
                for (int j = 0; j < J; ++j) dj[j] = o_norm3(v3sub(ldv(body_pos + (i * J + j) * 3), ldv(ref_pos + (i * J + j) * 3)));
                fallen = o_mean(dj, J) > term_dist[0];
            } else {                                                                /* :347-350 */
                for (int j = 0; j < J; ++j)
                    fallen |= o_norm3(v3sub(ldv(body_pos + (i * J + j) * 3), ldv(ref_pos + (i * J + j) * 3))) > term_dist[j];
            }
            fallen = fallen && (progress[i] > 1);                                   /* :354 */
        }
        terminated[i] = (uint8_t)fallen;                                            /* :356 */
        reset[i] = pass_time[i] ? 1 : (uint8_t)fallen;                              /* :362 */
    }
    return 0;
}

/* humanoid_phc.py:159-163: extras["mpjpe"] = (body_pos - rg_pos).norm(dim=-1).mean(dim=-1) */
int phc_oracle_mpjpe(const float *body_pos, const float *ref_pos, int64_t N, int J, float *out)
{
    for (int64_t i = 0; i < N; ++i) {
        float dj[32];
        for (int j = 0; j < J; ++j) dj[j] = o_norm3(v3sub(ldv(body_pos + (i * J + j) * 3), ldv(ref_pos + (i * J + j) * 3)));
        out[i] = o_mean(dj, J);
    }
    return 0;
}

/* torch_utils.py:333-365 exp_map_to_quat = quat_from_angle_axis(exp_map_to_angle_axis(exp_map)) */
static q4 o_exp_map_to_quat(v3 e)
{
    float angle = o_norm3(e);
    v3 axis = { e.x / angle, e.y / angle, e.z / angle };
    angle = o_normalize_angle(angle);
    if (!(fabsf(angle) > 1e-5f)) { angle = 0.0f; axis.x = 0.0f; axis.y = 0.0f; axis.z = 1.0f; }
    float th = angle / 2.0f;
    float an = o_norm3(axis);                       /* normalize(axis): x / max(norm, 1e-9) (:44-46) */
    if (an < 1e-9f) an = 1e-9f;
    float sn = sinf(th);
    float x = (axis.x / an) * sn, y = (axis.y / an) * sn, z = (axis.z / an) * sn, w = cosf(th);
    float n = sqrtf(fmaf(w, w, fmaf(z, z, fmaf(y, y, x * x))));
    if (n < 1e-9f) n = 1e-9f;
    q4 r = { x / n, y / n, z / n, w / n };
    return r;
}

/* common.py:192-267 build_amp_observations_smpl (+ dof_to_obs_smpl :179-189) without the shape / limb pass-through columns.
 * dof_subset: 3*nj indices into the 69-dof vector (or NULL = all 23 joints); key_body_pos [N,K,3].
 * obs: [N, (root_height_obs?1:0) + 12 + 9*nj + 3*K]. */
int phc_oracle_amp_obs(const float *root_pos, const float *root_rot, const float *root_vel, const float *root_ang,
                       const float *dof_pos, const float *dof_vel, const float *key_pos, const int64_t *dof_subset, int nj,
                       int K, int local_root_obs, int root_height_obs, int upright, int64_t N, float *obs)
{
    int W = (root_height_obs ? 1 : 0) + 12 + 9 * nj + 3 * K;
    for (int64_t i = 0; i < N; ++i) {
        float *o = obs + i * (int64_t)W;
        q4 rr = ldq(root_rot + i * 4);
        v3 rp = ldv(root_pos + i * 3);
        if (!upright) rr = o_remove_base_rot(rr);                                   /* :214-215 */
        q4 hinv = o_quat_from_angle_z(-o_heading(rr));                              /* :216 */
        if (root_height_obs) *o++ = rp.z;                                           /* :213, 250-251 */
        o_tan_norm(local_root_obs ? o_quat_mul(hinv, rr) : rr, o);                  /* :218-223 */
        v3 v = o_quat_rotate(hinv, ldv(root_vel + i * 3));                          /* :225 */
        v3 w = o_quat_rotate(hinv, ldv(root_ang + i * 3));                          /* :226 */
        o[6] = v.x; o[7] = v.y; o[8] = v.z; o[9] = w.x; o[10] = w.y; o[11] = w.z;
        float *dobs = o + 12, *dvel = dobs + 6 * nj, *kp = dvel + 3 * nj;
        for (int j = 0; j < nj; ++j) {
            v3 e;
            float dv[3];
            for (int c = 0; c < 3; ++c) {
                int64_t idx = dof_subset ? dof_subset[3 * j + c] : 3 * j + c;       /* :244-246 */
                ((float *)&e)[c] = dof_pos[i * 69 + idx];
                dv[c] = dof_vel[i * 69 + idx];
            }
            o_tan_norm(o_exp_map_to_quat(e), dobs + 6 * j);                         /* :186, 248 */
            dvel[3 * j] = dv[0]; dvel[3 * j + 1] = dv[1]; dvel[3 * j + 2] = dv[2];
        }
        for (int k = 0; k < K; ++k) {                                               /* :228-242 */
            v3 d = o_quat_rotate(hinv, v3sub(ldv(key_pos + (i * K + k) * 3), rp));
            kp[3 * k] = d.x; kp[3 * k + 1] = d.y; kp[3 * k + 2] = d.z;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* The per-step glue of HumanoidPHC.step() after physics (puffer_phc/envs/humanoid_phc.py */
/* :136-149): _compute_reward :1228-1303, _compute_reset :1311-1333,                      */
/* _compute_observations :935-959 (_compute_humanoid_obs :961-991, _compute_task_obs      */
/* :1048-1112).  body_state is the PhysX AoS record [N, env_stride floats], 13 floats per */
/* body: pos3, rot4, vel3, angvel3 (:542-549).                                            */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const float *body_state; int64_t env_stride;
    const int16_t *progress; const float *start_time, *start_offset; const int64_t *motion_ids;
    const float *global_offset;                 /* [N,3] */
    const float *dof_force, *dof_vel;           /* [N,69] or NULL (power reward off) */
    float dt;                                   /* float32(isaac dt) */
    float k[4], w[4]; float power_coef;
    const float *term_dist;                     /* [24], indexed by body id */
    uint32_t reset_body_mask;                   /* bit j set = body j is a reset body */
    int enable_early_termination, use_mean;
    int64_t N;
} phc_oracle_step_in;

typedef struct {
    float *obs;            /* [N,934] */
    float *reward;         /* [N] */
    float *reward_raw;     /* [N,5] when power reward on, else [N,4] */
    uint8_t *reset, *terminated;
    float *ref_t;          /* optional [N,312]: pos72 rot96 vel72 ang72 at t   */
    float *ref_t1;         /* optional [N,312] at t+1 */
} phc_oracle_step_out;

int phc_oracle_step(const phc_oracle_tables *T, const phc_oracle_step_in *in, phc_oracle_step_out *out)
{
    const int raw_w = in->dof_force ? 5 : 4;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < in->N; ++i) {
        const float *rec = in->body_state + i * in->env_stride;
        float bp[72], br[96], bv[72], ba[72];
        for (int j = 0; j < NB; ++j) {
            const float *r = rec + j * 13;
            memcpy(bp + j * 3, r, 12); memcpy(br + j * 4, r + 3, 16); memcpy(bv + j * 3, r + 7, 12); memcpy(ba + j * 3, r + 10, 12);
        }
        int64_t id = in->motion_ids[i];
        const float *off = in->global_offset + i * 3;
        int16_t prog = in->progress[i];
        /* humanoid_phc.py:1233-1235: progress_buf * dt + start + start_offset, fp32 */
        float t0 = ((float)prog * in->dt + in->start_time[i]) + in->start_offset[i];
        /* humanoid_phc.py:1060-1064: (progress_buf + 1) stays int16, then * dt */
        float t1 = ((float)(int16_t)(prog + 1) * in->dt + in->start_time[i]) + in->start_offset[i];
        float rp0[72], rr0[96], rv0[72], ra0[72], rp1[72], rr1[96], rv1[72], ra1[72];
        o_motion_state_one(T, id, t0, off, rp0, rr0, rv0, ra0, NULL, NULL, NULL, NULL, NULL, NULL);
        o_motion_state_one(T, id, t1, off, rp1, rr1, rv1, ra1, NULL, NULL, NULL, NULL, NULL, NULL);
        if (out->ref_t) { float *d = out->ref_t + i * 312; memcpy(d, rp0, 288); memcpy(d + 72, rr0, 384); memcpy(d + 168, rv0, 288); memcpy(d + 240, ra0, 288); }
        if (out->ref_t1) { float *d = out->ref_t1 + i * 312; memcpy(d, rp1, 288); memcpy(d + 72, rr1, 384); memcpy(d + 168, rv1, 288); memcpy(d + 240, ra1, 288); }

        /* reward at t (:1257-1270) */
        float raw4[4];
        phc_oracle_imitation_reward(bp, br, bv, ba, rp0, rr0, rv0, ra0, 1, NB, in->k, in->w, &out->reward[i], raw4);
        memcpy(out->reward_raw + i * raw_w, raw4, 16);
        if (in->dof_force) {                                         /* :1295-1303 */
            float power = 0.0f;
            for (int d = 0; d < 69; ++d) power += fabsf(in->dof_force[i * 69 + d] * in->dof_vel[i * 69 + d]);
            float pr = -in->power_coef * power;
            if (prog <= 3) pr = 0.0f;
            out->reward[i] = out->reward[i] + pr;
            out->reward_raw[i * raw_w + 4] = pr;
        }

        /* reset at t (:1311-1333): gather the reset bodies in index order */
        float sp[72], sr[72], td[NB]; int nj = 0;
        for (int j = 0; j < NB; ++j)
            if (in->reset_body_mask >> j & 1u) {
                memcpy(sp + nj * 3, bp + j * 3, 12); memcpy(sr + nj * 3, rp0 + j * 3, 12); td[nj] = in->term_dist[j]; ++nj;
            }
        uint8_t pass = t0 >= T->motion_len[id];                      /* :1315 */
        phc_oracle_im_reset(&prog, sp, sr, &pass, in->enable_early_termination, td, in->use_mean, 1, nj,
                            &out->reset[i], &out->terminated[i]);

        /* obs: self (358) then task (576) with the reference at t+1 (:947, :979-991, :1099-1112) */
        float *o = out->obs + i * 934;
        phc_oracle_self_obs(bp, br, bv, ba, 1, NB, 1, 1, 1, o);
        phc_oracle_imitation_obs_v6(bp, br, bp, br, bv, ba, rp1, rr1, rv1, ra1, 1, NB, 1, o + 358);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* puffer_phc/policies/running_norm.py                                                    */
/* ------------------------------------------------------------------------------------ */

/* running_norm.py:15-20 forward */
int phc_oracle_rms_forward(const float *x, const float *mean, const float *var, float eps, float clip,
                           int64_t B, int C, float *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B; ++i)
        for (int c = 0; c < C; ++c) {
            float v = (x[i * C + c] - mean[c]) / sqrtf(var[c] + eps);
            /* torch.clamp(v, -clip, clip): min(max(v, lo), hi), NaN propagates */
            if (v == v) v = v < -clip ? -clip : (v > clip ? clip : v);
            y[i * C + c] = v;
        }
    return 0;
}

/* running_norm.py:23-34 update: batch mean and biased variance (accumulated in double,
 * rounded to fp32), then running = running*(1-w) + batch*w with w = 1/count; count += 1. */
int phc_oracle_rms_update(const float *x, int64_t B, int C, float *running_mean, float *running_var, float *count)
{
    float wgt = 1.0f / count[0];
    for (int c = 0; c < C; ++c) {
        double s = 0.0;
        for (int64_t i = 0; i < B; ++i) s += (double)x[i * C + c];
        double m = s / (double)B, q = 0.0;
        for (int64_t i = 0; i < B; ++i) { double d = (double)x[i * C + c] - m; q += d * d; }
        float mean = (float)m, var = (float)(q / (double)B);
        running_mean[c] = running_mean[c] * (1.0f - wgt) + mean * wgt;
        running_var[c] = running_var[c] * (1.0f - wgt) + var * wgt;
    }
    count[0] += 1.0f;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* puffer_phc/c_gae.pyx:11-32 compute_gae -- flat serial reverse scan, all locals float.  */
/* ------------------------------------------------------------------------------------ */
int phc_oracle_gae(const float *dones, const float *values, const float *rewards, int64_t L,
                   float gamma, float gae_lambda, float *adv)
{
    for (int64_t i = 0; i < L; ++i) adv[i] = 0.0f;
    float last = 0.0f;
    for (int64_t t = 0; t + 1 < L; ++t) {
        int64_t cur = L - 2 - t, nxt = L - 1 - t;
        float nnt = 1.0f - dones[nxt];
        float delta = (rewards[nxt] + (gamma * values[nxt]) * nnt) - values[cur];
        last = delta + ((gamma * gae_lambda) * nnt) * last;
        adv[cur] = last;
    }
    return 0;
}
