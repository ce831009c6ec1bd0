/*
 * loader_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, scalar, CPU restatement of the motion TABLE BUILD of howird/puffer-phc (SURVEY.md
 * section 8 row f4): what MotionLibSMPL.load_motion_with_skeleton + SkeletonMotion.from_skeleton_state
 * + compute_motion_dof_vels_jit turn one raw clip into.  Used only as the parity checker of
 * csrc/build_tables.cu.  Nothing under puffer_phc_b200/ may import, link or call this file.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py (make_loader) runs the reference's own loader on
 * the real sample clip and three synthetic clips in the build container and commits the raw clips and
 * the tables it built (tests/golden/loader.npz); tests/test_oracle_golden.py holds this file to them.
 *
 * The reference mixes precisions, and this file follows it operation by operation:
 *   - the input rotations are GLOBAL and float64 (pkl format, scripts/convert_amass_data.py:186-196);
 *   - local rotations are computed in float64 but stored into a float32 tensor
 *     (poselib_skeleton.py:575-591: quat_identity_like() allocates float32);
 *   - forward kinematics therefore runs entirely in float32 (local rotation f32, skeleton offsets f32,
 *     the f64 root translation rounded on assignment; poselib_skeleton.py:516-536, 603-617);
 *   - linear velocity = np.gradient of the float32 positions, divided by float32(1/fps), then
 *     scipy.ndimage gaussian_filter1d(sigma 2, mode nearest), which accumulates in double and stores
 *     float32 (poselib_skeleton.py:1228-1235);
 *   - angular velocity is float64 end to end (the global rotations are the float64 input), filtered the
 *     same way (poselib_skeleton.py:1238-1249);
 *   - dof velocities are float32 torch ops on the float32 local rotations (motion_lib.py:119-140).
 * scipy is a third-party dependency (scipy 1.18.1 in this image; any 1.x has the same correlate1d):
 * its published algorithm -- symmetric-kernel correlate1d over a "nearest"-extended line, kernel
 * exp(-x^2/(2 sigma^2)) over radius int(4 sigma + 0.5), normalised by its sum -- is restated below.
 *
 * Build with -ffp-contract=off (oracle/Makefile).  Quaternions are xyzw.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y, z, w; } dq4;
typedef struct { float x, y, z, w; } fq4;
typedef struct { float x, y, z; } fv3;

/* torch_utils.py:55-75 quat_mul, float64 instantiation (expression order kept verbatim). */
static dq4 d_quat_mul(dq4 a, dq4 b)
{
    double x1 = a.x, y1 = a.y, z1 = a.z, w1 = a.w, x2 = b.x, y2 = b.y, z2 = b.z, w2 = b.w;
    double ww = (z1 + x1) * (x2 + y2);
    double yy = (w1 - y1) * (w2 + z2);
    double zz = (w1 + y1) * (w2 - z2);
    double xx = (ww + yy) + zz;
    double qq = 0.5 * (xx + (z1 - x1) * (x2 - y2));
    dq4 r;
    r.w = (qq - ww) + (z1 - y1) * (y2 - z2);
    r.x = (qq - xx) + (x1 + w1) * (x2 + w2);
    r.y = (qq - yy) + (w1 - x1) * (y2 + z2);
    r.z = (qq - zz) + (z1 + y1) * (w2 - x2);
    return r;
}
static fq4 f_quat_mul(fq4 a, fq4 b)
{
    float x1 = a.x, y1 = a.y, z1 = a.z, w1 = a.w, x2 = b.x, y2 = b.y, z2 = b.z, w2 = b.w;
    float ww = (z1 + x1) * (x2 + y2);
    float yy = (w1 - y1) * (w2 + z2);
    float zz = (w1 + y1) * (w2 - z2);
    float xx = (ww + yy) + zz;
    float qq = 0.5f * (xx + (z1 - x1) * (x2 - y2));
    fq4 r;
    r.w = (qq - ww) + (z1 - y1) * (y2 - z2);
    r.x = (qq - xx) + (x1 + w1) * (x2 + w2);
    r.y = (qq - yy) + (w1 - x1) * (y2 + z2);
    r.z = (qq - zz) + (z1 + y1) * (w2 - x2);
    return r;
}
static dq4 d_conj(dq4 a) { dq4 r = { -a.x, -a.y, -a.z, a.w }; return r; }    /* torch_utils.py:79-82, 232-236 */
static fq4 f_conj(fq4 a) { fq4 r = { -a.x, -a.y, -a.z, a.w }; return r; }

/* torch_utils.py:154-196 quat_normalize = quat_unit(quat_pos(q)): flip the sign when w < 0, divide by
 * max(norm, 1e-9).  torch's CPU norm over a contiguous last dimension of 4 sums the squares left to
 * right, every product and sum rounded (probed in the build container; the 3-wide norm differs). */
static dq4 d_normalize(dq4 q)
{
    double s = (q.w < 0.0) ? -1.0 : 1.0;
    q.x *= s; q.y *= s; q.z *= s; q.w *= s;
    double n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    if (n < 1e-9) n = 1e-9;
    dq4 r = { q.x / n, q.y / n, q.z / n, q.w / n };
    return r;
}
static fq4 f_normalize(fq4 q)
{
    float s = (q.w < 0.0f) ? -1.0f : 1.0f;
    q.x *= s; q.y *= s; q.z *= s; q.w *= s;
    float n = sqrtf(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    if (n < 1e-9f) n = 1e-9f;
    fq4 r = { q.x / n, q.y / n, q.z / n, q.w / n };
    return r;
}

/* torch_utils.py:263-269 quat_rotate: imag(quat_mul(quat_mul(rot, [v, 0]), conj(rot))), float32. */
static fv3 f_quat_rotate(fq4 rot, fv3 v)
{
    fq4 o = { v.x, v.y, v.z, 0.0f };
    fq4 r = f_quat_mul(f_quat_mul(rot, o), f_conj(rot));
    fv3 out = { r.x, r.y, r.z };
    return out;
}

/* scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius): exp(-0.5/sigma^2 * x^2) / sum; the sum is
 * numpy's pairwise reduction (8 running sums over blocks of 8, combined as a tree, remainder appended). */
#define G_RADIUS 8      /* int(truncate * sigma + 0.5) = int(4 * 2 + 0.5) */
static void gaussian_weights(double w[2 * G_RADIUS + 1])
{
    const double sigma = 2.0, sigma2 = sigma * sigma;
    double r[8];
    for (int k = 0; k < 2 * G_RADIUS + 1; ++k) { double x = (double)(k - G_RADIUS); w[k] = exp(-0.5 / sigma2 * (x * x)); }
    for (int k = 0; k < 8; ++k) r[k] = w[k];
    for (int k = 0; k < 8; ++k) r[k] += w[8 + k];
    double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    s += w[16];
    for (int k = 0; k < 2 * G_RADIUS + 1; ++k) w[k] = w[k] / s;
}

/* scipy ni_filters.c NI_Correlate1D, symmetric branch, mode "nearest": the line is read as doubles,
 * tmp = x[l]*w[c]; for jj = -c..-1: tmp += (x[l+jj] + x[l-jj]) * w[jj+c].  `col` has stride `stride`. */
static double filter_at(const double *line, int64_t T, int64_t l, const double w[2 * G_RADIUS + 1])
{
    double tmp = line[l] * w[G_RADIUS];
    for (int jj = -G_RADIUS; jj < 0; ++jj) {
        int64_t a = l + jj, b = l - jj;
        if (a < 0) a = 0;
        if (b > T - 1) b = T - 1;
        tmp += (line[a] + line[b]) * w[jj + G_RADIUS];
    }
    return tmp;
}

/* torch_utils.py:86-106 quat_to_angle_axis + motion_lib.py:131-133, float32. */
static void f_dof_vel(fq4 l0, fq4 l1, float dt, float out[3])
{
    fq4 d = f_quat_mul(f_conj(l0), l1);
    float sin_theta = sqrtf(1.0f - d.w * d.w);
    float angle = 2.0f * acosf(d.w);
    angle = atan2f(sinf(angle), cosf(angle));                       /* normalize_angle :50-51 */
    float ax = d.x / sin_theta, ay = d.y / sin_theta, az = d.z / sin_theta;
    if (!(fabsf(sin_theta) > 1e-5f)) { angle = 0.0f; ax = 0.0f; ay = 0.0f; az = 1.0f; }   /* NaN -> default */
    out[0] = (ax * angle) / dt; out[1] = (ay * angle) / dt; out[2] = (az * angle) / dt;
}

/*
 * One clip.  pose_quat_global [T,J,4] f64, root_trans [T,3] f64 (already cropped to [start,end),
 * motion_lib.py:773-785), parents [J] (-1 for the root, parents precede children), local_translation
 * [J,3] f32 (SkeletonTree.local_translation), fps.  Outputs are the per-clip slices of the tables
 * load_motions concatenates (motion_lib.py:405-412): gts [T,J,3], grs [T,J,4], lrs [T,J,4], gvs [T,J,3],
 * gavs [T,J,3], dvs [T,J-1,3], all float32.  Returns 0, or -1 for T < 2 (np.gradient and the dof-velocity
 * loop both raise in the reference) / J > 64.
 */
int phc_oracle_build_clip(const double *pose_quat_global, const double *root_trans, int64_t T, int J,
                          const int64_t *parents, const float *local_translation, int fps,
                          float *gts, float *grs, float *lrs, float *gvs, float *gavs, float *dvs)
{
    if (T < 2 || J > 64 || J < 1) return -1;
    const int C = J * 3;
    double w[2 * G_RADIUS + 1];
    gaussian_weights(w);

    /* ---- rotations and forward kinematics -------------------------------------------------------- */
    for (int64_t t = 0; t < T; ++t) {
        const dq4 *G = (const dq4 *)(pose_quat_global + t * J * 4);
        fq4 *L = (fq4 *)(lrs + t * J * 4);
        fq4 *Gf = (fq4 *)(grs + t * J * 4);
        fv3 *P = (fv3 *)(gts + t * J * 3);
        fq4 Gr[64];
        for (int j = 0; j < J; ++j) {
            Gf[j].x = (float)G[j].x; Gf[j].y = (float)G[j].y; Gf[j].z = (float)G[j].z; Gf[j].w = (float)G[j].w;   /* grs :406 */
            /* poselib_skeleton.py:579-590: local = quat_mul_norm(quat_inverse(global[parent]), global[j]) in f64,
             * stored into a float32 tensor; the root keeps its global rotation. */
            dq4 l = (parents[j] < 0) ? G[j] : d_normalize(d_quat_mul(d_conj(G[parents[j]]), G[j]));
            L[j].x = (float)l.x; L[j].y = (float)l.y; L[j].z = (float)l.z; L[j].w = (float)l.w;
        }
        /* poselib_skeleton.py:516-536 global_transformation, torch_utils.py:322-330 transform_mul -- float32:
         * r = quat_mul_norm(r_parent, local); t = quat_rotate(r_parent, local_translation) + t_parent. */
        for (int j = 0; j < J; ++j) {
            int p = (int)parents[j];
            if (p < 0) {
                Gr[j] = L[j];
                P[j].x = (float)root_trans[t * 3 + 0]; P[j].y = (float)root_trans[t * 3 + 1]; P[j].z = (float)root_trans[t * 3 + 2];
            } else {
                fv3 lt = { local_translation[j * 3 + 0], local_translation[j * 3 + 1], local_translation[j * 3 + 2] };
                Gr[j] = f_normalize(f_quat_mul(Gr[p], L[j]));
                fv3 r = f_quat_rotate(Gr[p], lt);
                P[j].x = r.x + P[p].x; P[j].y = r.y + P[p].y; P[j].z = r.z + P[p].z;
            }
        }
    }

    /* ---- linear velocity: np.gradient (f32) / f32(1/fps), gaussian filter (double accumulate, f32 store) ---- */
    double *col = (double *)malloc(sizeof(double) * (size_t)T);
    if (!col) return -2;
    const double time_delta = 1.0 / (double)fps;                      /* poselib_skeleton.py:1183: 1 / fps */
    const float td_f = (float)time_delta;
    for (int c = 0; c < C; ++c) {
        for (int64_t t = 0; t < T; ++t) {
            float g;
            if (t == 0) g = (gts[1 * C + c] - gts[0 * C + c]) / 1.0f;
            else if (t == T - 1) g = (gts[(T - 1) * C + c] - gts[(T - 2) * C + c]) / 1.0f;
            else g = (gts[(t + 1) * C + c] - gts[(t - 1) * C + c]) / 2.0f;
            col[t] = (double)(g / td_f);
        }
        for (int64_t t = 0; t < T; ++t) gvs[t * C + c] = (float)filter_at(col, T, t, w);
    }

    /* ---- angular velocity, float64 (poselib_skeleton.py:1238-1249; torch_utils.py:219-228 quat_angle_axis) ---- */
    double *av = (double *)malloc(sizeof(double) * (size_t)T * (size_t)C);
    if (!av) { free(col); return -2; }
    for (int64_t t = 0; t < T; ++t) {
        for (int j = 0; j < J; ++j) {
            dq4 d = { 0.0, 0.0, 0.0, 1.0 };                             /* last frame keeps the identity */
            if (t < T - 1) {
                const dq4 *G0 = (const dq4 *)(pose_quat_global + t * J * 4), *G1 = (const dq4 *)(pose_quat_global + (t + 1) * J * 4);
                d = d_normalize(d_quat_mul(G1[j], d_conj(G0[j])));
            }
            double s = 2.0 * (d.w * d.w) - 1.0;
            if (s < -1.0) s = -1.0;
            if (s > 1.0) s = 1.0;
            double angle = acos(s);
            double n = sqrt((d.x * d.x + d.y * d.y) + d.z * d.z);
            if (n < 1e-9) n = 1e-9;
            av[(t * J + j) * 3 + 0] = ((d.x / n) * angle) / time_delta;
            av[(t * J + j) * 3 + 1] = ((d.y / n) * angle) / time_delta;
            av[(t * J + j) * 3 + 2] = ((d.z / n) * angle) / time_delta;
        }
    }
    for (int c = 0; c < C; ++c) {
        for (int64_t t = 0; t < T; ++t) col[t] = av[t * C + c];
        for (int64_t t = 0; t < T; ++t) gavs[t * C + c] = (float)filter_at(col, T, t, w);
    }
    free(av);
    free(col);

    /* ---- dof velocities (motion_lib.py:119-140): frames 0..T-2 from (t, t+1), the last repeats T-2 ---- */
    const float dt_f = (float)(1.0 / (double)fps);
    for (int64_t t = 0; t < T; ++t) {
        int64_t a = (t < T - 1) ? t : T - 2;
        const fq4 *L0 = (const fq4 *)(lrs + a * J * 4), *L1 = (const fq4 *)(lrs + (a + 1) * J * 4);
        for (int j = 1; j < J; ++j) f_dof_vel(L0[j], L1[j], dt_f, dvs + (t * (J - 1) + (j - 1)) * 3);
    }
    return 0;
}
