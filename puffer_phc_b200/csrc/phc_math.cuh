// phc_math.cuh -- fp32 quaternion / heading / frame-blend math of the PHC hot path.
//
// Every function mirrors the *operation order* of the reference's torch code (one rounding per
// torch op, no FMA contraction) so that integer results and branch decisions are reproducible
// bit for bit and floating-point results stay within 1e-5 of the reference.  The translation
// unit that includes this header MUST be compiled with -fmad=false (nvcc) / -ffp-contract=off
// (host harness); explicit fmaf() is used only where the reference's CPU kernel fuses.
// Quaternions are xyzw (reference puffer_phc/torch_utils.py:61-62).
//
// REFERENCE DEVICE FLAVOUR (`dev` arguments; PHC_REF_CPU = 0, PHC_REF_CUDA = 1).  torch rounds three reductions of this path
// differently on its two devices (measured on the B200 box with torch 2.11, profiles/r2_torch_device_flavours.md):
//     torch.sum(q0*q1, -1) over 4      CPU ((p0+p1)+p2)+p3                 CUDA (p0+p2)+(p1+p3)          (slerp, torch_utils.py:113)
//     torch.norm(d, dim=-1) over 3     CPU sqrt(fma(z,z,fma(y,y,x*x)))     CUDA sqrt((x*x+z*z)+y*y)      (termination test, common.py:343)
//     .mean(-1) over J                 CPU 8 lanes, tail first, / J        CUDA 16-lane tree * f32(1/J)  (eval variant, common.py:344)
// They decide flags (termination) and slerp's fall-back branches, and the dot product feeds a cancelling 1-c*c, so every
// function that contains one takes the flavour to reproduce; everything else is flavour-independent.
//
// The header compiles for the host as well (tests/host_math_harness.cpp) so the math can be
// checked on a CPU-only box before it ever runs on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PHC_HD __host__ __device__ __forceinline__
#else
#define PHC_HD inline
#endif

namespace phc {

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };

PHC_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }

constexpr int PHC_REF_CPU = 0, PHC_REF_CUDA = 1;

// torch.sum(q0 * q1, dim=-1) (products individually rounded on both devices; only the order of the three additions differs)
PHC_HD float dot4(Q4 a, Q4 b, int dev) {
    const float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z, p3 = a.w * b.w;
    return dev == PHC_REF_CUDA ? (p0 + p2) + (p1 + p3) : ((p0 + p1) + p2) + p3;
}

// torch.norm over 3 components: CPU sqrt(fma(z,z,fma(y,y,x*x))); CUDA sqrt((x*x + z*z) + y*y) (two reduction lanes: lane 0 takes
// components 0 and 2, lane 1 component 1, then one shuffle-add) -- both bit-exact against torch on 262144 random triples.
PHC_HD float norm3(V3 d, int dev = PHC_REF_CPU) {
    if (dev == PHC_REF_CUDA) return sqrtf((d.x * d.x + d.z * d.z) + d.y * d.y);
    return sqrtf(fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x)));
}

// torch.clip(x, 0, 1): NaN propagates (fminf/fmaxf would drop it).
PHC_HD float clip01(float x) {
    if (x != x) return x;
    return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
}

// MotionLibBase._calc_frame_blend (reference puffer_phc/motion_lib.py:655-665).
// phase uses the un-clamped time; only the lower clamp is applied to time afterwards.
PHC_HD void frame_blend(float time, float len, int64_t nf, float dt, int64_t& idx0, int64_t& idx1, float& blend) {
    float phase = clip01(time / len);
    if (time < 0.0f) time = 0.0f;
    idx0 = (int64_t)(phase * (float)(nf - 1));
    idx1 = idx0 + 1 < nf - 1 ? idx0 + 1 : nf - 1;
    blend = clip01((time - (float)idx0 * dt) / dt);
}

// quat_mul (torch_utils.py:55-75): the 8-multiply form, expression order kept.
PHC_HD Q4 quat_mul(Q4 a, Q4 b) {
    float ww = (a.z + a.x) * (b.x + b.y);
    float yy = (a.w - a.y) * (b.w + b.z);
    float zz = (a.w + a.y) * (b.w - b.z);
    float xx = (ww + yy) + zz;
    float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
    Q4 r;
    r.w = (qq - ww) + (a.z - a.y) * (b.y - b.z);
    r.x = (qq - xx) + (a.x + a.w) * (b.x + b.w);
    r.y = (qq - yy) + (a.w - a.x) * (b.y + b.z);
    r.z = (qq - zz) + (a.z + a.y) * (b.w - b.x);
    return r;
}

// quat_conjugate (torch_utils.py:79-82)
PHC_HD Q4 quat_conj(Q4 a) { return Q4{-a.x, -a.y, -a.z, a.w}; }

// my_quat_rotate (torch_utils.py:274-281): v*(2w^2-1) + cross(q,v)*w*2 + q*dot(q,v)*2
PHC_HD V3 quat_rotate(Q4 q, V3 v) {
    float s = 2.0f * (q.w * q.w) - 1.0f;
    float cx = q.y * v.z - q.z * v.y, cy = q.z * v.x - q.x * v.z, cz = q.x * v.y - q.y * v.x;
    float d = (q.x * v.x + q.y * v.y) + q.z * v.z;
    V3 r;
    r.x = (v.x * s + (cx * q.w) * 2.0f) + (q.x * d) * 2.0f;
    r.y = (v.y * s + (cy * q.w) * 2.0f) + (q.y * d) * 2.0f;
    r.z = (v.z * s + (cz * q.w) * 2.0f) + (q.z * d) * 2.0f;
    return r;
}

// my_quat_rotate for a pure z-rotation q = (0, 0, qz, qw) (the heading quaternions): the terms that
// multiply the zero components are dropped; x+0 and 0*x are exact so the result is unchanged.
PHC_HD V3 rotate_z(float qz, float qw, V3 v) {
    float s = 2.0f * (qw * qw) - 1.0f;
    float cx = -(qz * v.y), cy = qz * v.x;
    float d = qz * v.z;
    V3 r;
    r.x = v.x * s + (cx * qw) * 2.0f;
    r.y = v.y * s + (cy * qw) * 2.0f;
    r.z = v.z * s + (qz * d) * 2.0f;
    return r;
}

// quat_to_tan_norm (torch_utils.py:285-297): my_quat_rotate of (1,0,0) then (0,0,1); zero terms dropped.
PHC_HD void tan_norm(Q4 q, float* o) {
    float s = 2.0f * (q.w * q.w) - 1.0f;
    o[0] = s + (q.x * q.x) * 2.0f;
    o[1] = (q.z * q.w) * 2.0f + (q.y * q.x) * 2.0f;
    o[2] = ((-q.y) * q.w) * 2.0f + (q.z * q.x) * 2.0f;
    o[3] = (q.y * q.w) * 2.0f + (q.x * q.z) * 2.0f;
    o[4] = ((-q.x) * q.w) * 2.0f + (q.y * q.z) * 2.0f;
    o[5] = s + (q.z * q.z) * 2.0f;
}

// ---- fused-multiply-add flavours for the fused step's observation math ----------------------------------------
// Same quantities as quat_mul / rotate_z / tan_norm above, written as explicit fmaf chains (about half the
// instructions).  They round differently from the reference's op-by-op sequence (a few ulp), which is fine for
// the fp32 observation values (tolerance 1e-5) and is never used for anything that feeds an index or a flag.
struct ZRot { float A, B, C; };     // rotation about z by the heading quaternion (0,0,qz,qw): A = cos h, B = sin h, C = 1
PHC_HD ZRot zrot_make(float qz, float qw) {
    ZRot r;
    r.A = fmaf(2.0f * qw, qw, -1.0f);
    r.B = (2.0f * qz) * qw;
    r.C = fmaf(2.0f * qz, qz, r.A);
    return r;
}
// my_quat_rotate((0,0,-qz,qw), v): the inverse heading rotation
PHC_HD V3 zrot_inv(const ZRot& h, V3 v) { return V3{fmaf(h.B, v.y, h.A * v.x), fmaf(-h.B, v.x, h.A * v.y), h.C * v.z}; }
// (0,0,z1,w1) * q
PHC_HD Q4 quat_mul_zl(float z1, float w1, Q4 q) {
    return Q4{fmaf(-z1, q.y, w1 * q.x), fmaf(z1, q.x, w1 * q.y), fmaf(z1, q.w, w1 * q.z), fmaf(-z1, q.z, w1 * q.w)};
}
// q * (0,0,z2,w2)
PHC_HD Q4 quat_mul_zr(Q4 q, float z2, float w2) {
    return Q4{fmaf(q.y, z2, q.x * w2), fmaf(-q.x, z2, q.y * w2), fmaf(q.w, z2, q.z * w2), fmaf(-q.z, z2, q.w * w2)};
}
// a * conj(b)
PHC_HD Q4 quat_mul_conj_fma(Q4 a, Q4 b) {
    Q4 r;
    r.x = fmaf(-a.w, b.x, fmaf(a.x, b.w, fmaf(-a.y, b.z, a.z * b.y)));
    r.y = fmaf(-a.w, b.y, fmaf(a.y, b.w, fmaf(-a.z, b.x, a.x * b.z)));
    r.z = fmaf(-a.w, b.z, fmaf(a.z, b.w, fmaf(-a.x, b.y, a.y * b.x)));
    r.w = fmaf(a.w, b.w, fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)));
    return r;
}
PHC_HD void tan_norm_fma(Q4 q, float* o) {
    const float x2 = 2.0f * q.x, y2 = 2.0f * q.y, z2 = 2.0f * q.z;
    const float s = fmaf(2.0f * q.w, q.w, -1.0f);
    o[0] = fmaf(x2, q.x, s);
    o[1] = fmaf(z2, q.w, y2 * q.x);
    o[2] = fmaf(-y2, q.w, z2 * q.x);
    o[3] = fmaf(y2, q.w, x2 * q.z);
    o[4] = fmaf(-x2, q.w, y2 * q.z);
    o[5] = fmaf(z2, q.z, s);
}

// calc_heading (torch_utils.py:369-380): atan2 of the rotated x axis.
PHC_HD float calc_heading(Q4 q) {
    float s = 2.0f * (q.w * q.w) - 1.0f;
    float dx = s + (q.x * q.x) * 2.0f;
    float dy = (q.z * q.w) * 2.0f + (q.y * q.x) * 2.0f;
    return atan2f(dy, dx);
}

// calc_heading_quat (torch_utils.py:384-394) = quat_from_angle_axis(heading, z) (:354-358):
// (0, 0, sin(h/2), cos(h/2)) divided by its norm (quat_unit :174-179; torch's CPU norm
// accumulates squares with an fma chain).  calc_heading_quat_inv (:398-408) uses -heading,
// i.e. exactly the conjugate, because sin is odd and cos even in every libm used here.
PHC_HD void heading_quat(float heading, float& qz, float& qw) {
    float th = heading / 2.0f;
    float sn = sinf(th), cs = cosf(th);
    float n = sqrtf(fmaf(cs, cs, sn * sn));
    if (n < 1e-9f) n = 1e-9f;
    qz = sn / n;
    qw = cs / n;
}

// Heading quaternion without the atan2 -> sin/cos round trip: with (dx, dy) the rotated x axis projected on the
// ground plane, cos h = dx/r and sin h = dy/r, and the half-angle identities give (sin(h/2), cos(h/2)) directly
// (the branch keeps the square root away from cancellation).  Equal to heading_quat(calc_heading(q)) to a few ulp;
// used by the fused step, whose heading only feeds fp32 observation values (tolerance 1e-5), never a flag.
PHC_HD void heading_quat_direct(Q4 q, float& qz, float& qw) {
    float s = 2.0f * (q.w * q.w) - 1.0f;
    float dx = s + (q.x * q.x) * 2.0f;
    float dy = (q.z * q.w) * 2.0f + (q.y * q.x) * 2.0f;
    float r = sqrtf(dx * dx + dy * dy);
    if (!(r > 0.0f)) { qz = 0.0f; qw = 1.0f; return; }      // atan2(0, 0) = 0 in the reference
    float inv = 1.0f / r;
    float ch = dx * inv, sh = dy * inv;
    if (ch >= 0.0f) {
        qw = sqrtf(0.5f * (1.0f + ch));
        qz = sh / (2.0f * qw);
    } else {
        float z = sqrtf(0.5f * (1.0f - ch));
        qz = sh < 0.0f ? -z : z;
        qw = sh / (2.0f * qz);
    }
}

// sin(x) for x in [0, pi/2] (the only range slerp needs: x = t * acos(c), c in [0,1), t in [0,1]):
// odd polynomial through x^13 in Horner form, relative error < 1e-7 (about 1 ulp), no range reduction.
PHC_HD float sin_0_halfpi(float x) {
    const float x2 = x * x;
    float p = 1.6059043836821613e-10f;                  //  1/13!
    p = fmaf(p, x2, -2.5052108385441720e-08f);          // -1/11!
    p = fmaf(p, x2, 2.7557319223985893e-06f);           //  1/9!
    p = fmaf(p, x2, -1.9841269841269841e-04f);          // -1/7!
    p = fmaf(p, x2, 8.3333333333333332e-03f);           //  1/5!
    p = fmaf(p, x2, -1.6666666666666666e-01f);          // -1/3!
    return fmaf(x * x2, p, x);
}

// remove_base_rot (reference puffer_phc/envs/common.py:15-19), used when upright == false.
PHC_HD Q4 remove_base_rot(Q4 q) { return quat_mul(q, Q4{-0.5f, -0.5f, -0.5f, 0.5f}); }

// quat_to_angle_axis (torch_utils.py:86-106) -> angle only (normalize_angle :50-51 applied).
PHC_HD float quat_angle(Q4 q) {
    float s = sqrtf(1.0f - q.w * q.w);
    if (!(fabsf(s) > 1e-5f)) return 0.0f;      // NaN -> masked, like torch.where on a false mask
    float a = 2.0f * acosf(q.w);
    return atan2f(sinf(a), cosf(a));
}

// Squared rotation angle of q for the rotation reward (common.py:304-306).  normalize_angle(2 acos w) is
// 2 acos w wrapped into (-pi, pi]; only its square is used, so the wrap is done in closed form instead of the
// reference's atan2(sin, cos) round trip (equal to a few ulp; the round trip itself carries ~5e-7 of noise).
PHC_HD float quat_angle_sq(Q4 q) {
    // the reference masks on |sqrt(1 - w^2)| > 1e-5; in fp32 1 - w^2 is either <= 0 / NaN or >= 2^-24, so the test is "> 0" and the
    // square root (only ever used for this decision) is not needed: identical mask, shorter dependent chain
    const float x = 1.0f - q.w * q.w;
    if (!(x > 1e-10f)) return 0.0f;
    float a = 2.0f * acosf(q.w);
    if (a > 3.14159265358979323846f) a = a - 6.28318530717958647692f;
    return a * a;
}

// quat_to_exp_map (torch_utils.py:144-150)
PHC_HD V3 quat_exp_map(Q4 q) {
    float s = sqrtf(1.0f - q.w * q.w);
    if (!(fabsf(s) > 1e-5f)) return V3{0.0f, 0.0f, 0.0f};     // angle 0 times axis (0,0,1)
    float a = 2.0f * acosf(q.w);
    a = atan2f(sinf(a), cosf(a));
    return V3{a * (q.x / s), a * (q.y / s), a * (q.z / s)};
}

// quat_to_exp_map with the angle wrapped in closed form: normalize_angle(2 acos w) is 2 acos w - 2 pi when it exceeds pi
// (the reference's atan2(sin a, cos a) round trip gives the same value to a few ulp; they only differ in sign for an exact
// 180-degree rotation, w == 0, a measure-zero knife edge).
PHC_HD V3 quat_exp_map_fast(Q4 q) {
    float s = sqrtf(1.0f - q.w * q.w);
    if (!(fabsf(s) > 1e-5f)) return V3{0.0f, 0.0f, 0.0f};
    float a = 2.0f * acosf(q.w);
    if (a > 3.14159265358979323846f) a = a - 6.28318530717958647692f;
    const float k = a / s;
    return V3{k * q.x, k * q.y, k * q.z};
}

// slerp (torch_utils.py:110-131).  The two torch.where fall-backs are evaluated first (they
// discard the trigonometric result anyway): q0 when |cos| >= 1, the un-normalised midpoint when
// |sin| < 1e-3.  No renormalisation.
PHC_HD Q4 slerp(Q4 q0, Q4 q1, float t, int dev = PHC_REF_CPU) {
    float c = dot4(q0, q1, dev);
    if (c < 0.0f) { q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z; q1.w = -q1.w; }
    c = fabsf(c);
    if (c >= 1.0f) return q0;
    float s = sqrtf(1.0f - c * c);
    if (fabsf(s) < 0.001f)
        return Q4{0.5f * q0.x + 0.5f * q1.x, 0.5f * q0.y + 0.5f * q1.y, 0.5f * q0.z + 0.5f * q1.z, 0.5f * q0.w + 0.5f * q1.w};
    float h = acosf(c);
    float ra = sinf((1.0f - t) * h) / s;
    float rb = sinf(t * h) / s;
    return Q4{ra * q0.x + rb * q1.x, ra * q0.y + rb * q1.y, ra * q0.z + rb * q1.z, ra * q0.w + rb * q1.w};
}

// Same slerp for the fused step: identical branch decisions (they depend on c and s only), but one IEEE
// reciprocal shared by the two ratios instead of two divisions, and a bounded-range polynomial sine
// (the arguments lie in [0, pi/2]); a few ulp from slerp().
PHC_HD Q4 slerp_rcp(Q4 q0, Q4 q1, float t, int dev = PHC_REF_CPU) {
    float c = dot4(q0, q1, dev);
    if (c < 0.0f) { q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z; q1.w = -q1.w; }
    c = fabsf(c);
    if (c >= 1.0f) return q0;
    float s = sqrtf(1.0f - c * c);
    if (fabsf(s) < 0.001f)
        return Q4{0.5f * q0.x + 0.5f * q1.x, 0.5f * q0.y + 0.5f * q1.y, 0.5f * q0.z + 0.5f * q1.z, 0.5f * q0.w + 0.5f * q1.w};
    // NOTE: no shortcut for t == 0 or 1.  The reference's ratio sin(h)/sqrt(1-c*c) is NOT 1 there: 1-c*c cancels, so the
    // ratio is off by up to ~1.5e-4 for nearby frames, and parity means reproducing that value (same c, same 1-c*c).
    float h = acosf(c);
    float inv = 1.0f / s;
    float ra = sin_0_halfpi((1.0f - t) * h) * inv;
    float rb = sin_0_halfpi(t * h) * inv;
    return Q4{ra * q0.x + rb * q1.x, ra * q0.y + rb * q1.y, ra * q0.z + rb * q1.z, ra * q0.w + rb * q1.w};
}

// ---- slerp with the pair quantities precomputed (motion library "pair aux" table, csrc/motion_state.cu) ---------------------------
// Everything in slerp that depends only on the two table rotations -- the dot product c, the sign flip, the two fall-back decisions,
// h = acos(c) and 1/sqrt(1-c*c) -- is a property of the frame pair (f, f+1), not of the query.  slerp_pair_make() computes it ONCE per
// pair with exactly the operations slerp_rcp() uses (so the results are bit-identical), the fused step reads it back as two floats:
//     h >= 0 : regular slerp, inv = +-1/sin_half (sign = the reference's q1 = -q1 flip)
//     h = -1 : |cos| >= 1             -> q0
//     h = -2 : |sin_half| < 0.001     -> 0.5*q0 + 0.5*(+-q1), inv = +-1
struct SlerpPair { float h, inv; };
PHC_HD SlerpPair slerp_pair_make(Q4 q0, Q4 q1, int dev) {
    float c = dot4(q0, q1, dev);
    const float sgn = c < 0.0f ? -1.0f : 1.0f;
    c = fabsf(c);
    if (c >= 1.0f) return SlerpPair{-1.0f, 1.0f};
    const float s = sqrtf(1.0f - c * c);
    if (fabsf(s) < 0.001f) return SlerpPair{-2.0f, sgn};
    return SlerpPair{acosf(c), sgn * (1.0f / s)};
}
PHC_HD Q4 slerp_pair(Q4 q0, Q4 q1, float t, SlerpPair a) {
    if (a.h == -1.0f) return q0;
    if (a.h == -2.0f) {
        const float hb = 0.5f * a.inv;
        return Q4{0.5f * q0.x + hb * q1.x, 0.5f * q0.y + hb * q1.y, 0.5f * q0.z + hb * q1.z, 0.5f * q0.w + hb * q1.w};
    }
    const float ia = fabsf(a.inv);
    const float ra = sin_0_halfpi((1.0f - t) * a.h) * ia;
    const float rb = sin_0_halfpi(t * a.h) * a.inv;
    return Q4{ra * q0.x + rb * q1.x, ra * q0.y + rb * q1.y, ra * q0.z + rb * q1.z, ra * q0.w + rb * q1.w};
}
// t == 0 and not the midpoint fall-back (h == -2, the caller's business): the second rotation contributes rb * q1 = +-0 only
PHC_HD Q4 slerp_pair_t0(Q4 q0, SlerpPair a) {
    if (a.h < 0.0f) return q0;
    const float ra = sin_0_halfpi(a.h) * fabsf(a.inv);
    return Q4{ra * q0.x, ra * q0.y, ra * q0.z, ra * q0.w};
}

// exp_map_to_quat (torch_utils.py:333-365) = quat_from_angle_axis(exp_map_to_angle_axis(e)); norms use the CPU reference's
// fma chain, the small-angle mask falls back to angle 0 about (0,0,1).
PHC_HD Q4 exp_map_to_quat(V3 e, int dev = PHC_REF_CPU) {
    float angle = norm3(e, dev);
    V3 axis{e.x / angle, e.y / angle, e.z / angle};
    angle = atan2f(sinf(angle), cosf(angle));                       // normalize_angle (:50-51)
    if (!(fabsf(angle) > 1e-5f)) { angle = 0.0f; axis = V3{0.0f, 0.0f, 1.0f}; }
    const float th = angle / 2.0f;
    float an = norm3(axis, dev);
    if (an < 1e-9f) an = 1e-9f;
    const float sn = sinf(th);
    const float x = (axis.x / an) * sn, y = (axis.y / an) * sn, z = (axis.z / an) * sn, w = cosf(th);
    float n = sqrtf(fmaf(w, w, fmaf(z, z, fmaf(y, y, x * x))));
    if (n < 1e-9f) n = 1e-9f;
    return Q4{x / n, y / n, z / n, w / n};
}

// lerp as written in get_motion_state (motion_lib.py:596-603): (1-b)*x0 + b*x1
PHC_HD float lerp(float a, float b, float one_m, float t) { return one_m * a + t * b; }

// .mean(dim=-1) over n <= 32 contiguous values in torch's own summation order (both fitted bit-exactly, n = 20 and 24):
//   CUDA: the reduce kernel runs bx = last_pow2(n) lanes, lane k first adds v[k + bx] (when it exists), then a shuffle tree (offsets
//         bx/2 .. 1), and the sum is MULTIPLIED by float(1/n);
//   CPU:  eight vector lanes a_k = v[k] + v[k+8] + ... over the full rows of 8, the scalar tail (n % 8 elements) summed first, then
//         the lanes added one by one, and the sum is DIVIDED by n.
PHC_HD float mean_ordered(const float* v, int n, int dev) {
    if (dev == PHC_REF_CUDA) {
        int bx = 1;
        while (bx * 2 <= n) bx *= 2;
        float s[32];
        for (int k = 0; k < bx; ++k) s[k] = (k + bx < n) ? v[k] + v[k + bx] : v[k];
        for (int off = bx / 2; off >= 1; off /= 2)
            for (int k = 0; k < off; ++k) s[k] = s[k] + s[k + off];
        return s[0] * (1.0f / (float)n);
    }
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

// mean over xyz of squares: ((x^2 + y^2) + z^2) / 3  ((diff**2).mean(dim=-1), common.py:300)
PHC_HD float mean_sq3(V3 d) { return ((d.x * d.x + d.y * d.y) + d.z * d.z) / 3.0f; }
// sum of squares only; the /3 and /J of the two means are applied once per env (reward_from_sq_sums)
PHC_HD float sum_sq3(V3 d) { return (d.x * d.x + d.y * d.y) + d.z * d.z; }

}  // namespace phc
