// phc_body.cuh -- per-body pieces of the observation / reward math of reference puffer_phc/envs/common.py,
// shared by the stand-alone kernels (imitation.cu) and the fused step (step_fused.cu).  The heading
// quaternion of the simulated root is h = (0, 0, hz, hw) and h^-1 = (0, 0, -hz, hw) (see heading_quat).
// Host-compilable (tests/host_math_harness.cpp).
#pragma once
#include "phc_common.cuh"

namespace phc {

struct BodyState { V3 p; Q4 q; V3 v; V3 w; };   // pos, rot (xyzw), lin vel, ang vel of one body

PHC_HD void put3(float* o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

// get_motion_state blend of one body (motion_lib.py:596-610); offset is added to the position only.
PHC_HD BodyState blend_frames(const BodyState& a, const BodyState& b, float blend, V3 off, int dev = PHC_REF_CPU) {
    const float om = 1.0f - blend;
    BodyState r;
    r.p = V3{lerp(a.p.x, b.p.x, om, blend) + off.x, lerp(a.p.y, b.p.y, om, blend) + off.y, lerp(a.p.z, b.p.z, om, blend) + off.z};
    r.q = slerp_rcp(a.q, b.q, blend, dev);
    // velocities feed fp32 outputs only (no index, no flag): one fused multiply-add per component, <= 1 ulp from the reference's
    // (1-b)*x0 + b*x1; the position above keeps the reference's three roundings because the termination test reads it
    r.v = V3{fmaf(blend, b.v.x, om * a.v.x), fmaf(blend, b.v.y, om * a.v.y), fmaf(blend, b.v.z, om * a.v.z)};
    r.w = V3{fmaf(blend, b.w.x, om * a.w.x), fmaf(blend, b.w.y, om * a.w.y), fmaf(blend, b.w.z, om * a.w.z)};
    return r;
}


// The same blend with the pair's slerp quantities read from the motion library's pair-aux table (bit-identical to blend_frames).
PHC_HD BodyState blend_frames_pair(const BodyState& a, const BodyState& b, float blend, V3 off, SlerpPair sp) {
    const float om = 1.0f - blend;
    BodyState r;
    r.p = V3{lerp(a.p.x, b.p.x, om, blend) + off.x, lerp(a.p.y, b.p.y, om, blend) + off.y, lerp(a.p.z, b.p.z, om, blend) + off.z};
    r.q = slerp_pair(a.q, b.q, blend, sp);
    r.v = V3{fmaf(blend, b.v.x, om * a.v.x), fmaf(blend, b.v.y, om * a.v.y), fmaf(blend, b.v.z, om * a.v.z)};
    r.w = V3{fmaf(blend, b.w.x, om * a.w.x), fmaf(blend, b.w.y, om * a.w.y), fmaf(blend, b.w.z, om * a.w.z)};
    return r;
}
// blend == 0 exactly (the query time sits on a table frame: ~84 % of the queries when control and motion run at the same rate) and no
// body of the pair takes the midpoint fall-back: frame 1 only ever contributes 0 * x = +-0, so it is neither fetched nor read.
// Equal to blend_frames(a, b, 0, off) up to the sign of an exact zero.
PHC_HD BodyState blend_frames_t0(const BodyState& a, V3 off, SlerpPair sp) {
    BodyState r;
    r.p = V3{a.p.x + off.x, a.p.y + off.y, a.p.z + off.z};
    r.q = slerp_pair_t0(a.q, sp);
    r.v = a.v;
    r.w = a.w;
    return r;
}

// compute_imitation_observations_v6 (common.py:137-173) and compute_humanoid_observations_smpl_max (common.py:57-89) for one body;
// b = simulated body, r = reference body, h = heading rotation of the simulated root.  Explicit fmaf chains (see phc_math.cuh): a few ulp
// from the reference's op order.  Every kernel that produces observation rows (stand-alone, fused step, auto-reset tail) uses these,
// and tests/host_math_harness.cpp replays them on the host against the reference's golden vectors.
PHC_HD void task_obs_body_fma(const BodyState& b, const BodyState& r, V3 root_pos, float hz, float hw, const ZRot& h, float* o_dpos,
                              float* o_drot, float* o_dvel, float* o_dang, float* o_lpos, float* o_lrot) {
    put3(o_dpos, zrot_inv(h, r.p - b.p));                                                 // common.py:138-139
    const Q4 dq = quat_mul_conj_fma(r.q, b.q);                                            // :142-145
    tan_norm_fma(quat_mul_zr(quat_mul_zl(-hz, hw, dq), hz, hw), o_drot);                  // :146-149, :169
    put3(o_dvel, zrot_inv(h, r.v - b.v));                                                 // :152-153
    put3(o_dang, zrot_inv(h, r.w - b.w));                                                 // :155-156
    put3(o_lpos, zrot_inv(h, r.p - root_pos));                                            // :159-162
    tan_norm_fma(quat_mul_zl(-hz, hw, r.q), o_lrot);                                      // :164-165
}
PHC_HD void self_obs_pos_rot_fma(const BodyState& b, V3 root_pos, float hz, float hw, const ZRot& h, int j, float* o_pos, float* o_rot) {
    if (j >= 1) put3(o_pos, zrot_inv(h, b.p - root_pos));                                 // common.py:57-66
    tan_norm_fma(quat_mul_zl(-hz, hw, b.q), o_rot);                                       // :68-75
}
PHC_HD void self_obs_vel_ang_fma(const BodyState& b, const ZRot& h, float* o_vel, float* o_ang) {
    put3(o_vel, zrot_inv(h, b.v));                                                        // common.py:81-83
    put3(o_ang, zrot_inv(h, b.w));                                                        // :85-89
}
PHC_HD void reward_terms_body_fma(const BodyState& b, const BodyState& r, float& sp, float& sr, float& sv, float& sa) {
    const V3 dp = r.p - b.p, dv = r.v - b.v, da = r.w - b.w;
    sp = fmaf(dp.z, dp.z, fmaf(dp.y, dp.y, dp.x * dp.x));
    sr = quat_angle_sq(quat_mul_conj_fma(r.q, b.q));
    sv = fmaf(dv.z, dv.z, fmaf(dv.y, dv.y, dv.x * dv.x));
    sa = fmaf(da.z, da.z, fmaf(da.y, da.y, da.x * da.x));
}

// env-level tail (common.py:300-320) for sums of squares: sp/sv/sa are sums over J bodies x 3 components, sr over J bodies.
PHC_HD float reward_from_sq_sums(float sp, float sr, float sv, float sa, float J, const float* k, const float* w, float* raw) {
    const float inv3j = 1.0f / (3.0f * J), invj = 1.0f / J;
    raw[0] = expf(-k[0] * (sp * inv3j));
    raw[1] = expf(-k[1] * (sr * invj));
    raw[2] = expf(-k[2] * (sv * inv3j));
    raw[3] = expf(-k[3] * (sa * inv3j));
    return ((w[0] * raw[0] + w[1] * raw[1]) + w[2] * raw[2]) + w[3] * raw[3];
}

}  // namespace phc
