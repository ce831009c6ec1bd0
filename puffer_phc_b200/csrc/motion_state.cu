// motion_state.cu -- MotionLibBase.get_motion_state / get_root_pos_smpl / sample_time_interval
// (reference puffer_phc/motion_lib.py:526-535, 549-665) and the packed-frame table builder.
//
// One warp per query, lane j = body j: frame-index/blend (bit-exact op order), the row gathers of the
// two bracketing frames, lerp of pos/vel/ang-vel/dof-vel (bit-exact: pure mul/add), slerp of global and local
// rotations and the quaternion -> exp-map of the local rotations, all in one pass; nothing is materialised.
// The rotations use the cheaper, algebraically identical forms of phc_math.cuh (same branch decisions, a few ulp
// from the reference's op sequence): the kernel is instruction-bound with the libm sin/cos/atan2 chain.
#include "phc_common.cuh"

namespace phc {

struct StateArgs {
    phc_motion_tables t;
    const int64_t* ids;
    const float* times;
    const float* offset;
    int64_t B;
    phc_motion_state_out o;
    int dev;
};

constexpr int MS_WARPS = 4;

__global__ void __launch_bounds__(MS_WARPS * 32) motion_state_kernel(const StateArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * MS_WARPS + (threadIdx.x >> 5);
    if (q >= a.B) return;
    const phc_motion_tables& T = a.t;
    const phc_motion_state_out& o = a.o;

    const int64_t id = __ldg(a.ids + q);
    const float time = __ldg(a.times + q);
    int64_t i0, i1;
    float blend;
    frame_blend(time, __ldg(T.motion_len + id), __ldg(T.num_frames + id), __ldg(T.motion_dt + id), i0, i1, blend);
    const int64_t ls = __ldg(T.length_starts + id);
    const int64_t f0 = i0 + ls, f1 = i1 + ls;
    const float one_m = 1.0f - blend;
    if (lane == 0) {
        if (o.frame_idx0) o.frame_idx0[q] = i0;
        if (o.frame_idx1) o.frame_idx1[q] = i1;
        if (o.blend) o.blend[q] = blend;
    }

    if (lane < NB) {
        const int j = lane;
        if (o.rg_pos || o.root_pos) {
            V3 p0 = ldg3(T.gts + (f0 * NB + j) * 3), p1 = ldg3(T.gts + (f1 * NB + j) * 3);
            V3 p{lerp(p0.x, p1.x, one_m, blend), lerp(p0.y, p1.y, one_m, blend), lerp(p0.z, p1.z, one_m, blend)};
            if (a.offset) {   // motion_lib.py:599: (lerp) + offset
                p.x = p.x + __ldg(a.offset + q * 3 + 0);
                p.y = p.y + __ldg(a.offset + q * 3 + 1);
                p.z = p.z + __ldg(a.offset + q * 3 + 2);
            }
            if (o.rg_pos) st3(o.rg_pos + (q * NB + j) * 3, p);
            if (o.root_pos && j == 0) st3(o.root_pos + q * 3, p);
        }
        if (o.body_vel || o.root_vel) {
            V3 p0 = ldg3(T.gvs + (f0 * NB + j) * 3), p1 = ldg3(T.gvs + (f1 * NB + j) * 3);
            V3 p{lerp(p0.x, p1.x, one_m, blend), lerp(p0.y, p1.y, one_m, blend), lerp(p0.z, p1.z, one_m, blend)};
            if (o.body_vel) st3(o.body_vel + (q * NB + j) * 3, p);
            if (o.root_vel && j == 0) st3(o.root_vel + q * 3, p);
        }
        if (o.body_ang_vel || o.root_ang_vel) {
            V3 p0 = ldg3(T.gavs + (f0 * NB + j) * 3), p1 = ldg3(T.gavs + (f1 * NB + j) * 3);
            V3 p{lerp(p0.x, p1.x, one_m, blend), lerp(p0.y, p1.y, one_m, blend), lerp(p0.z, p1.z, one_m, blend)};
            if (o.body_ang_vel) st3(o.body_ang_vel + (q * NB + j) * 3, p);
            if (o.root_ang_vel && j == 0) st3(o.root_ang_vel + q * 3, p);
        }
        if (o.rb_rot || o.root_rot) {
            Q4 r = slerp_rcp(ldg4a(T.grs + (f0 * NB + j) * 4), ldg4a(T.grs + (f1 * NB + j) * 4), blend, a.dev);
            if (o.rb_rot) *reinterpret_cast<float4*>(o.rb_rot + (q * NB + j) * 4) = make_float4(r.x, r.y, r.z, r.w);
            if (o.root_rot && j == 0) st4(o.root_rot + q * 4, r);
        }
        if (j >= 1) {
            if (o.dof_pos) {   // motion_lib.py:605-606, 670-673
                Q4 r = slerp_rcp(ldg4a(T.lrs + (f0 * NB + j) * 4), ldg4a(T.lrs + (f1 * NB + j) * 4), blend, a.dev);
                st3(o.dof_pos + q * NDOF + (j - 1) * 3, quat_exp_map_fast(r));
            }
            if (o.dof_vel) {
                V3 p0 = ldg3(T.dvs + (f0 * 23 + (j - 1)) * 3), p1 = ldg3(T.dvs + (f1 * 23 + (j - 1)) * 3);
                V3 p{lerp(p0.x, p1.x, one_m, blend), lerp(p0.y, p1.y, one_m, blend), lerp(p0.z, p1.z, one_m, blend)};
                st3(o.dof_vel + q * NDOF + (j - 1) * 3, p);
            }
        }
    }
    if (o.motion_aa)   // motion_lib.py:619: frame f0 only, not blended
        for (int c = lane; c < 72; c += 32) o.motion_aa[q * 72 + c] = __ldg(T.motion_aa + f0 * 72 + c);
    if (o.motion_bodies && lane < 17) o.motion_bodies[q * 17 + lane] = __ldg(T.motion_bodies + id * 17 + lane);
    if (o.motion_limb_weights && lane < 10) o.motion_limb_weights[q * 10 + lane] = __ldg(T.limb_weights + id * 10 + lane);
}

// Reset path: the same query, written straight into the env's state tensors (humanoid_phc.py:843-873, 899-929).
struct ResetArgs2 {
    phc_motion_tables t;
    const int64_t* env_ids; const int64_t* motion_ids; const float* times; const float* offset; int64_t K;
    float* root_states; float* dof_pos; float* dof_vel; float* body_state; int64_t env_stride; int dev;
};

__global__ void __launch_bounds__(MS_WARPS * 32) reset_ref_state_kernel(const ResetArgs2 a) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * MS_WARPS + (threadIdx.x >> 5);
    if (i >= a.K) return;
    const phc_motion_tables& T = a.t;
    const int64_t e = __ldg(a.env_ids + i);
    const int64_t id = __ldg(a.motion_ids + e);
    int64_t i0, i1;
    float blend;
    frame_blend(__ldg(a.times + i), __ldg(T.motion_len + id), __ldg(T.num_frames + id), __ldg(T.motion_dt + id), i0, i1, blend);
    const int64_t ls = __ldg(T.length_starts + id);
    const int64_t f0 = i0 + ls, f1 = i1 + ls;
    const float one_m = 1.0f - blend;
    if (lane >= NB) return;
    const int j = lane;
    V3 off{0.0f, 0.0f, 0.0f};
    if (a.offset) off = ldg3(a.offset + e * 3);
    const bool need_body = a.body_state || (a.root_states && j == 0);
    if (need_body) {
        V3 p0 = ldg3(T.gts + (f0 * NB + j) * 3), p1 = ldg3(T.gts + (f1 * NB + j) * 3);
        V3 p{lerp(p0.x, p1.x, one_m, blend), lerp(p0.y, p1.y, one_m, blend), lerp(p0.z, p1.z, one_m, blend)};
        if (a.offset) { p.x = p.x + off.x; p.y = p.y + off.y; p.z = p.z + off.z; }
        const Q4 r = slerp_rcp(ldg4a(T.grs + (f0 * NB + j) * 4), ldg4a(T.grs + (f1 * NB + j) * 4), blend, a.dev);
        V3 v0 = ldg3(T.gvs + (f0 * NB + j) * 3), v1 = ldg3(T.gvs + (f1 * NB + j) * 3);
        const V3 v{lerp(v0.x, v1.x, one_m, blend), lerp(v0.y, v1.y, one_m, blend), lerp(v0.z, v1.z, one_m, blend)};
        V3 w0 = ldg3(T.gavs + (f0 * NB + j) * 3), w1 = ldg3(T.gavs + (f1 * NB + j) * 3);
        const V3 w{lerp(w0.x, w1.x, one_m, blend), lerp(w0.y, w1.y, one_m, blend), lerp(w0.z, w1.z, one_m, blend)};
        if (a.body_state) {
            float* b = a.body_state + e * a.env_stride + REC * j;
            st3(b, p); st4(b + 3, r); st3(b + 7, v); st3(b + 10, w);
        }
        if (a.root_states && j == 0) {
            float* b = a.root_states + e * REC;
            st3(b, p); st4(b + 3, r); st3(b + 7, v); st3(b + 10, w);
        }
    }
    if (j >= 1) {
        if (a.dof_pos) {
            const Q4 r = slerp_rcp(ldg4a(T.lrs + (f0 * NB + j) * 4), ldg4a(T.lrs + (f1 * NB + j) * 4), blend, a.dev);
            st3(a.dof_pos + e * NDOF + (j - 1) * 3, quat_exp_map_fast(r));
        }
        if (a.dof_vel) {
            V3 p0 = ldg3(T.dvs + (f0 * 23 + (j - 1)) * 3), p1 = ldg3(T.dvs + (f1 * 23 + (j - 1)) * 3);
            st3(a.dof_vel + e * NDOF + (j - 1) * 3, V3{lerp(p0.x, p1.x, one_m, blend), lerp(p0.y, p1.y, one_m, blend), lerp(p0.z, p1.z, one_m, blend)});
        }
    }
}

__global__ void sample_time_interval_kernel(const float* __restrict__ phase, const float* __restrict__ len, int64_t n,
                                            int div_mode, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fps = (float)(1.0 / 30.0);     // curr_fps = 1/30 as a Python double, cast at the op (motion_lib.py:532)
    const float x = phase[i] * len[i];
    // torch-CUDA divides a tensor by a Python scalar as a multiplication by float(1.0 / scalar) with the reciprocal formed in DOUBLE:
    // 1.0 / (1/30) = 30.000000000000004 -> 30.0f (bit-exact against torch-CUDA on 262144 samples; 1.0f / fps would be 29.999998f)
    const float qv = div_mode ? x * (float)(1.0 / (1.0 / 30.0)) : x / fps;
    out[i] = (float)(int64_t)qv * fps;
}

// packed[f] = gts[f] (72) | grs[f] (96) | gvs[f] (72) | gavs[f] (72)
__global__ void pack_frames_kernel(const phc_motion_tables T, float* __restrict__ packed) {
    const int64_t total = T.F * FRAME_F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / FRAME_F;
        const int c = (int)(i - f * FRAME_F);
        float v;
        if (c < 72) v = __ldg(T.gts + f * 72 + c);
        else if (c < 168) v = __ldg(T.grs + f * 96 + (c - 72));
        else if (c < 240) v = __ldg(T.gvs + f * 72 + (c - 168));
        else v = __ldg(T.gavs + f * 72 + (c - 240));
        packed[i] = v;
    }
}

// pair_aux[f][j] = slerp_pair_make(grs[f][j], grs[f1][j]) with f1 = min(f + 1, last frame of f's clip); pair_flags[f] bit 0 = some
// body takes the midpoint fall-back.  One warp per frame, lane = body; the clip of a frame by binary search over length_starts.
__global__ void __launch_bounds__(128) pair_aux_kernel(const phc_motion_tables T, int dev, float* __restrict__ aux, uint8_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= T.F) return;
    int64_t lo = 0, hi = T.M;                     // largest clip with length_starts <= f
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(T.length_starts + mid) <= f) lo = mid; else hi = mid;
    }
    const int64_t last = __ldg(T.length_starts + lo) + __ldg(T.num_frames + lo) - 1;
    const int64_t f1 = f + 1 <= last ? f + 1 : last;
    bool mid2 = false;
    if (lane < NB) {
        const SlerpPair sp = slerp_pair_make(ldg4a(T.grs + (f * NB + lane) * 4), ldg4a(T.grs + (f1 * NB + lane) * 4), dev);
        *reinterpret_cast<float2*>(aux + (f * NB + lane) * 2) = make_float2(sp.h, sp.inv);
        mid2 = sp.h == -2.0f;
    }
    const bool any = __any_sync(FULL, mid2);
    if (lane == 0) flags[f] = any ? 1 : 0;
}

}  // namespace phc

using namespace phc;

extern "C" int phc_build_pair_aux(const phc_motion_tables* t, int ref_device, float* pair_aux, uint8_t* pair_flags, phc_stream_t stream) {
    const char* fn = "phc_build_pair_aux";
    PHC_REQUIRE(t && pair_aux && pair_flags, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(t->grs && t->num_frames && t->length_starts, PHC_EINVAL, "%s: grs / num_frames / length_starts required", fn);
    PHC_REQUIRE(aligned16(t->grs) && aligned8(pair_aux), PHC_EALIGN, "%s: grs must be 16-byte, pair_aux 8-byte aligned", fn);
    PHC_REQUIRE(ref_device == PHC_REF_DEVICE_CPU || ref_device == PHC_REF_DEVICE_CUDA, PHC_EINVAL, "%s: ref_device must be 0 or 1", fn);
    if (t->F <= 0 || t->M <= 0) return PHC_OK;
    pair_aux_kernel<<<(unsigned)((t->F + 3) / 4), 128, 0, (cudaStream_t)stream>>>(*t, ref_device, pair_aux, pair_flags);
    return check_launch(fn);
}

extern "C" int phc_motion_state(const phc_motion_tables* t, const int64_t* motion_ids, const float* motion_times,
                                const float* offset, int64_t B, const phc_motion_state_out* out, int ref_device, phc_stream_t stream) {
    PHC_REQUIRE(t && out, PHC_EINVAL, "phc_motion_state: tables/out is NULL");
    PHC_REQUIRE(B >= 0, PHC_EINVAL, "phc_motion_state: B=%lld < 0", (long long)B);
    if (B == 0) return PHC_OK;
    PHC_REQUIRE(motion_ids && motion_times, PHC_EINVAL, "phc_motion_state: motion_ids/motion_times is NULL");
    PHC_REQUIRE(t->motion_len && t->motion_dt && t->num_frames && t->length_starts, PHC_EINVAL,
                "phc_motion_state: per-motion tables missing");
    const phc_motion_state_out& o = *out;
    PHC_REQUIRE(!(o.rg_pos || o.root_pos) || t->gts, PHC_EINVAL, "phc_motion_state: gts table missing");
    PHC_REQUIRE(!(o.rb_rot || o.root_rot) || t->grs, PHC_EINVAL, "phc_motion_state: grs table missing");
    PHC_REQUIRE(!o.dof_pos || t->lrs, PHC_EINVAL, "phc_motion_state: lrs table missing");
    PHC_REQUIRE(!(o.body_vel || o.root_vel) || t->gvs, PHC_EINVAL, "phc_motion_state: gvs table missing");
    PHC_REQUIRE(!(o.body_ang_vel || o.root_ang_vel) || t->gavs, PHC_EINVAL, "phc_motion_state: gavs table missing");
    PHC_REQUIRE(!o.dof_vel || t->dvs, PHC_EINVAL, "phc_motion_state: dvs table missing");
    PHC_REQUIRE(!o.motion_aa || t->motion_aa, PHC_EINVAL, "phc_motion_state: motion_aa table missing");
    PHC_REQUIRE(!o.motion_bodies || t->motion_bodies, PHC_EINVAL, "phc_motion_state: motion_bodies table missing");
    PHC_REQUIRE(!o.motion_limb_weights || t->limb_weights, PHC_EINVAL, "phc_motion_state: limb_weights table missing");
    PHC_REQUIRE(aligned16(t->grs) && aligned16(t->lrs) && aligned16(o.rb_rot), PHC_EALIGN,
                "phc_motion_state: grs/lrs tables and rb_rot output must be 16-byte aligned");
    PHC_REQUIRE(ref_device == PHC_REF_DEVICE_CPU || ref_device == PHC_REF_DEVICE_CUDA, PHC_EINVAL, "phc_motion_state: ref_device must be 0 or 1");
    StateArgs a{*t, motion_ids, motion_times, offset, B, o, ref_device};
    const int64_t blocks = (B + MS_WARPS - 1) / MS_WARPS;
    motion_state_kernel<<<(unsigned)blocks, MS_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch("phc_motion_state");
}

extern "C" int phc_reset_ref_state(const phc_motion_tables* t, const int64_t* env_ids, const int64_t* sampled_motion_ids,
                                   const float* motion_times, const float* global_offset, int64_t K, float* root_states,
                                   float* dof_pos, float* dof_vel, float* body_state, int64_t env_stride, int ref_device,
                                   phc_stream_t stream) {
    const char* fn = "phc_reset_ref_state";
    PHC_REQUIRE(t, PHC_EINVAL, "%s: tables is NULL", fn);
    PHC_REQUIRE(K >= 0, PHC_EINVAL, "%s: K < 0", fn);
    if (K == 0) return PHC_OK;
    PHC_REQUIRE(env_ids && sampled_motion_ids && motion_times, PHC_EINVAL, "%s: env_ids / sampled_motion_ids / motion_times is NULL", fn);
    PHC_REQUIRE(t->motion_len && t->motion_dt && t->num_frames && t->length_starts, PHC_EINVAL, "%s: per-motion tables missing", fn);
    PHC_REQUIRE(!(body_state || root_states) || (t->gts && t->grs && t->gvs && t->gavs), PHC_EINVAL, "%s: gts/grs/gvs/gavs tables missing", fn);
    PHC_REQUIRE(!dof_pos || t->lrs, PHC_EINVAL, "%s: lrs table missing", fn);
    PHC_REQUIRE(!dof_vel || t->dvs, PHC_EINVAL, "%s: dvs table missing", fn);
    PHC_REQUIRE(!body_state || env_stride >= NB * REC, PHC_ESHAPE, "%s: env_stride=%lld < 312", fn, (long long)env_stride);
    PHC_REQUIRE(aligned16(t->grs) && aligned16(t->lrs), PHC_EALIGN, "%s: grs/lrs tables must be 16-byte aligned", fn);
    ResetArgs2 a{*t, env_ids, sampled_motion_ids, motion_times, global_offset, K, root_states, dof_pos, dof_vel, body_state, env_stride,
                 ref_device};
    reset_ref_state_kernel<<<(unsigned)((K + MS_WARPS - 1) / MS_WARPS), MS_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}

extern "C" int phc_sample_time_interval(const float* phase, const float* motion_len, int64_t n, int div_mode, float* out,
                                        phc_stream_t stream) {
    PHC_REQUIRE(n >= 0, PHC_EINVAL, "phc_sample_time_interval: n < 0");
    if (n == 0) return PHC_OK;
    PHC_REQUIRE(phase && motion_len && out, PHC_EINVAL, "phc_sample_time_interval: NULL pointer");
    PHC_REQUIRE(div_mode == 0 || div_mode == 1, PHC_EINVAL, "phc_sample_time_interval: div_mode must be 0 or 1");
    sample_time_interval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(phase, motion_len, n, div_mode, out);
    return check_launch("phc_sample_time_interval");
}

// MotionLibBase._calc_frame_blend on its own (motion_lib.py:655-665): thread = element.
__global__ void frame_blend_kernel(const float* __restrict__ time, const float* __restrict__ len, const int64_t* __restrict__ nf,
                                   const float* __restrict__ dt, int64_t n, int64_t* __restrict__ i0, int64_t* __restrict__ i1,
                                   float* __restrict__ blend) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t a, b;
    float bl;
    phc::frame_blend(__ldg(time + i), __ldg(len + i), __ldg(nf + i), __ldg(dt + i), a, b, bl);
    i0[i] = a; i1[i] = b; blend[i] = bl;
}

extern "C" int phc_frame_blend(const float* time, const float* len, const int64_t* num_frames, const float* dt, int64_t n,
                               int64_t* frame_idx0, int64_t* frame_idx1, float* blend, phc_stream_t stream) {
    PHC_REQUIRE(n >= 0, PHC_EINVAL, "phc_frame_blend: n < 0");
    if (n == 0) return PHC_OK;
    PHC_REQUIRE(time && len && num_frames && dt && frame_idx0 && frame_idx1 && blend, PHC_EINVAL, "phc_frame_blend: NULL pointer");
    frame_blend_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(time, len, num_frames, dt, n, frame_idx0, frame_idx1, blend);
    return check_launch("phc_frame_blend");
}

extern "C" int phc_pack_frames(const phc_motion_tables* t, float* packed, phc_stream_t stream) {
    PHC_REQUIRE(t && packed, PHC_EINVAL, "phc_pack_frames: NULL pointer");
    PHC_REQUIRE(t->gts && t->grs && t->gvs && t->gavs, PHC_EINVAL, "phc_pack_frames: gts/grs/gvs/gavs required");
    PHC_REQUIRE(aligned16(packed), PHC_EALIGN, "phc_pack_frames: packed must be 16-byte aligned");
    if (t->F <= 0) return PHC_OK;
    const int blocks = sm_count() * 8;
    pack_frames_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*t, packed);
    return check_launch("phc_pack_frames");
}
