// build_tables.cu -- the motion TABLE BUILD ("next" row f4): what MotionLibSMPL.load_motion_with_skeleton
// (reference puffer_phc/motion_lib.py:744-825), SkeletonState.local_rotation / global_transformation
// (poselib_skeleton.py:516-536, 575-591), SkeletonMotion._compute_velocity / _compute_angular_velocity
// (poselib_skeleton.py:1228-1249) and compute_motion_dof_vels_jit (motion_lib.py:119-140) turn raw clips into,
// for ALL clips of a library in ONE launch, written straight into the concatenated tables load_motions builds
// (motion_lib.py:405-412) and, optionally, into the packed 1248-byte frame records the fused step gathers.
//
// The reference does this clip by clip on the host (torch CPU + numpy + scipy; "~20s for 4096 envs",
// motion_lib.py:404).  Here a CTA owns a tile of BT_TILE consecutive frames of one clip:
//   phase A  warp = frame, lane = body: float64 local rotation (rounded to float32 like the reference's float32
//            local_rotation tensor) for tile + halo (9 frames either side: gaussian radius 8 + gradient 1), into
//            shared memory;
//   phase D  thread = (frame, joint): float32 dof velocities from the local rotations;
//   phase FK lane = frame (two warps), bodies walked in tree order: float32 forward kinematics with the parent's
//            rotation / position read back from shared memory (conflict-free strides; the global rotations overwrite
//            the local ones in place) -- every lane busy, instead of one tree level per warp step.  The other six
//            warps compute the float64 frame-to-frame angular velocity meanwhile (warp = frame, lane = body);
//   phase G  np.gradient / float32(1/fps) of the float32 positions;
//   phase B  thread = (frame, component): the sigma-2 gaussian (scipy correlate1d, symmetric form, "nearest"
//            edges, double accumulation) of both velocities, coalesced row stores.  Nothing intermediate goes to HBM.
// The precision mix (f64 rotations, f32 FK, f32 gradient, f64 filter) is the reference's own and is what makes
// positions / velocities bit-identical to it (DESIGN.md section 2 lists the precision of every stage).
// Compiled with -fmad=false: every operation individually rounded.
#include <math.h>

#include "phc_common.cuh"

namespace phc {

constexpr int BT_TILE = 32;                 // frames per CTA (PHC_BUILD_TILE in the header)
constexpr int BT_R = 8;                     // gaussian radius int(4 * sigma + 0.5), sigma = 2
constexpr int BT_HALO = BT_R + 1;           // + 1 for the central difference
constexpr int BT_THREADS = 256;
constexpr int BT_WARPS = BT_THREADS / 32;
constexpr int BT_PROWS = BT_TILE + 2 * BT_HALO;   // position rows held
constexpr int BT_VROWS = BT_TILE + 2 * BT_R;      // velocity rows held

static_assert(BT_TILE == PHC_BUILD_TILE, "header and kernel disagree on the tile size");

template <typename T> struct Quat { T x, y, z, w; };

// torch_utils.py:55-75 quat_mul, the 8-multiply form, for float and double.
template <typename T>
__device__ __forceinline__ Quat<T> qmul(Quat<T> a, Quat<T> b) {
    T ww = (a.z + a.x) * (b.x + b.y);
    T yy = (a.w - a.y) * (b.w + b.z);
    T zz = (a.w + a.y) * (b.w - b.z);
    T xx = (ww + yy) + zz;
    T qq = T(0.5) * (xx + (a.z - a.x) * (b.x - b.y));
    Quat<T> r;
    r.w = (qq - ww) + (a.z - a.y) * (b.y - b.z);
    r.x = (qq - xx) + (a.x + a.w) * (b.x + b.w);
    r.y = (qq - yy) + (a.w - a.x) * (b.y + b.z);
    r.z = (qq - zz) + (a.z + a.y) * (b.w - b.x);
    return r;
}
template <typename T> __device__ __forceinline__ Quat<T> qconj(Quat<T> a) { return Quat<T>{-a.x, -a.y, -a.z, a.w}; }

// torch_utils.py:154-196 quat_normalize = quat_unit(quat_pos(q)); torch's CPU 4-wide norm sums left to right.
template <typename T>
__device__ __forceinline__ Quat<T> qnormalize(Quat<T> q) {
    const T s = (q.w < T(0)) ? T(-1) : T(1);
    q.x *= s; q.y *= s; q.z *= s; q.w *= s;
    T n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    if (n < T(1e-9)) n = T(1e-9);
    return Quat<T>{q.x / n, q.y / n, q.z / n, q.w / n};
}

__device__ __forceinline__ Quat<double> ldq(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return Quat<double>{a.x, a.y, b.x, b.y};
}

struct BuildArgs {
    phc_build_in in;
    phc_build_out out;
    double w[2 * BT_R + 1];     // gaussian weights (host-computed, scipy's formula)
};

// scipy ni_filters.c NI_Correlate1D, symmetric branch: tmp = x[l]*w[c]; tmp += (x[l+jj] + x[l-jj]) * w[jj+c].
// `line` points at row `lo` of a shared-memory column with row stride C; indices are clamped to [0, nf-1].
template <bool CLAMP, typename T>
__device__ __forceinline__ double filter_at(const T* col, int C, int lo, int nf, int l, const double* w) {
    const T* x = col + (l - lo) * C;                    // interior tiles (CLAMP = false): every tap is an immediate offset
    double tmp = (double)x[0] * w[BT_R];
#pragma unroll
    for (int jj = -BT_R; jj < 0; ++jj) {
        int a = jj, b = -jj;
        if (CLAMP) {
            a = l + jj < 0 ? -l : jj;
            b = l - jj > nf - 1 ? nf - 1 - l : -jj;
        }
        tmp = fma((double)x[a * C] + (double)x[b * C], w[jj + BT_R], tmp);      // fused: ~1e-16 from scipy's mul-then-add
    }
    return tmp;
}

// JT = 24: the SMPL humanoid, every row stride / tap offset is a compile-time constant (the index arithmetic otherwise rivals
// the floating-point work); JT = 0: J read from the arguments.
template <int JT>
__global__ void __launch_bounds__(BT_THREADS, 3) build_tables_kernel(const __grid_constant__ BuildArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const phc_build_in& in = A.in;
    const phc_build_out& o = A.out;
    const int J = JT ? JT : in.J, C = 3 * J;
    const int QS = 4 * J + 4;        // quaternion row stride (floats): lane = frame float4 accesses are bank-conflict free
    const int PS = C + 1;            // position row stride (floats): odd, lane = frame scalar accesses are conflict free
    double* sW = reinterpret_cast<double*>(smem_raw);                    // [BT_VROWS][C]  raw angular velocity (float64)
    float* sL = reinterpret_cast<float*>(sW + BT_VROWS * C);             // [BT_PROWS][QS] local rotations (float32); the forward
    float* sG = sL;                                                      // kinematics overwrites them in place with its global
    float* sV = sL;                                                      // rotations; then [BT_VROWS][C] raw linear velocity
    float* sP = sL + BT_PROWS * QS;                                      // [BT_PROWS][PS] positions
    __shared__ int s_clip;
    __shared__ int s_parent[32];

    // ---- which clip / tile: binary search of the tile prefix sum ------------------------------------
    if (threadIdx.x == 0) {
        const int64_t tile = blockIdx.x;
        int64_t lo = 0, hi = in.M;                       // largest m with tile_prefix[m] <= tile
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(in.tile_prefix + mid) <= tile) lo = mid; else hi = mid;
        }
        s_clip = (int)lo;
    }
    if (threadIdx.x < 32) s_parent[threadIdx.x] = threadIdx.x < J ? __ldg(in.parents + threadIdx.x) : -1;
    __syncthreads();
    const int64_t m = s_clip;
    const int nf = (int)__ldg(in.num_frames + m);
    const int64_t src0 = __ldg(in.in_start + m), dst0 = __ldg(in.out_start + m);
    const int a = (int)((int64_t)blockIdx.x - __ldg(in.tile_prefix + m)) * BT_TILE;
    const int b = min(a + BT_TILE, nf);
    const int lo9 = max(a - BT_HALO, 0), hi9 = min(b + BT_HALO, nf);
    const int lo8 = max(a - BT_R, 0), hi8 = min(b + BT_R, nf);
    const int fps = __ldg(in.fps + m);
    const double time_delta = 1.0 / (double)fps;                          // poselib_skeleton.py:1183
    const bool has_heading = in.heading != nullptr;
    double hs = 0.0, hc = 1.0;
    if (has_heading) { const double th = __ldg(in.heading + m); hs = sin(0.5 * th); hc = cos(0.5 * th); }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // motion_lib.py:789-799 random heading: scipy's Rotation product h * from_quat(q) (from_quat normalises, the
    // product is normalised again) and trans @ R(h)^T.  h = (0, 0, sin(th/2), cos(th/2)).
    auto load_rot = [&](int f, int body) -> Quat<double> {
        Quat<double> q = ldq(in.pose_quat_global + ((src0 + f) * J + body) * 4);
        if (has_heading) {
            double n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
            q.x /= n; q.y /= n; q.z /= n; q.w /= n;
            Quat<double> r{hc * q.x - hs * q.y, hc * q.y + hs * q.x, hc * q.z + q.w * hs, hc * q.w - hs * q.z};
            n = sqrt(((r.x * r.x + r.y * r.y) + r.z * r.z) + r.w * r.w);
            q = Quat<double>{r.x / n, r.y / n, r.z / n, r.w / n};
        }
        return q;
    };

    // ---- phase A (warp = frame, lane = body): float64 local rotation -------------------------------------------
    if (lane < J) {
        const int j = lane, parent = s_parent[j];
        for (int f = lo9 + warp; f < hi9; f += BT_WARPS) {
            const Quat<double> G = load_rot(f, j);
            // poselib_skeleton.py:579-590: float64 quat_mul_norm(quat_inverse(parent), self) stored into a float32 tensor
            // (the exact divisions stay: this value is rounded to float32 and feeds the bit-exact forward kinematics)
            const Quat<double> Ld = parent < 0 ? G : qnormalize(qmul(qconj(load_rot(f, parent)), G));
            const float4 L4 = make_float4((float)Ld.x, (float)Ld.y, (float)Ld.z, (float)Ld.w);
            *reinterpret_cast<float4*>(sL + (f - lo9) * QS + j * 4) = L4;
            if (f >= a && f < b) {
                const int64_t row = dst0 + f;
                const float4 g4 = make_float4((float)G.x, (float)G.y, (float)G.z, (float)G.w);
                *reinterpret_cast<float4*>(o.grs + (row * J + j) * 4) = g4;
                *reinterpret_cast<float4*>(o.lrs + (row * J + j) * 4) = L4;
                if (o.packed) *reinterpret_cast<float4*>(o.packed + row * FRAME_F + 72 + j * 4) = g4;
            }
        }
    }
    __syncthreads();

    {
        // ---- dof velocities (motion_lib.py:119-140, float32): frame t from the local rotations of (t, t+1); the last
        // frame repeats T-2 ----
        const float dt_f = (float)(1.0 / (double)fps);
        const int JD = J - 1;
        for (int i = threadIdx.x; i < (b - a) * JD; i += BT_THREADS) {
            const int f = a + i / JD, jj = 1 + (i - (i / JD) * JD);
            float* dst = o.dvs + ((dst0 + f) * JD + (jj - 1)) * 3;
            if (nf < 2) { dst[0] = 0.0f; dst[1] = 0.0f; dst[2] = 0.0f; continue; }     // the reference raises; stay in bounds
            const int f0 = f < nf - 1 ? f : nf - 2;
            const float4 l0 = *reinterpret_cast<const float4*>(sL + (f0 - lo9) * QS + jj * 4);
            const float4 l1 = *reinterpret_cast<const float4*>(sL + (f0 + 1 - lo9) * QS + jj * 4);
            const Quat<float> d = qmul(qconj(Quat<float>{l0.x, l0.y, l0.z, l0.w}), Quat<float>{l1.x, l1.y, l1.z, l1.w});
            // torch_utils.py:86-106 quat_to_angle_axis
            const float sin_theta = sqrtf(1.0f - d.w * d.w);
            float angle = 2.0f * acosf(d.w);
            angle = atan2f(sinf(angle), cosf(angle));
            float ax = d.x / sin_theta, ay = d.y / sin_theta, az = d.z / sin_theta;
            if (!(fabsf(sin_theta) > 1e-5f)) { angle = 0.0f; ax = 0.0f; ay = 0.0f; az = 1.0f; }
            dst[0] = (ax * angle) / dt_f; dst[1] = (ay * angle) / dt_f; dst[2] = (az * angle) / dt_f;
        }
    }
    __syncthreads();     // the forward kinematics below overwrites the local rotations

    if (threadIdx.x < hi9 - lo9) {
        // ---- forward kinematics (lane = frame, bodies in tree order; float32, the reference's operation order) ----
        // poselib_skeleton.py:516-536 + torch_utils.py:322-330: r = quat_mul_norm(r_parent, local),
        // t = quat_rotate(r_parent, local_translation) + t_parent; the root keeps (local rotation, f32(root translation)).
        const int fr = threadIdx.x, f = lo9 + fr;
        const float* lt = in.local_translation + m * in.lt_clip_stride;
        for (int j = 0; j < J; ++j) {
            const int p = s_parent[j];
            const float4 L4 = *reinterpret_cast<const float4*>(sL + fr * QS + j * 4);      // consumed before slot j is overwritten
            const Quat<float> L{L4.x, L4.y, L4.z, L4.w};
            Quat<float> Gr;
            float px, py, pz;
            if (p < 0) {
                const double* tr = in.root_trans + (src0 + f) * 3;
                double tx = __ldg(tr), ty = __ldg(tr + 1), tz = __ldg(tr + 2);
                if (has_heading) {      // torch.matmul(trans, R^T), R = [[c,-s,0],[s,c,0],[0,0,1]] from the unit quaternion
                    const double r00 = 1.0 - 2.0 * (hs * hs), r01 = -2.0 * (hs * hc), r10 = 2.0 * (hs * hc);
                    const double nx = (tx * r00 + ty * r01) + tz * 0.0, ny = (tx * r10 + ty * r00) + tz * 0.0;
                    tx = nx; ty = ny;
                }
                Gr = L; px = (float)tx; py = (float)ty; pz = (float)tz;
            } else {
                const float4 P4 = *reinterpret_cast<const float4*>(sG + fr * QS + p * 4);
                const Quat<float> Pr{P4.x, P4.y, P4.z, P4.w};
                Gr = qnormalize(qmul(Pr, L));
                const Quat<float> rv = qmul(qmul(Pr, Quat<float>{__ldg(lt + j * 3), __ldg(lt + j * 3 + 1), __ldg(lt + j * 3 + 2), 0.0f}),
                                            qconj(Pr));                                                   // quat_rotate :263-269
                px = rv.x + sP[fr * PS + p * 3]; py = rv.y + sP[fr * PS + p * 3 + 1]; pz = rv.z + sP[fr * PS + p * 3 + 2];
            }
            *reinterpret_cast<float4*>(sG + fr * QS + j * 4) = make_float4(Gr.x, Gr.y, Gr.z, Gr.w);
            sP[fr * PS + j * 3] = px; sP[fr * PS + j * 3 + 1] = py; sP[fr * PS + j * 3 + 2] = pz;
        }
    } else if (threadIdx.x >= 64) {
        // ---- meanwhile, warps 2..7 (warp = frame, lane = body): raw angular velocity, float64 ----
        // poselib_skeleton.py:1238-1246: quat_mul_norm(r[t+1], quat_inverse(r[t])) -> quat_angle_axis (torch_utils.py:219-228)
        // -> axis * angle / time_delta; the last frame keeps the identity (= 0).  Nothing here feeds a float32 chain, so the
        // seven divisions are folded into two (differences ~1e-16, the output is rounded to float32).
        if (lane < J) {
            const int j = lane;
            for (int f = lo8 + warp - 2; f < hi8; f += BT_WARPS - 2) {
                double wx = 0.0, wy = 0.0, wz = 0.0;
                if (f < nf - 1) {
                    Quat<double> d = qmul(load_rot(f + 1, j), qconj(load_rot(f, j)));
                    const double sg = d.w < 0.0 ? -1.0 : 1.0;
                    const double rn = sg / fmax(sqrt(((d.x * d.x + d.y * d.y) + d.z * d.z) + d.w * d.w), 1e-9);
                    d.x *= rn; d.y *= rn; d.z *= rn; d.w *= rn;
                    double sc = 2.0 * (d.w * d.w) - 1.0;
                    sc = sc < -1.0 ? -1.0 : (sc > 1.0 ? 1.0 : sc);
                    const double k = acos(sc) / (fmax(sqrt((d.x * d.x + d.y * d.y) + d.z * d.z), 1e-9) * time_delta);
                    wx = d.x * k; wy = d.y * k; wz = d.z * k;
                }
                double* sw = sW + (f - lo8) * C + j * 3;
                sw[0] = wx; sw[1] = wy; sw[2] = wz;
            }
        }
    }
    __syncthreads();

    // ---- np.gradient over frames (float32), / float32(time_delta) (poselib_skeleton.py:1229); sV overwrites sG ----
    const float td_f = (float)time_delta;
    const int dr = BT_THREADS / C, dc = BT_THREADS - dr * C;         // i += BT_THREADS without a division per element
    for (int i = threadIdx.x, r = threadIdx.x / C, c = threadIdx.x - (threadIdx.x / C) * C; i < (hi8 - lo8) * C;
         i += BT_THREADS, r += dr + (c + dc >= C), c = c + dc >= C ? c + dc - C : c + dc) {
        const int f = lo8 + r;
        const float* col = sP + c - lo9 * PS;
        float g;
        if (nf < 2) g = 0.0f;                        // the reference raises for one-frame clips; keep the kernel in bounds
        else if (f == 0) g = (col[1 * PS] - col[0]) / 1.0f;
        else if (f == nf - 1) g = (col[(nf - 1) * PS] - col[(nf - 2) * PS]) / 1.0f;
        else g = (col[(f + 1) * PS] - col[(f - 1) * PS]) / 2.0f;
        sV[i] = g / td_f;
    }
    __syncthreads();

    // ---- gaussian filter of both velocities, positions, coalesced stores ---------------------------------
    const bool interior = a - BT_R >= 0 && b + BT_R <= nf;            // no tap of this tile leaves the clip
    for (int i = threadIdx.x, r = threadIdx.x / C, c = threadIdx.x - (threadIdx.x / C) * C; i < (b - a) * C;
         i += BT_THREADS, r += dr + (c + dc >= C), c = c + dc >= C ? c + dc - C : c + dc) {
        const int f = a + r;
        const int64_t row = dst0 + f;
        const float gv = (float)(interior ? filter_at<false>(sV + c, C, lo8, nf, f, A.w) : filter_at<true>(sV + c, C, lo8, nf, f, A.w));
        const float gav = (float)(interior ? filter_at<false>(sW + c, C, lo8, nf, f, A.w) : filter_at<true>(sW + c, C, lo8, nf, f, A.w));
        const float p = sP[(f - lo9) * PS + c];
        o.gts[row * C + c] = p;
        o.gvs[row * C + c] = gv;
        o.gavs[row * C + c] = gav;
        if (o.packed) {
            float* rec = o.packed + row * FRAME_F;
            rec[c] = p; rec[168 + c] = gv; rec[240 + c] = gav;
        }
    }
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ x, int64_t n, float* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = (float)__ldg(x + i);
}

// _motion_aa (motion_lib.py:381, 399): the float64 pose_aa rows of every loaded slot's WHOLE clip (the reference appends the
// uncropped array), cast to float32.  Warp = row; the slot of an output row is found by a warp-uniform binary search of the row prefix.
// With a heading, the root rotation vector (columns 0..2) of the rows inside the crop is replaced by
// (h * from_rotvec(rv)).as_rotvec() (motion_lib.py:793; scipy's from_rotvec / compose / as_rotvec formulas).
struct AaArgs {
    const double* pose_aa;
    const int64_t *seg_src, *seg_dst_prefix, *crop_lo, *crop_hi;
    const double* heading;
    float* out;
    int64_t S, n_rows;
    int row_len;
};

__global__ void __launch_bounds__(256) build_motion_aa_kernel(const AaArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5), wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t chunk = (a.n_rows + warps - 1) / warps;     // a warp owns a run of consecutive rows: one search, then a walk
    const int64_t row0 = wid * chunk, row1 = row0 + chunk < a.n_rows ? row0 + chunk : a.n_rows;
    if (row0 >= row1) return;
    int64_t lo = 0, hi = a.S;                              // warp-uniform search: largest slot with seg_dst_prefix <= row0
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a.seg_dst_prefix + mid) <= row0) lo = mid; else hi = mid;
    }
    int64_t seg_end = __ldg(a.seg_dst_prefix + lo + 1);
    for (int64_t row = row0; row < row1; ++row) {
        while (row >= seg_end) { ++lo; seg_end = __ldg(a.seg_dst_prefix + lo + 1); }     // empty segments are skipped too
        const int64_t r = row - __ldg(a.seg_dst_prefix + lo);
        const double* src = a.pose_aa + (__ldg(a.seg_src + lo) + r) * a.row_len;
        float* dst = a.out + row * a.row_len;
        const bool rotate = a.heading && r >= __ldg(a.crop_lo + lo) && r < __ldg(a.crop_hi + lo);
        for (int c = lane; c < a.row_len; c += 32) {
            double v = __ldg(src + c);
            if (rotate && c < 3) {
                const double th = __ldg(a.heading + lo), hs = sin(0.5 * th), hc = cos(0.5 * th);
                const double rx = __ldg(src), ry = __ldg(src + 1), rz = __ldg(src + 2);
                double angle = sqrt((rx * rx + ry * ry) + rz * rz), scale;
                if (angle <= 1e-3) { const double a2 = angle * angle; scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0; }
                else scale = sin(angle / 2.0) / angle;
                const double qx = scale * rx, qy = scale * ry, qz = scale * rz, qw = cos(angle / 2.0);
                double px = hc * qx - hs * qy, py = hc * qy + hs * qx, pz = hc * qz + qw * hs, pw = hc * qw - hs * qz;
                const double n = sqrt(((px * px + py * py) + pz * pz) + pw * pw);
                px /= n; py /= n; pz /= n; pw /= n;
                if (pw < 0.0) { px = -px; py = -py; pz = -pz; pw = -pw; }
                angle = 2.0 * atan2(sqrt((px * px + py * py) + pz * pz), pw);
                if (angle <= 1e-3) { const double a2 = angle * angle; scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0; }
                else scale = angle / sin(angle / 2.0);
                v = scale * (c == 0 ? px : (c == 1 ? py : pz));
            }
            dst[c] = (float)v;
        }
    }
}

static size_t build_smem_bytes(int J) {
    const size_t C = 3 * (size_t)J;
    return sizeof(double) * BT_VROWS * C + sizeof(float) * (BT_PROWS * (4 * (size_t)J + 4) + BT_PROWS * (C + 1));
}

}  // namespace phc

extern "C" int phc_build_motion_tables(const phc_build_in* in, const phc_build_out* out, phc_stream_t stream) {
    using namespace phc;
    PHC_REQUIRE(in && out, PHC_EINVAL, "phc_build_motion_tables: NULL argument struct");
    PHC_REQUIRE(in->pose_quat_global && in->root_trans && in->in_start && in->num_frames && in->out_start && in->fps &&
                    in->tile_prefix && in->parents && in->local_translation,
                PHC_EINVAL, "phc_build_motion_tables: NULL input pointer");
    PHC_REQUIRE(out->gts && out->grs && out->lrs && out->gvs && out->gavs && out->dvs, PHC_EINVAL,
                "phc_build_motion_tables: NULL output table");
    PHC_REQUIRE(in->M >= 0 && in->n_tiles >= 0 && in->lt_clip_stride >= 0, PHC_EINVAL, "phc_build_motion_tables: negative size");
    PHC_REQUIRE(in->J >= 2 && in->J <= 32, PHC_ESHAPE, "phc_build_motion_tables: J=%d outside [2,32]", in->J);
    PHC_REQUIRE(!out->packed || in->J == NB, PHC_ESHAPE, "phc_build_motion_tables: packed records need J == 24");
    PHC_REQUIRE(aligned16(in->pose_quat_global) && aligned16(out->grs) && aligned16(out->lrs) && (!out->packed || aligned16(out->packed)),
                PHC_EALIGN, "phc_build_motion_tables: pose_quat_global / grs / lrs / packed must be 16-byte aligned");
    PHC_REQUIRE(in->n_tiles < (int64_t)1 << 31, PHC_ESHAPE, "phc_build_motion_tables: too many tiles");
    if (in->M == 0 || in->n_tiles == 0) return PHC_OK;

    BuildArgs args;
    args.in = *in;
    args.out = *out;
    {   // scipy.ndimage._filters._gaussian_kernel1d(2, 0, 8): exp(-0.5/sigma^2 x^2) / sum, the sum in numpy's pairwise order
        const double sigma2 = 4.0;
        double r[8];
        for (int k = 0; k < 2 * BT_R + 1; ++k) { const double x = (double)(k - BT_R); args.w[k] = exp(-0.5 / sigma2 * (x * x)); }
        for (int k = 0; k < 8; ++k) r[k] = args.w[k];
        for (int k = 0; k < 8; ++k) r[k] += args.w[8 + k];
        double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        s += args.w[16];
        for (int k = 0; k < 2 * BT_R + 1; ++k) args.w[k] = args.w[k] / s;
    }
    const size_t smem = build_smem_bytes(in->J);
    {   // per device; cheap enough to repeat on every call of a load-time function
        cudaError_t e = cudaFuncSetAttribute(build_tables_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_smem_bytes(32));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(build_tables_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_smem_bytes(NB));
        if (e != cudaSuccess) return fail((int)e, "phc_build_motion_tables: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    if (in->J == NB) build_tables_kernel<NB><<<(unsigned)in->n_tiles, BT_THREADS, smem, (cudaStream_t)stream>>>(args);
    else build_tables_kernel<0><<<(unsigned)in->n_tiles, BT_THREADS, smem, (cudaStream_t)stream>>>(args);
    return check_launch("phc_build_motion_tables");
}

extern "C" int phc_cast_f64_f32(const double* x, int64_t n, float* y, phc_stream_t stream) {
    using namespace phc;
    PHC_REQUIRE(n >= 0, PHC_EINVAL, "phc_cast_f64_f32: negative size");
    if (n == 0) return PHC_OK;
    PHC_REQUIRE(x && y, PHC_EINVAL, "phc_cast_f64_f32: NULL pointer");
    const int64_t blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (int64_t)sm_count() * 16 ? blocks : (int64_t)sm_count() * 16);
    cast_f64_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, y);
    return check_launch("phc_cast_f64_f32");
}

extern "C" int phc_build_motion_aa(const double* pose_aa, int row_len, const int64_t* seg_src, const int64_t* seg_dst_prefix, int64_t S,
                                   int64_t n_rows, const double* heading, const int64_t* crop_lo, const int64_t* crop_hi, float* out,
                                   phc_stream_t stream) {
    using namespace phc;
    PHC_REQUIRE(S >= 0 && n_rows >= 0 && row_len >= 3, PHC_EINVAL, "phc_build_motion_aa: bad size");
    if (S == 0 || n_rows == 0) return PHC_OK;
    PHC_REQUIRE(pose_aa && seg_src && seg_dst_prefix && out, PHC_EINVAL, "phc_build_motion_aa: NULL pointer");
    PHC_REQUIRE(!heading || (crop_lo && crop_hi), PHC_EINVAL, "phc_build_motion_aa: heading needs crop_lo / crop_hi");
    AaArgs a{pose_aa, seg_src, seg_dst_prefix, crop_lo, crop_hi, heading, out, S, n_rows, row_len};
    const int64_t blocks = (n_rows + 63) / 64;                // a warp walks >= 8 consecutive rows
    const int grid = (int)(blocks < (int64_t)sm_count() * 16 ? blocks : (int64_t)sm_count() * 16);
    build_motion_aa_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("phc_build_motion_aa");
}
