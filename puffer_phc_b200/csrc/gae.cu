// gae.cu -- c_gae.compute_gae (reference puffer_phc/c_gae.pyx:11-32): a serial reverse scan over the
// WHOLE flat rollout array (the carry crosses env boundaries), reward/done indexed at t+1, adv[L-1] = 0:
//     nnt  = 1 - dones[t+1]
//     d    = (rewards[t+1] + (gamma * values[t+1]) * nnt) - values[t]
//     last = d + ((gamma * lambda) * nnt) * last          adv[t] = last
//
// Blocked kernel: each thread owns 32 consecutive elements and runs the SAME sequential recurrence, after
// a warm-up over the K elements to its right started from last = 0.  The unknown true carry enters the
// warm-up multiplied by prod(gamma*lambda*nnt) <= (gamma*lambda)^K, which the host chooses below 2^-40, far
// under fp32 resolution (and exactly 0 after any done), so the chunk starts from the exact carry and the
// output is bit-identical to the serial scan; chunks whose window reaches the end of the array are exact by
// construction.  Tiles are staged in shared memory with coalesced loads/stores (stride-(CH+1) padding).
// Serial kernel: one thread, used when gamma*lambda is too close to 1 for a bounded window.
// Direct kernel (short windows, K <= 64: PHC's gamma = 0.98, lambda = 0.2 gives K = 32): the same chunked recurrence without staging
// the raw arrays -- see gae_direct_kernel.
//
// PRECONDITION of the blocked kernel (mode 0 picks it silently, mode 1 forces it): FINITE inputs.  In the reference's serial scan a
// NaN / Inf in rewards or values poisons every earlier element up to the previous done; the blocked kernel would carry it only K
// elements back, and a carry more than ~2^16 times larger than the local values can differ in the last bits.  Callers that must
// reproduce the reference's behaviour on corrupted rollouts (core.py notes "Nans in adversarial reward and gae") use mode 2, which
// is the reference's scan verbatim; tests/test_gpu_parity.py holds both modes to the serial result incl. NaN and large-magnitude cases.
#include "phc_common.cuh"

namespace phc {

constexpr int GAE_THREADS = 64;     // threads per CTA
constexpr int GAE_KMAX = 2048;

// one padding word per CH elements: thread t starts at (CH+1)*t, and CH+1 is odd, so a warp hits 32 distinct banks
template <int CH> __device__ __forceinline__ int padc(int i) { return i + i / CH; }

// GAE_CH = elements owned by a thread, THREADS = threads per CTA.  This one-tile-per-CTA kernel serves the long windows (K > 64) with
// CH = 32, which keeps the redundant warm-up work at K/32 per element; short windows go to gae_direct_kernel below.  Tiles are
// staged with 16-byte loads when the three arrays allow it (tile starts are multiples of 4 elements).
template <int GAE_CH, int THREADS>
__global__ void __launch_bounds__(THREADS) gae_blocked_kernel(const float* __restrict__ dones, const float* __restrict__ values,
                                                              const float* __restrict__ rewards, int64_t L, float gamma,
                                                              float gl, int K, float* __restrict__ adv, int vec) {
    extern __shared__ float sm[];
    constexpr int GAE_TILE = GAE_CH * THREADS;
    const int span = GAE_TILE + K + 1;                 // elements [tile0, tile0 + span) are needed
    const int pspan = padc<GAE_CH>(span) + 1;
    float* s_d = sm;
    float* s_v = s_d + pspan;
    float* s_r = s_v + pspan;
    float* s_a = s_r + pspan;                          // [padc<GAE_CH>(GAE_TILE)+1]
    const int64_t tile0 = (int64_t)blockIdx.x * GAE_TILE;
    const int64_t avail = (L - tile0 < span) ? (L - tile0) : span;
    const int n4 = vec ? (int)(avail >> 2) : 0;
    const float4* d4 = reinterpret_cast<const float4*>(dones + tile0);
    const float4* v4 = reinterpret_cast<const float4*>(values + tile0);
    const float4* r4 = reinterpret_cast<const float4*>(rewards + tile0);
    for (int q = threadIdx.x; q < n4; q += THREADS) {
        const float4 a = __ldg(d4 + q), b = __ldg(v4 + q), c = __ldg(r4 + q);
        const int i = 4 * q;
        const int p0 = padc<GAE_CH>(i), p1 = padc<GAE_CH>(i + 1), p2 = padc<GAE_CH>(i + 2), p3 = padc<GAE_CH>(i + 3);
        s_d[p0] = a.x; s_d[p1] = a.y; s_d[p2] = a.z; s_d[p3] = a.w;
        s_v[p0] = b.x; s_v[p1] = b.y; s_v[p2] = b.z; s_v[p3] = b.w;
        s_r[p0] = c.x; s_r[p1] = c.y; s_r[p2] = c.z; s_r[p3] = c.w;
    }
    for (int i = 4 * n4 + threadIdx.x; i < avail; i += THREADS) {
        const int p = padc<GAE_CH>(i);
        s_d[p] = __ldg(dones + tile0 + i);
        s_v[p] = __ldg(values + tile0 + i);
        s_r[p] = __ldg(rewards + tile0 + i);
    }
    __syncthreads();
    const int c0 = threadIdx.x * GAE_CH;               // chunk [c0, c0 + CH) relative to the tile
    if (tile0 + c0 < L) {
        // highest t_cur this thread evaluates: warm-up start, clipped to L-2 (adv[L-1] = 0 starts the true scan)
        int64_t hi = tile0 + c0 + GAE_CH - 1 + K;
        if (hi > L - 2) hi = L - 2;
        float last = 0.0f;
        for (int i = (int)(hi - tile0); i >= c0; --i) {
            const int pn = padc<GAE_CH>(i + 1), pc = padc<GAE_CH>(i);
            const float nnt = 1.0f - s_d[pn];
            const float delta = (s_r[pn] + (gamma * s_v[pn]) * nnt) - s_v[pc];
            last = delta + (gl * nnt) * last;
            if (i < c0 + GAE_CH) s_a[pc] = last;
        }
        if (tile0 + c0 + GAE_CH > L - 1 && tile0 + c0 <= L - 1) s_a[padc<GAE_CH>((int)(L - 1 - tile0))] = 0.0f;
    }
    __syncthreads();
    const int64_t nout = (L - tile0 < GAE_TILE) ? (L - tile0) : GAE_TILE;
    if (vec && (nout & 3) == 0) {
        float4* o4 = reinterpret_cast<float4*>(adv + tile0);
        for (int q = threadIdx.x; q < (int)(nout >> 2); q += THREADS) {
            const int i = 4 * q;
            o4[q] = make_float4(s_a[padc<GAE_CH>(i)], s_a[padc<GAE_CH>(i + 1)], s_a[padc<GAE_CH>(i + 2)], s_a[padc<GAE_CH>(i + 3)]);
        }
    } else {
        for (int i = threadIdx.x; i < nout; i += THREADS) adv[tile0 + i] = s_a[padc<GAE_CH>(i)];
    }
}


// (delta, coefficient) pairs live in groups of CH with two pad pairs behind each group: a chain thread's group starts 8 * (CH + 2) bytes
// after its neighbour's -- an odd number of 16-byte units for CH = 8 and 16, so the 16-byte loads of a quarter warp hit distinct
// banks -- and groups of four elements stay 16-byte aligned for the precompute pass's stores.
template <int CH> __host__ __device__ __forceinline__ int padp(int i) { return i + 2 * (i / CH); }

// c_gae.pyx:24-27, the part of a step that does not involve the carry: element i needs dones / rewards / values at i + 1 and values at i
__device__ __forceinline__ float2 gae_pair(float d1, float v1, float r1, float v0, float gamma, float gl) {
    const float nnt = 1.0f - d1;
    return make_float2((r1 + (gamma * v1) * nnt) - v0, gl * nnt);
}

// ---- direct kernel for short warm-up windows ------------------------------------------------------------------------------------
// Staged variants of this kernel (the one-tile-per-CTA kernel above with 8-element chunks; a persistent cp.async-pipelined kernel with
// 2 or 3 stages, 8- or 16-element chunks, 8 to 24 warps per SM) all measured 10-13 us for 65536 x 32 whatever their occupancy and
// instruction count: they move ~75 bytes per element through SHARED memory (raw arrays in, raw arrays out again for the pairs, five
// chain reads per element), which at 128 B/clk per SM is the whole run time (profiles/README.md).  This kernel touches shared memory
// only with the pairs (8.0 us):
// global -> registers (16-byte loads, the element behind a group of four comes from the neighbour lane) -> (delta, coefficient) pairs in
// shared memory (8 B per element, written once) -> chains of CH = 16 elements (3 x 8 B read per element) -> advantages from registers
// straight to global memory.
template <int CH, int THREADS, int TILE>
__global__ void __launch_bounds__(THREADS) gae_direct_kernel(const float* __restrict__ dones, const float* __restrict__ values,
                                                             const float* __restrict__ rewards, int64_t L, float gamma, float gl,
                                                             int K, float* __restrict__ adv, int vec) {
    extern __shared__ float4 gae_sm4[];
    float2* pairs = reinterpret_cast<float2*>(gae_sm4);          // [padp(span)]
    constexpr int CHAINS = TILE / CH, PASSES = TILE / 4 / THREADS;
    static_assert(CHAINS <= THREADS && PASSES >= 1 && TILE % (4 * THREADS) == 0, "tile geometry");
    const int tid = threadIdx.x, lane = tid & 31;
    const int span = TILE + K + 1;
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    const int avail = (int)((L - tile0 < span) ? (L - tile0) : span);
    const int nelem = avail - 1;                       // elements i in [0, nelem) have a successor inside [0, avail): they get a pair
    const float* gd = dones + tile0;
    const float* gv = values + tile0;
    const float* gr = rewards + tile0;
    const int nq = vec ? (nelem >> 2) : 0;
    auto pair_group = [&](int q, const float4& d, const float4& v, const float4& r, float dn, float vn, float rn) {
        const float2 p0 = gae_pair(d.y, v.y, r.y, v.x, gamma, gl), p1 = gae_pair(d.z, v.z, r.z, v.y, gamma, gl);
        const float2 p2 = gae_pair(d.w, v.w, r.w, v.z, gamma, gl), p3 = gae_pair(dn, vn, rn, v.w, gamma, gl);
        float4* dst = reinterpret_cast<float4*>(pairs + padp<CH>(4 * q));
        dst[0] = make_float4(p0.x, p0.y, p1.x, p1.y);
        dst[1] = make_float4(p2.x, p2.y, p3.x, p3.y);
    };
    // element 4q + 4: the first element of the neighbour lane's group when that lane holds one, else one scalar load
    auto successor = [&](int q, bool ok, float mine, const float* g) {
        float nx = __shfl_down_sync(FULL, mine, 1);
        if (ok && (lane == 31 || q + 1 >= nq)) nx = __ldg(g + 4 * q + 4);
        return nx;
    };
    {   // the tile proper: all loads of the thread's passes are in flight before the first pair is formed
        float4 d[PASSES], v[PASSES], r[PASSES];
#pragma unroll
        for (int u = 0; u < PASSES; ++u) {
            const int q = tid + u * THREADS;
            d[u] = v[u] = r[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (q < nq) {
                d[u] = __ldg(reinterpret_cast<const float4*>(gd) + q);
                v[u] = __ldg(reinterpret_cast<const float4*>(gv) + q);
                r[u] = __ldg(reinterpret_cast<const float4*>(gr) + q);
            }
        }
#pragma unroll
        for (int u = 0; u < PASSES; ++u) {
            const int q = tid + u * THREADS;
            const bool ok = q < nq;
            const float dn = successor(q, ok, d[u].x, gd), vn = successor(q, ok, v[u].x, gv), rn = successor(q, ok, r[u].x, gr);
            if (ok) pair_group(q, d[u], v[u], r[u], dn, vn, rn);
        }
    }
    for (int q = tid + PASSES * THREADS; q - lane < nq; q += THREADS) {             // the warm-up halo behind the tile (whole warps)
        const bool ok = q < nq;
        float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f), v = d, r = d;
        if (ok) {
            d = __ldg(reinterpret_cast<const float4*>(gd) + q);
            v = __ldg(reinterpret_cast<const float4*>(gv) + q);
            r = __ldg(reinterpret_cast<const float4*>(gr) + q);
        }
        const float dn = successor(q, ok, d.x, gd), vn = successor(q, ok, v.x, gv), rn = successor(q, ok, r.x, gr);
        if (ok) pair_group(q, d, v, r, dn, vn, rn);
    }
    for (int i = 4 * nq + tid; i < nelem; i += THREADS)                              // ragged end / unaligned arrays
        pairs[padp<CH>(i)] = gae_pair(__ldg(gd + i + 1), __ldg(gv + i + 1), __ldg(gr + i + 1), __ldg(gv + i), gamma, gl);
    __syncthreads();
    if (tid >= CHAINS) return;
    const int c0 = tid * CH;
    const int64_t g0 = tile0 + c0;
    if (g0 >= L) return;
    int64_t hi = g0 + CH - 1 + K;                      // warm-up start, clipped to L-2 (adv[L-1] = 0 starts the true scan)
    if (hi > L - 2) hi = L - 2;
    float last = 0.0f;
    int i = (int)(hi - tile0);
    // warm-up: whole groups of CH steps, loads ahead of the dependent FMUL / FADD chain (an unclipped window ends on a group's last element)
#pragma unroll 2
    for (; i - (CH - 1) >= c0 + CH && (i & (CH - 1)) == CH - 1; i -= CH) {
        const float4* pp = reinterpret_cast<const float4*>(pairs + padp<CH>(i - (CH - 1)));
        float4 p[CH / 2];
#pragma unroll
        for (int u = 0; u < CH / 2; ++u) p[u] = pp[u];
#pragma unroll
        for (int u = CH / 2 - 1; u >= 0; --u) {
            last = p[u].z + p[u].w * last;             // c_gae.pyx:28
            last = p[u].x + p[u].y * last;
        }
    }
    for (; i >= c0 + CH; --i) {
        const float2 p = pairs[padp<CH>(i)];
        last = p.x + p.y * last;
    }
    float o[CH];
    {
        const float4* pp = reinterpret_cast<const float4*>(pairs + padp<CH>(c0));              // the chunk is one group
        float4 p[CH / 2];
#pragma unroll
        for (int u = 0; u < CH / 2; ++u) p[u] = pp[u];                                         // (slots past L-2 hold stale values, unused)
#pragma unroll
        for (int u = CH / 2 - 1; u >= 0; --u) {
            if (g0 + 2 * u + 1 <= L - 2) { last = p[u].z + p[u].w * last; o[2 * u + 1] = last; } else o[2 * u + 1] = 0.0f;   // adv[L-1] = 0
            if (g0 + 2 * u <= L - 2) { last = p[u].x + p[u].y * last; o[2 * u] = last; } else o[2 * u] = 0.0f;
        }
    }
    if (vec && g0 + CH <= L) {
        float4* o4 = reinterpret_cast<float4*>(adv + g0);
#pragma unroll
        for (int u = 0; u < CH / 4; ++u) o4[u] = make_float4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (g0 + k < L) adv[g0 + k] = o[k];
    }
}

// one warp: coalesced staging of 1024-element chunks, lane 0 runs the recurrence.
constexpr int GAE_SER_CHUNK = 1024;
__global__ void __launch_bounds__(32) gae_serial_kernel(const float* __restrict__ dones, const float* __restrict__ values,
                                                        const float* __restrict__ rewards, int64_t L, float gamma, float gl,
                                                        float* __restrict__ adv) {
    __shared__ float s_d[GAE_SER_CHUNK + 1], s_v[GAE_SER_CHUNK + 1], s_r[GAE_SER_CHUNK + 1], s_a[GAE_SER_CHUNK];
    const int lane = threadIdx.x;
    float last = 0.0f;
    if (lane == 0) adv[L - 1] = 0.0f;
    // chunks of t_cur in [lo, hi), walking down from L-1 (exclusive)
    for (int64_t hi = L - 1; hi > 0; hi -= GAE_SER_CHUNK) {
        const int64_t lo = hi > GAE_SER_CHUNK ? hi - GAE_SER_CHUNK : 0;
        const int n = (int)(hi - lo);
        for (int i = lane; i <= n; i += 32) {          // needs elements [lo, hi]
            s_d[i] = __ldg(dones + lo + i);
            s_v[i] = __ldg(values + lo + i);
            s_r[i] = __ldg(rewards + lo + i);
        }
        __syncwarp();
        if (lane == 0)
            for (int i = n - 1; i >= 0; --i) {
                const float nnt = 1.0f - s_d[i + 1];
                const float delta = (s_r[i + 1] + (gamma * s_v[i + 1]) * nnt) - s_v[i];
                last = delta + (gl * nnt) * last;
                s_a[i] = last;
            }
        __syncwarp();
        for (int i = lane; i < n; i += 32) adv[lo + i] = s_a[i];
        __syncwarp();
    }
}

}  // namespace phc

using namespace phc;

extern "C" int phc_gae(const float* dones, const float* values, const float* rewards, int64_t L, float gamma, float gae_lambda,
                       float* advantages, int mode, phc_stream_t stream) {
    const char* fn = "phc_gae";
    PHC_REQUIRE(L >= 0, PHC_EINVAL, "%s: L < 0", fn);
    PHC_REQUIRE(mode >= 0 && mode <= 2, PHC_EINVAL, "%s: mode must be 0, 1 or 2", fn);
    if (L == 0) return PHC_OK;
    PHC_REQUIRE(dones && values && rewards && advantages, PHC_EINVAL, "%s: NULL pointer", fn);
    cudaStream_t s = (cudaStream_t)stream;
    const float gl = gamma * gae_lambda;                  // formed in fp32 like the reference's C floats
    // warm-up window: (|gl|)^K <= 2^-40
    int K = -1;
    const float agl = fabsf(gl);
    constexpr int GAE_CH = 32;                            // K is kept a multiple of 32 (shared-memory padding period)
    if (agl == 0.0f) K = GAE_CH;
    else if (agl < 1.0f) {
        const double k = 40.0 * 0.6931471805599453 / -log((double)agl);
        if (k <= (double)GAE_KMAX) { K = ((int)ceil(k) + GAE_CH - 1) / GAE_CH * GAE_CH; if (K < GAE_CH) K = GAE_CH; }
    }
    if (mode == 1 && K < 0) return fail(PHC_EUNSUPPORTED, "%s: gamma*lambda=%g needs a warm-up window > %d; use mode 0 or 2", fn, (double)gl, GAE_KMAX);
    if (mode == 2 || K < 0) {
        gae_serial_kernel<<<1, 32, 0, s>>>(dones, values, rewards, L, gamma, gl, advantages);
        return check_launch(fn);
    }
    const int vec = aligned16(dones) && aligned16(values) && aligned16(rewards) && aligned16(advantages);
    if (K <= 64) {
        // short window: the direct kernel.  4096-element tiles (256 chains of 16) once the rollout fills the GPU, else 512-element
        // tiles (64 chains of 8): more CTAs in flight
#ifndef GAE_DCH
#define GAE_DCH 16                                    // elements per chain thread of the big tiles (A/B: 8 -> 8.6 us, 16 -> 8.0 us, 32 -> 8.4 us)
#endif
        constexpr int BIG_TILE = 256 * GAE_DCH;
        const bool big = (L + 2047) / 2048 >= 2 * (int64_t)sm_count();
        const int tile = big ? BIG_TILE : 512;
        const int ch = big ? GAE_DCH : 8;
        const int span = tile + K + 1;
        const size_t smem = (size_t)(span + 2 * (span / ch) + 4) * sizeof(float2);
        const unsigned grid = (unsigned)((L + tile - 1) / tile);
        if (big) {
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(gae_direct_kernel<GAE_DCH, 256, BIG_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
            }
            gae_direct_kernel<GAE_DCH, 256, BIG_TILE><<<grid, 256, smem, s>>>(dones, values, rewards, L, gamma, gl, K, advantages, vec);
        } else {
            gae_direct_kernel<8, 64, 512><<<grid, 64, smem, s>>>(dones, values, rewards, L, gamma, gl, K, advantages, vec);
        }
        return check_launch(fn);
    }
    {
        const int ch = 32;
        const int threads = GAE_THREADS;
        const int tile = ch * threads;
        const int span = tile + K + 1;
        const size_t smem = (size_t)(3 * ((span + span / ch) + 1) + (tile + tile / ch) + 1) * sizeof(float);
        const int64_t tiles = (L + tile - 1) / tile;
        if (smem > 48 * 1024) {      // per-device attribute and a cheap call: set it on every launch that needs it (no per-thread cache)
            cudaError_t e = cudaFuncSetAttribute(gae_blocked_kernel<32, GAE_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
        }
        gae_blocked_kernel<32, GAE_THREADS><<<(unsigned)tiles, GAE_THREADS, smem, s>>>(dones, values, rewards, L, gamma, gl, K, advantages, vec);
    }
    return check_launch(fn);
}
