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
// Pipelined kernel (short windows, K <= 64: PHC's gamma = 0.98, lambda = 0.2 gives K = 32): the same chunked recurrence, but a
// CTA walks several tiles with the raw arrays of the next two tiles in flight (cp.async, three stages), the per-element terms that do
// not depend on the carry -- delta and gamma*lambda*nnt, identical for every chunk whose window covers the element -- are computed ONCE
// per tile by all threads, so a chain step is one 8-byte shared-memory load + FMUL + FADD, and every thread stores its eight
// advantages straight to global memory (a warp's stores are contiguous).  The one-tile-per-CTA kernel spent its time in load ->
// barrier -> compute -> barrier -> store round trips with nothing in flight behind them.
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
// CH = 32, which keeps the redundant warm-up work at K/32 per element; short windows go to gae_pipelined_kernel below.  Tiles are
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


// ---- pipelined kernel for short warm-up windows ---------------------------------------------------------------------------
__device__ __forceinline__ void gae_cp16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gae_cp4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
constexpr int GAE_STAGES = 3;
constexpr int GAE_PCH = 8;          // elements per chain thread

__host__ __device__ inline int gae_pipe_rspan(int tile, int K) { return (tile + K + 1 + 3) & ~3; }     // floats per raw array per stage
__host__ __device__ inline size_t gae_pipe_smem(int tile, int K) {
    const int span = tile + K + 1;
    return (size_t)GAE_STAGES * 3 * gae_pipe_rspan(tile, K) * sizeof(float) + (size_t)(span + span / GAE_PCH + 1) * sizeof(float2);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) gae_pipelined_kernel(const float* __restrict__ dones, const float* __restrict__ values,
                                                                const float* __restrict__ rewards, int64_t L, float gamma, float gl,
                                                                int K, float* __restrict__ adv, int vec, int64_t tiles) {
    extern __shared__ float4 gae_sm4[];
    constexpr int CH = GAE_PCH, TILE = CH * THREADS;
    const int span = TILE + K + 1;                     // raw elements [tile0, tile0 + span) feed a tile
    const int rspan = gae_pipe_rspan(TILE, K);
    float* raw = reinterpret_cast<float*>(gae_sm4);    // [GAE_STAGES][3][rspan]   dones | values | rewards
    float2* pairs = reinterpret_cast<float2*>(raw + GAE_STAGES * 3 * rspan);   // [padc(span)]  (delta, gamma*lambda*nnt) per element
    const int tid = threadIdx.x;

    auto issue = [&](int64_t tile, int stage) {
        if (tile < tiles) {
            const int64_t tile0 = tile * TILE;
            const int avail = (int)((L - tile0 < span) ? (L - tile0) : span);
            float* sd = raw + stage * 3 * rspan;
            float* sv = sd + rspan;
            float* sr = sv + rspan;
            const int n4 = vec ? (avail >> 2) : 0;
            for (int q = tid; q < n4; q += THREADS) {
                gae_cp16(sd + 4 * q, dones + tile0 + 4 * q);
                gae_cp16(sv + 4 * q, values + tile0 + 4 * q);
                gae_cp16(sr + 4 * q, rewards + tile0 + 4 * q);
            }
            for (int i = 4 * n4 + tid; i < avail; i += THREADS) {
                gae_cp4(sd + i, dones + tile0 + i);
                gae_cp4(sv + i, values + tile0 + i);
                gae_cp4(sr + i, rewards + tile0 + i);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");          // one group per pipeline slot, empty past the last tile
    };

    int64_t tile = blockIdx.x;
#pragma unroll
    for (int s = 0; s < GAE_STAGES - 1; ++s) issue(tile + (int64_t)s * gridDim.x, s);
    for (int it = 0; tile < tiles; tile += gridDim.x, ++it) {
        const int stage = it % GAE_STAGES;
        // the stage refilled here was consumed by iteration it - 1, which every thread left through the barrier below
        issue(tile + (int64_t)(GAE_STAGES - 1) * gridDim.x, (it + GAE_STAGES - 1) % GAE_STAGES);
        asm volatile("cp.async.wait_group %0;" ::"n"(GAE_STAGES - 1) : "memory");
        __syncthreads();                               // this tile's raw arrays are visible; the previous tile's chains are done with pairs[]
        const int64_t tile0 = tile * TILE;
        const int avail = (int)((L - tile0 < span) ? (L - tile0) : span);
        const float* sd = raw + stage * 3 * rspan;
        const float* sv = sd + rspan;
        const float* sr = sv + rspan;
        for (int i = tid; i + 1 < avail; i += THREADS) {             // c_gae.pyx:24-27, the part that does not involve the carry
            const float nnt = 1.0f - sd[i + 1];
            const float delta = (sr[i + 1] + (gamma * sv[i + 1]) * nnt) - sv[i];
            pairs[padc<CH>(i)] = make_float2(delta, gl * nnt);
        }
        __syncthreads();
        const int c0 = tid * CH;
        const int64_t g0 = tile0 + c0;
        if (g0 < L) {
            int64_t hi = g0 + CH - 1 + K;              // warm-up start, clipped to L-2 (adv[L-1] = 0 starts the true scan)
            if (hi > L - 2) hi = L - 2;
            float last = 0.0f;
            int i = (int)(hi - tile0);
            // warm-up: groups of CH steps with the CH loads issued ahead of the dependent FMUL / FADD chain (an unclipped window starts
            // on the last element of a padding group, so the CH offsets are compile-time constants)
            for (; i - (CH - 1) >= c0 + CH && (i & (CH - 1)) == CH - 1; i -= CH) {
                const float2* pp = pairs + padc<CH>(i - (CH - 1));
                float2 p[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) p[u] = pp[u];
#pragma unroll
                for (int u = CH - 1; u >= 0; --u) last = p[u].x + p[u].y * last;               // c_gae.pyx:28
            }
            for (; i >= c0 + CH; --i) {
                const float2 p = pairs[padc<CH>(i)];
                last = p.x + p.y * last;
            }
            float o[CH];
            {
                const float2* pp = pairs + padc<CH>(c0);            // the chunk is one padding group: CH consecutive pairs
                float2 p[CH];
#pragma unroll
                for (int k = 0; k < CH; ++k) p[k] = pp[k];          // (slots past L-2 hold stale values that are never used)
#pragma unroll
                for (int k = CH - 1; k >= 0; --k) {
                    if (g0 + k <= L - 2) {
                        last = p[k].x + p[k].y * last;
                        o[k] = last;
                    } else {
                        o[k] = 0.0f;                   // adv[L-1]
                    }
                }
            }
            if (vec && g0 + CH <= L) {
                float4* o4 = reinterpret_cast<float4*>(adv + g0);
                o4[0] = make_float4(o[0], o[1], o[2], o[3]);
                o4[1] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int k = 0; k < CH; ++k)
                    if (g0 + k < L) adv[g0 + k] = o[k];
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
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
        // short window: the pipelined kernel.  2048-element tiles once there are two per SM, else 512-element tiles (more CTAs in
        // flight for rollouts that do not fill the GPU); CTAs walk equal numbers of tiles.
        const int sms = sm_count();
        const bool big = (L + 2047) / 2048 >= 2 * (int64_t)sms;
        const int tile = big ? 2048 : 512;
        const size_t smem = gae_pipe_smem(tile, K);
        const int64_t tiles = (L + tile - 1) / tile;
        const int64_t resident = (int64_t)sms * (big ? 2 : 8);
        const int64_t per = (tiles + resident - 1) / resident;
        const unsigned grid = (unsigned)((tiles + per - 1) / per);
        if (big) {
            cudaError_t e = cudaFuncSetAttribute(gae_pipelined_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
            gae_pipelined_kernel<256><<<grid, 256, smem, s>>>(dones, values, rewards, L, gamma, gl, K, advantages, vec, tiles);
        } else {
            gae_pipelined_kernel<64><<<grid, 64, smem, s>>>(dones, values, rewards, L, gamma, gl, K, advantages, vec, tiles);
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
