// rms.cu -- RunningNorm (reference puffer_phc/policies/running_norm.py:5-53).
//   forward  (:15-20)  y = clamp((x - mean) / sqrt(var + eps), -clip, clip)
//   update   (:23-34)  split into   moments (per-column sum / sum of squares, fp64, deterministic)
//                                   [all-reduce across ranks happens here, on the moments buffer]
//                                   finalize (batch mean / biased var -> running average, count += 1)
// Memory-bound column pass: thread = column (coalesced rows), 8 independent rows in flight per thread.
#include "phc_common.cuh"

namespace phc {

constexpr int RMS_THREADS = 256;

__device__ __forceinline__ float norm_clamp(float x, float mean, float den, float clip) {
    float y = (x - mean) / den;
    return (y != y) ? y : fminf(fmaxf(y, -clip), clip);     // torch.clamp propagates NaN
}

// contiguous [B*C] fast path: float4 in, float4 out.
__global__ void __launch_bounds__(RMS_THREADS) rms_forward_vec_kernel(const float4* __restrict__ x, const float* __restrict__ mean,
                                                                      const float* __restrict__ var, float eps, float clip,
                                                                      int64_t n4, int C, float4* __restrict__ y) {
    extern __shared__ float sm[];
    float* s_mean = sm;
    float* s_den = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s_mean[c] = __ldg(mean + c); s_den[c] = sqrtf(__ldg(var + c) + eps); }
    __syncthreads();
    // four independent 16-byte loads in flight per thread before any arithmetic (the division chain is long)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + k * stride;
            if (i < n4) v[k] = __ldg(x + i);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + k * stride;
            if (i < n4) {
                int c = (int)((i << 2) % C);
                float r[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    r[u] = norm_clamp(r[u], s_mean[c], s_den[c], clip);
                    c = (c + 1 == C) ? 0 : c + 1;
                }
                y[i] = make_float4(r[0], r[1], r[2], r[3]);
            }
        }
    }
}

// generic strided rows.
__global__ void __launch_bounds__(RMS_THREADS) rms_forward_kernel(const float* __restrict__ x, int64_t xs, const float* __restrict__ mean,
                                                                  const float* __restrict__ var, float eps, float clip, int64_t B,
                                                                  int C, float* __restrict__ y, int64_t ys) {
    extern __shared__ float sm[];
    float* s_mean = sm;
    float* s_den = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s_mean[c] = __ldg(mean + c); s_den[c] = sqrtf(__ldg(var + c) + eps); }
    __syncthreads();
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x)
        for (int c = threadIdx.x; c < C; c += blockDim.x) y[b * ys + c] = norm_clamp(__ldg(x + b * xs + c), s_mean[c], s_den[c], clip);
}

// partial[blockIdx.x][0][c] = sum over this block's rows of x[:,c]; [1][c] = sum of squares.  blockIdx.y = column chunk.
__global__ void __launch_bounds__(RMS_THREADS) rms_moments_kernel(const float* __restrict__ x, int64_t xs, int64_t B, int C,
                                                                  int64_t rows_per_block, double* __restrict__ partial) {
    const int c = blockIdx.y * RMS_THREADS + threadIdx.x;
    if (c >= C) return;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < B) ? r0 + rows_per_block : B;
    double s = 0.0, q = 0.0;
    int64_t r = r0;
    for (; r + 8 <= r1; r += 8) {          // eight independent rows in flight per thread
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __ldg(x + (r + k) * xs + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const double d = a[k]; s += d; q = fma(d, d, q); }
    }
    for (; r < r1; ++r) { const double d = __ldg(x + r * xs + c); s += d; q = fma(d, d, q); }
    double* p = partial + (int64_t)blockIdx.x * 2 * C;
    p[c] = s;
    p[C + c] = q;
}

// Streaming variant for even C with 8-byte aligned rows: a block owns a contiguous range of rows and sweeps them front to
// back (purely sequential DRAM traffic).  thread = ONE column pair (float2 loads; the block is as wide as the row, rounded
// up to a warp, so every thread carries the same work), RMS_ROWS rows in flight per thread before any arithmetic --
// ~64 B per thread x ~1900 resident threads per SM keeps well over the ~45 KB per SM that HBM3e needs in flight.
// Rows are added in order into one accumulator per column, so the result does not depend on RMS_ROWS.
constexpr int RMS_ROWS = 8;
constexpr int RMS_PAIR_THREADS_MAX = 512;
__global__ void __launch_bounds__(RMS_PAIR_THREADS_MAX) rms_moments_rows_kernel(const float* __restrict__ x, int64_t xs, int64_t B, int C,
                                                                                int64_t rows_per_block, double* __restrict__ partial) {
    const int cp = blockIdx.y * blockDim.x + threadIdx.x;
    if (cp >= (C >> 1)) return;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < B) ? r0 + rows_per_block : B;
    double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
    const float2* col = reinterpret_cast<const float2*>(x) + cp;        // row r is at col + r * (xs / 2)
    const int64_t xs2 = xs >> 1;
    int64_t r = r0;
    for (; r + RMS_ROWS <= r1; r += RMS_ROWS) {
        float2 a[RMS_ROWS];
#pragma unroll
        for (int k = 0; k < RMS_ROWS; ++k) a[k] = __ldcs(col + (r + k) * xs2);     // streamed once: evict-first
#pragma unroll
        for (int k = 0; k < RMS_ROWS; ++k) {
            const double d0 = a[k].x, d1 = a[k].y;
            s0 += d0; q0 = fma(d0, d0, q0);
            s1 += d1; q1 = fma(d1, d1, q1);
        }
    }
    for (; r < r1; ++r) {
        const float2 a = __ldcs(col + r * xs2);
        const double d0 = a.x, d1 = a.y;
        s0 += d0; q0 = fma(d0, d0, q0);
        s1 += d1; q1 = fma(d1, d1, q1);
    }
    double* p = partial + (int64_t)blockIdx.x * 2 * C;
    p[2 * cp] = s0; p[2 * cp + 1] = s1; p[C + 2 * cp] = q0; p[C + 2 * cp + 1] = q1;
}

// moments[1 + i] += sum_p partial[p][i] for i in [0, 2C); moments[0] += rows.  Deterministic: a block owns 32
// columns, its 8 warps each sum a fixed stride-8 subset of the P partial rows (4 independent loads in flight),
// and the 8 sub-sums are combined in a fixed order.
__global__ void __launch_bounds__(256) rms_reduce_kernel(const double* __restrict__ partial, int P, int64_t rows, int C,
                                                         double* __restrict__ moments) {
    __shared__ double sh[8][33];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (i < 2 * C) {
        int p = g;
        for (; p + 24 < P; p += 32) {
            s0 += partial[(int64_t)p * 2 * C + i];
            s1 += partial[(int64_t)(p + 8) * 2 * C + i];
            s2 += partial[(int64_t)(p + 16) * 2 * C + i];
            s3 += partial[(int64_t)(p + 24) * 2 * C + i];
        }
        for (; p < P; p += 8) s0 += partial[(int64_t)p * 2 * C + i];
    }
    sh[g][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (g == 0 && i < 2 * C) {
        double s = sh[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) s += sh[k][lane];
        moments[1 + i] += s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) moments[0] += (double)rows;
}

// running_norm.py:26-34 on the accumulated moments; a single block so that count is read before it is bumped.
__global__ void rms_finalize_kernel(const double* __restrict__ moments, int C, float* __restrict__ running_mean,
                                    float* __restrict__ running_var, float* __restrict__ count) {
    const double n = moments[0];
    if (!(n > 0.0)) return;         // nothing pending (double finalize, empty rollout): leave mean / var / count untouched (uniform exit)
    const float weight = 1.0f / count[0];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double m = moments[1 + c] / n;
        double v = moments[1 + C + c] / n - m * m;          // biased variance (unbiased=False)
        if (v < 0.0) v = 0.0;
        const float mean_b = (float)m, var_b = (float)v;
        running_mean[c] = running_mean[c] * (1.0f - weight) + mean_b * weight;
        running_var[c] = running_var[c] * (1.0f - weight) + var_b * weight;
    }
    __syncthreads();
    if (threadIdx.x == 0) count[0] = count[0] + 1.0f;
}

static int moments_row_blocks() { return 4 * sm_count(); }

}  // namespace phc

using namespace phc;

extern "C" int phc_rms_forward(const float* x, int64_t x_stride, const float* mean, const float* var, float eps, float clip,
                               int64_t B, int C, float* y, int64_t y_stride, phc_stream_t stream) {
    const char* fn = "phc_rms_forward";
    PHC_REQUIRE(B >= 0, PHC_EINVAL, "%s: B < 0", fn);
    PHC_REQUIRE(C >= 1 && C <= 6000, PHC_ESHAPE, "%s: C=%d outside [1,6000]", fn, C);
    if (B == 0) return PHC_OK;
    PHC_REQUIRE(x && mean && var && y, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(x_stride >= C && y_stride >= C, PHC_ESHAPE, "%s: row stride < C", fn);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = 2 * (size_t)C * sizeof(float);
    const int64_t total = B * C;
    if (x_stride == C && y_stride == C && (total & 3) == 0 && aligned16(x) && aligned16(y)) {
        const int64_t n4 = total >> 2;
        int64_t blocks = (n4 + RMS_THREADS - 1) / RMS_THREADS;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        rms_forward_vec_kernel<<<(unsigned)blocks, RMS_THREADS, smem, s>>>(reinterpret_cast<const float4*>(x), mean, var, eps, clip, n4,
                                                                          C, reinterpret_cast<float4*>(y));
    } else {
        int64_t blocks = B;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        rms_forward_kernel<<<(unsigned)blocks, RMS_THREADS, smem, s>>>(x, x_stride, mean, var, eps, clip, B, C, y, y_stride);
    }
    return check_launch(fn);
}

extern "C" int64_t phc_rms_scratch_doubles(int C) { return C < 1 ? 0 : (int64_t)moments_row_blocks() * 2 * C; }

extern "C" int phc_rms_moments(const float* x, int64_t x_stride, int64_t B, int C, double* moments, double* scratch,
                               phc_stream_t stream) {
    const char* fn = "phc_rms_moments";
    PHC_REQUIRE(B >= 0, PHC_EINVAL, "%s: B < 0", fn);
    PHC_REQUIRE(C >= 1, PHC_ESHAPE, "%s: C < 1", fn);
    if (B == 0) return PHC_OK;
    PHC_REQUIRE(x && moments && scratch, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(x_stride >= C, PHC_ESHAPE, "%s: row stride < C", fn);
    cudaStream_t s = (cudaStream_t)stream;
    int64_t row_blocks = (B + 63) / 64;
    if (row_blocks > moments_row_blocks()) row_blocks = moments_row_blocks();
    const int64_t rows_per_block = (B + row_blocks - 1) / row_blocks;
    row_blocks = (B + rows_per_block - 1) / rows_per_block;
    if ((C & 1) == 0 && (x_stride & 1) == 0 && aligned8(x)) {
        const int npairs = C >> 1;
        int threads = ((npairs + 31) / 32) * 32;                  // as wide as the row ...
        int chunks = 1;
        if (threads > RMS_PAIR_THREADS_MAX) {                     // ... or the row split into equal column chunks
            chunks = (npairs + RMS_PAIR_THREADS_MAX - 1) / RMS_PAIR_THREADS_MAX;
            threads = (((npairs + chunks - 1) / chunks + 31) / 32) * 32;
        }
        dim3 grid((unsigned)row_blocks, (unsigned)chunks);
        rms_moments_rows_kernel<<<grid, threads, 0, s>>>(x, x_stride, B, C, rows_per_block, scratch);
    } else {
        dim3 grid((unsigned)row_blocks, (unsigned)((C + RMS_THREADS - 1) / RMS_THREADS));
        rms_moments_kernel<<<grid, RMS_THREADS, 0, s>>>(x, x_stride, B, C, rows_per_block, scratch);
    }
    int rc = check_launch(fn);
    if (rc) return rc;
    rms_reduce_kernel<<<(2 * C + 31) / 32, 256, 0, s>>>(scratch, (int)row_blocks, B, C, moments);
    return check_launch(fn);
}

extern "C" int phc_rms_reduce_partials(const double* partials, int num_partials, int64_t rows, int C, double* moments,
                                       phc_stream_t stream) {
    const char* fn = "phc_rms_reduce_partials";
    PHC_REQUIRE(partials && moments, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(num_partials >= 0 && rows >= 0 && C >= 1, PHC_EINVAL, "%s: bad size", fn);
    rms_reduce_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partials, num_partials, rows, C, moments);
    return check_launch(fn);
}

extern "C" int phc_rms_finalize(const double* moments, int C, float* running_mean, float* running_var, float* count,
                                phc_stream_t stream) {
    const char* fn = "phc_rms_finalize";
    PHC_REQUIRE(moments && running_mean && running_var && count, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE(C >= 1, PHC_ESHAPE, "%s: C < 1", fn);
    rms_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(moments, C, running_mean, running_var, count);
    return check_launch(fn);
}
