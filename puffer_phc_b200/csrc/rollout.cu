// rollout.cu -- the device-resident rollout buffer around c_gae.compute_gae: what the reference's
// Experience.store / sort_training_data (reference puffer_phc/clean_pufferl/structs.py:108-145) and the GAE call site
// (clean_pufferl/core.py:213-259) do with host numpy arrays, a Python sorted() over batch_size tuples and four host<->device copies.
//
// Layout: values / rewards / dones / truncateds fp32 [T, N] and mask u8 [T, N], one row per env step (env id = column).  The
// reference appends, per step, the rows of the non-masked envs in env order until batch_size rows are stored (the last step may be
// stored partially, structs.py:116), so the ARRIVAL index of element (t, e) is
//     rows stored before step t  +  non-masked envs of step t before e,
// and sort_training_data orders the stored rows by (env, step).  Integer work only, bit-exact by construction:
//   phc_rollout_store   ONE launch per env step: copies the five per-env vectors into row t, counts the row's non-masked envs per
//                       32-env group (ballot), leaves each element's rank inside its group (1 byte), adds the row total to c[t]
//                       and to the running row count (integer atomics: deterministic).
//   phc_rollout_sort    four launches per rollout, no host synchronisation:
//     1. scan_rows      block t turns row t's group counts into exclusive prefixes; block 0 also derives R[t] (rows before step t),
//                       the step t* that crosses batch_size, the rows it may still store and the total row count;
//     2. env_counts     thread = env: kept elements per env, exclusive scan inside 1024-env blocks, block sums;
//     3. scan (1 block) exclusive scan of the block sums;
//     4. gather         CTA = 32-env group: rows staged through a shared-memory tile (coalesced reads), then warp = env, lane = step:
//                       the env's kept elements are compacted with a ballot and written CONTIGUOUSLY (sorted order is env-major)
//                       -- dones / values / rewards for phc_gae, the arrival index (the reference's idxs) and the element's
//                       position in the env-major [N, T] flattening (for gathering anything else).
#include "phc_common.cuh"

namespace phc {

constexpr int RO_STORE_THREADS = 256;
constexpr int RO_ENVS_PER_BLOCK = 1024;

// block-wide exclusive scan of one int per thread (blockDim.x a multiple of 32, <= 1024); returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_exclusive_scan(int x, int* s_warp /* [33] */, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads();                      // s_warp may still be read by a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;            // exclusive prefix of the warp sums
        if (lane == 31) s_warp[32] = wi;  // block total
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + incl - x;
}

template <typename FLAG>   // FLAG = float or uint8_t: dtype of the done / truncated vectors handed over
__global__ void __launch_bounds__(RO_STORE_THREADS) rollout_store_kernel(const float* __restrict__ value, const float* __restrict__ reward,
                                                                         const FLAG* __restrict__ done, const FLAG* __restrict__ trunc,
                                                                         const uint8_t* __restrict__ mask, int64_t N, float* values_row,
                                                                         float* rewards_row, float* dones_row, float* truncs_row,
                                                                         uint8_t* mask_row, uint8_t* subrank_row, int32_t* group_counts_row,
                                                                         int32_t* row_count, int64_t* stored) {
    __shared__ int s_cnt[RO_STORE_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * RO_STORE_THREADS + threadIdx.x;
    const bool valid = e < N;
    bool m = false;
    if (valid) {
        m = mask[e] != 0;
        values_row[e] = value[e];
        rewards_row[e] = reward[e];
        dones_row[e] = (float)done[e];
        truncs_row[e] = trunc ? (float)trunc[e] : 0.0f;
        mask_row[e] = m ? 1 : 0;
    }
    const unsigned b = __ballot_sync(FULL, m);
    if (valid) subrank_row[e] = (uint8_t)__popc(b & ((1u << lane) - 1u));
    if (lane == 0) {
        if (valid) group_counts_row[e >> 5] = __popc(b);
        s_cnt[warp] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < RO_STORE_THREADS / 32; ++w) c += s_cnt[w];
        if (c) {
            atomicAdd(row_count, c);
            atomicAdd(reinterpret_cast<unsigned long long*>(stored), (unsigned long long)c);
        }
    }
}

// meta: [0] t* (first step that would pass batch_size; T when none), [1] rows step t* may still store, [2] rows kept in total, [3] T
__global__ void __launch_bounds__(1024) rollout_scan_rows_kernel(const int32_t* __restrict__ group_counts, int32_t* __restrict__ prefix,
                                                                 int64_t groups, const int32_t* __restrict__ row_counts, int T,
                                                                 int64_t batch_size, int64_t* __restrict__ R, int64_t* __restrict__ meta) {
    __shared__ int s_warp[33];
    const int32_t* row = group_counts + (int64_t)blockIdx.x * groups;
    int32_t* prow = prefix + (int64_t)blockIdx.x * groups;
    int carry = 0;
    for (int64_t base = 0; base < groups; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int x = i < groups ? row[i] : 0;
        int total;
        const int ex = block_exclusive_scan(x, s_warp, &total);
        if (i < groups) prow[i] = carry + ex;
        carry += total;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t run = 0, tstar = T, remaining = 0;
        for (int t = 0; t < T; ++t) {
            R[t] = run;
            if (tstar == T && run + row_counts[t] > batch_size) { tstar = t; remaining = batch_size - run; }
            run += row_counts[t];
        }
        R[T] = run;
        meta[0] = tstar; meta[1] = remaining; meta[2] = run < batch_size ? run : batch_size; meta[3] = T;
    }
}

// is element (t, e) one of the stored rows?
__device__ __forceinline__ bool ro_kept(const uint8_t* mask, const uint8_t* subrank, const int32_t* prefix, int64_t N, int64_t group_stride,
                                        int t, int64_t e, int64_t tstar, int64_t remaining) {
    if (t > tstar || mask[(int64_t)t * N + e] == 0) return false;
    if (t < tstar) return true;
    return (int64_t)prefix[(int64_t)t * group_stride + (e >> 5)] + subrank[(int64_t)t * N + e] < remaining;
}

__global__ void __launch_bounds__(RO_ENVS_PER_BLOCK) rollout_env_counts_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ subrank,
                                                                               const int32_t* __restrict__ prefix, int64_t N,
                                                                               int64_t group_stride, const int64_t* __restrict__ meta,
                                                                               int32_t* __restrict__ offs_local, int32_t* __restrict__ block_sums) {
    __shared__ int s_warp[33];
    const int64_t e = (int64_t)blockIdx.x * RO_ENVS_PER_BLOCK + threadIdx.x;
    const int64_t tstar = meta[0], remaining = meta[1];
    const int T = (int)meta[3];
    int k = 0;
    if (e < N) {
        // rows before t* are kept whenever they are not masked (mask holds 0 / 1): plain independent loads, eight in flight -- one
        // dependent round trip per row made this kernel the slowest of the four
        const int tfull = tstar < T ? (int)tstar : T;
        const uint8_t* mp = mask + e;
        int t = 0;
        for (; t + 8 <= tfull; t += 8) {
            int m[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) m[u] = mp[(int64_t)(t + u) * N];
#pragma unroll
            for (int u = 0; u < 8; ++u) k += m[u];
        }
        for (; t < tfull; ++t) k += mp[(int64_t)t * N];
        if (tstar < T) k += ro_kept(mask, subrank, prefix, N, group_stride, (int)tstar, e, tstar, remaining) ? 1 : 0;
    }
    int total;
    const int ex = block_exclusive_scan(k, s_warp, &total);
    if (e < N) offs_local[e] = ex;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) rollout_scan_blocks_kernel(const int32_t* __restrict__ block_sums, int nblocks, int64_t* __restrict__ base) {
    __shared__ int s_warp[33];
    int64_t carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int x = i < nblocks ? block_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(x, s_warp, &total);
        if (i < nblocks) base[i] = carry + ex;
        carry += total;
    }
}

// CTA = one 32-env group, 8 warps.  Rows are read the way they are stored (a warp reads one row's 32 consecutive envs: full sectors)
// into a shared-memory tile, RO_TCH rows at a time; then every warp takes four of the envs with lane = step, compacts the kept elements
// with a ballot and writes them contiguously (an env's kept elements are adjacent in the sorted order).  Reading with lane = step
// straight from global memory cost one sector request per lane and row: 57 us at 65536 x 32.
constexpr int RO_TCH = 64;
__global__ void __launch_bounds__(256) rollout_gather_kernel(const float* __restrict__ dones, const float* __restrict__ values,
                                                             const float* __restrict__ rewards, const uint8_t* __restrict__ mask,
                                                             const uint8_t* __restrict__ subrank, const int32_t* __restrict__ prefix, int64_t N,
                                                             int64_t group_stride, const int64_t* __restrict__ R, const int64_t* __restrict__ meta,
                                                             const int32_t* __restrict__ offs_local, const int64_t* __restrict__ base,
                                                             float* __restrict__ s_dones, float* __restrict__ s_values, float* __restrict__ s_rewards,
                                                             int64_t* __restrict__ idxs, int64_t* __restrict__ pos_em) {
    __shared__ float t_d[RO_TCH][33], t_v[RO_TCH][33], t_r[RO_TCH][33];      // [row][env]: lane = row reads are conflict-free (pitch 33)
    __shared__ uint8_t t_m[RO_TCH][36], t_s[RO_TCH][36];                      // pitch 36 bytes = 9 words
    __shared__ int32_t t_pref[RO_TCH];
    __shared__ int64_t t_R[RO_TCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = blockIdx.x, e0 = g * 32;
    const int64_t tstar = meta[0], remaining = meta[1];
    const int T = (int)meta[3];
    const int tend = tstar < T ? (int)tstar + 1 : T;
    int64_t out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t e = e0 + warp * 4 + i;
        out[i] = e < N ? base[e / RO_ENVS_PER_BLOCK] + offs_local[e] : 0;
    }
    for (int t0 = 0; t0 < tend; t0 += RO_TCH) {
        const int rows = tend - t0 < RO_TCH ? tend - t0 : RO_TCH;
        __syncthreads();                               // the previous chunk has been consumed
        {   // every load of the warp's (up to) eight rows is requested before the first shared-memory store: one memory round trip per
            // chunk instead of one per row (the kernel is latency-bound: 2 048 short-lived CTAs)
            const int64_t e = e0 + lane;
            const bool ok = e < N;
            float rd[RO_TCH / 8], rv[RO_TCH / 8], rr[RO_TCH / 8];
            uint8_t rm[RO_TCH / 8], rs[RO_TCH / 8];
            int32_t rp[RO_TCH / 8];
            int64_t rR[RO_TCH / 8];
#pragma unroll
            for (int u = 0; u < RO_TCH / 8; ++u) {
                const int tr = warp + 8 * u, t = t0 + tr;
                rd[u] = rv[u] = rr[u] = 0.0f; rm[u] = rs[u] = 0; rp[u] = 0; rR[u] = 0;
                if (tr < rows) {
                    const int64_t src = (int64_t)t * N + e;
                    if (ok) { rd[u] = dones[src]; rv[u] = values[src]; rr[u] = rewards[src]; rm[u] = mask[src]; rs[u] = subrank[src]; }
                    if (lane == 0) { rp[u] = prefix[(int64_t)t * group_stride + g]; rR[u] = R[t]; }
                }
            }
#pragma unroll
            for (int u = 0; u < RO_TCH / 8; ++u) {
                const int tr = warp + 8 * u;
                if (tr < rows) {
                    t_d[tr][lane] = rd[u]; t_v[tr][lane] = rv[u]; t_r[tr][lane] = rr[u]; t_m[tr][lane] = rm[u]; t_s[tr][lane] = rs[u];
                    if (lane == 0) { t_pref[tr] = rp[u]; t_R[tr] = rR[u]; }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int el = warp * 4 + i;
            const int64_t e = e0 + el;
            if (e >= N) continue;                      // warp-uniform
            for (int tt = 0; tt < rows; tt += 32) {
                const int tr = tt + lane, t = t0 + tr;
                bool k = tr < rows && t_m[tr][el] != 0;
                if (k && t == tstar) k = (int64_t)t_pref[tr] + t_s[tr][el] < remaining;
                const unsigned b = __ballot_sync(FULL, k);
                if (k) {
                    const int64_t p = out[i] + __popc(b & ((1u << lane) - 1u));
                    s_dones[p] = t_d[tr][el];
                    s_values[p] = t_v[tr][el];
                    s_rewards[p] = t_r[tr][el];
                    idxs[p] = t_R[tr] + t_pref[tr] + t_s[tr][el];                               // structs.py:116-131: arrival row
                    pos_em[p] = e * (int64_t)T + t;
                }
                out[i] += __popc(b);
            }
        }
    }
}

}  // namespace phc

using namespace phc;

extern "C" int phc_rollout_store(const float* value, const float* reward, const void* done, const void* trunc, int flags_are_float,
                                 const uint8_t* mask, int64_t N, float* values_row, float* rewards_row, float* dones_row,
                                 float* truncateds_row, uint8_t* mask_row, uint8_t* subrank_row, int32_t* group_counts_row,
                                 int32_t* row_count, int64_t* stored, phc_stream_t stream) {
    const char* fn = "phc_rollout_store";
    PHC_REQUIRE(N >= 0, PHC_EINVAL, "%s: N < 0", fn);
    if (N == 0) return PHC_OK;
    PHC_REQUIRE(value && reward && done && mask && values_row && rewards_row && dones_row && truncateds_row && mask_row && subrank_row &&
                    group_counts_row && row_count && stored, PHC_EINVAL, "%s: NULL pointer", fn);
    const unsigned grid = (unsigned)((N + RO_STORE_THREADS - 1) / RO_STORE_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    if (flags_are_float)
        rollout_store_kernel<float><<<grid, RO_STORE_THREADS, 0, s>>>(value, reward, (const float*)done, (const float*)trunc, mask, N, values_row,
                                                                      rewards_row, dones_row, truncateds_row, mask_row, subrank_row,
                                                                      group_counts_row, row_count, stored);
    else
        rollout_store_kernel<uint8_t><<<grid, RO_STORE_THREADS, 0, s>>>(value, reward, (const uint8_t*)done, (const uint8_t*)trunc, mask, N,
                                                                        values_row, rewards_row, dones_row, truncateds_row, mask_row, subrank_row,
                                                                        group_counts_row, row_count, stored);
    return check_launch(fn);
}

extern "C" int64_t phc_rollout_scratch_bytes(int64_t N, int T) {
    if (N < 0 || T < 0) return 0;
    const int64_t nblk = (N + RO_ENVS_PER_BLOCK - 1) / RO_ENVS_PER_BLOCK;
    const int64_t groups = (N + 31) / 32;
    // R [T + 1] int64 | base [nblk] int64 | prefix [T][groups] int32 | offs_local [N] int32 | block_sums [nblk] int32
    return (int64_t)(T + 1) * 8 + nblk * 8 + ((int64_t)T * groups + N + nblk + 2) * 4;
}

extern "C" int phc_rollout_sort(const float* dones, const float* values, const float* rewards, const uint8_t* mask, const uint8_t* subrank,
                                const int32_t* group_counts, const int32_t* row_counts, int64_t N, int T, int64_t batch_size, void* scratch,
                                int64_t* meta, float* sorted_dones, float* sorted_values, float* sorted_rewards, int64_t* idxs,
                                int64_t* pos_em, phc_stream_t stream) {
    const char* fn = "phc_rollout_sort";
    PHC_REQUIRE(N >= 1 && T >= 1 && batch_size >= 1, PHC_EINVAL, "%s: N, T and batch_size must be positive", fn);
    PHC_REQUIRE(dones && values && rewards && mask && subrank && group_counts && row_counts && scratch && meta && sorted_dones &&
                    sorted_values && sorted_rewards && idxs && pos_em, PHC_EINVAL, "%s: NULL pointer", fn);
    PHC_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, PHC_EALIGN, "%s: scratch must be 8-byte aligned", fn);
    const int64_t groups = (N + 31) / 32;
    const int64_t nblk = (N + RO_ENVS_PER_BLOCK - 1) / RO_ENVS_PER_BLOCK;
    PHC_REQUIRE(nblk <= (1 << 22), PHC_EUNSUPPORTED, "%s: N too large", fn);
    int64_t* R = static_cast<int64_t*>(scratch);
    int64_t* base = R + (T + 1);
    int32_t* prefix = reinterpret_cast<int32_t*>(base + nblk);
    int32_t* offs_local = prefix + (int64_t)T * groups;
    int32_t* block_sums = offs_local + N;
    cudaStream_t s = (cudaStream_t)stream;
    rollout_scan_rows_kernel<<<T, 1024, 0, s>>>(group_counts, prefix, groups, row_counts, T, batch_size, R, meta);
    rollout_env_counts_kernel<<<(unsigned)nblk, RO_ENVS_PER_BLOCK, 0, s>>>(mask, subrank, prefix, N, groups, meta, offs_local, block_sums);
    rollout_scan_blocks_kernel<<<1, 1024, 0, s>>>(block_sums, (int)nblk, base);
    rollout_gather_kernel<<<(unsigned)groups, 256, 0, s>>>(dones, values, rewards, mask, subrank, prefix, N, groups, R, meta,
                                                                  offs_local, base, sorted_dones, sorted_values, sorted_rewards, idxs, pos_em);
    return check_launch(fn);
}
