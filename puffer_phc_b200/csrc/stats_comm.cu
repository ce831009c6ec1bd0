// stats_comm.cu -- the ONE exchange step of the hot path as ONE kernel over NVLink peer memory:
//     fold the rank's per-CTA partial sums  ->  all-reduce [n, sum x, sum x^2 | episode metrics] across the GPUs of the box
//     ->  RunningNorm's running-average update (reference puffer_phc/policies/running_norm.py:23-34)
// instead of  phc_stats_reduce + ncclAllReduce (15 KB, latency-bound) + phc_rms_finalize + a memset  on the compute stream.
//
// Every rank owns an exchange buffer that all ranks of the node can address (torch symmetric memory: the driver maps the peers'
// allocations into this process; NVSwitch routes the stores).  The kernel is a push all-gather + local ordered sum:
//   block b (16 observation columns: lanes 0..15 carry sum x, lanes 16..31 sum x^2, lane 32 the row count; the last block carries
//   the episode metrics) folds its columns from the rank's partial slots, STORES the 33 doubles into slot [parity][my rank][b] of
//   EVERY rank's buffer (plain st.global to peer addresses), fences (system scope) and publishes flag[parity][my rank][b] = epoch
//   on every rank; then it waits until its own buffer holds this epoch's flags of all ranks for block b and adds the contributions
//   IN RANK ORDER -- every rank performs the identical sequence of fp64 additions, so mean / var / count / metric sums are
//   bit-identical on all ranks by construction -- and finalises its 16 columns of running_mean / running_var.
// No block waits for another block of the same GPU, only for the matching block of the peers (60 blocks, all co-resident), so the
// kernel cannot deadlock against itself; a peer that never arrives trips a 20 s watchdog trap instead of hanging the GPU.
// Buffers are double-buffered by epoch parity: a rank can only be one exchange ahead of its slowest peer (it needs every peer's
// flags of epoch e to finish e, and a peer publishes e only after its own kernel of e - 1 has completed).
#include "phc_common.cuh"

namespace phc {

constexpr int SC_COLS = 16;                         // observation columns per block
constexpr int SC_MSG = 33;                          // doubles per block message: 16 sum x | 16 sum x^2 | n   (metrics block: 16 metrics | - | n)

__host__ __device__ inline int sc_blocks(int C) { return (C + SC_COLS - 1) / SC_COLS + 1; }        // + the metrics block
__host__ __device__ inline size_t sc_data_doubles(int world, int C) { return (size_t)2 * world * sc_blocks(C) * SC_MSG; }

struct SCArgs {
    double* partial; int P; int C; double rows; double* row_adjust;
    double* mpartial; int MP;
    double* stats;
    phc_stats_comm comm;
    float* running_mean; float* running_var; float* count;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) stats_allreduce_finalize_kernel(const SCArgs a) {
    __shared__ double sh[8][SC_MSG];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int C = a.C, W = a.comm.world, me = a.comm.rank;
    const int nblk = sc_blocks(C), b = blockIdx.x;
    const bool metrics_block = b == nblk - 1;
    const float cnt = a.count ? a.count[0] : 1.0f;                 // read before any block can bump it (see the ticket below)
    // ---- 1. fold this block's columns from the rank's partial slots (fixed order: 8 warps x stride-8 rows, then 8 sub-sums) ------
    double s = 0.0;
    if (!metrics_block) {
        const int c = b * SC_COLS + (lane & 15);
        const int i = (lane < 16) ? c : C + c;                     // index into a partial slot: [sum x (C) | sum x^2 (C)]
        if (c < C && a.partial) {               // four independent loads in flight per thread, then their clears (fixed order)
            const size_t ps = (size_t)2 * C;
            int p = g;
            for (; p + 24 < a.P; p += 32) {
                double* q = a.partial + (size_t)p * ps + i;
                const double x0 = q[0], x1 = q[8 * ps], x2 = q[16 * ps], x3 = q[24 * ps];
                q[0] = 0.0; q[8 * ps] = 0.0; q[16 * ps] = 0.0; q[24 * ps] = 0.0;
                s += (x0 + x1) + (x2 + x3);
            }
            for (; p < a.P; p += 8) {
                double* q = a.partial + (size_t)p * ps + i;
                s += *q;
                *q = 0.0;
            }
        }
    } else if (a.mpartial && lane < PHC_NUM_METRICS && g == 0) {
        for (int p = 0; p < a.MP; ++p) {
            double* q = a.mpartial + (size_t)p * PHC_NUM_METRICS + lane;
            s += *q;
            *q = 0.0;
        }
    }
    sh[g][lane] = s;
    __syncthreads();
    if (g != 0) return;                                            // one warp does the exchange and the finalisation
    double v = sh[0][lane];
    if (!metrics_block) {
#pragma unroll
        for (int k = 1; k < 8; ++k) v += sh[k][lane];
    }
    double n_mine = a.rows + (a.row_adjust ? a.row_adjust[0] : 0.0);
    // ---- 2. push to every rank (own buffer included), fence, publish -------------------------------------------------------------
    const unsigned long long epoch = a.comm.epoch;
    const int par = (int)(epoch & 1ull);
    const size_t data_doubles = sc_data_doubles(W, C);
    const size_t slot = ((size_t)(par * W + me) * nblk + b) * SC_MSG;
    for (int r = 0; r < W; ++r) {
        double* dst = reinterpret_cast<double*>(a.comm.peer_bufs[r]) + slot;
        dst[lane] = v;
        if (lane == 0) dst[32] = n_mine;
    }
    __threadfence_system();
    __syncwarp();
    if (lane < W) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(a.comm.peer_bufs[lane]) + data_doubles);
        st_release_sys(flags + (size_t)(par * W + me) * nblk + b, epoch);
    }
    // ---- 3. wait for every rank's message of this block, add in rank order -----------------------------------------------------
    const double* mine = reinterpret_cast<const double*>(a.comm.peer_bufs[me]);
    const unsigned long long* myflags = reinterpret_cast<const unsigned long long*>(mine + data_doubles);
    if (lane < W) {
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(myflags + (size_t)(par * W + lane) * nblk + b) != epoch) {
            if (globaltimer_ns() - t0 > 20000000000ull) __trap();                 // a peer never arrived
            __nanosleep(200);
        }
    }
    __syncwarp();
    double tot = 0.0, n = 0.0;
    for (int r = 0; r < W; ++r) {
        const double* src = mine + ((size_t)(par * W + r) * nblk + b) * SC_MSG;
        tot += src[lane];
        n += src[32];
    }
    // ---- 4. results ------------------------------------------------------------------------------------------------------------
    if (metrics_block) {
        if (lane < PHC_NUM_METRICS) a.stats[1 + 2 * C + lane] += tot;            // global sums (all ranks hold the same values)
    } else {
        const double sxx = __shfl_down_sync(FULL, tot, 16);
        const int c = b * SC_COLS + lane;
        if (lane < 16 && c < C && n > 0.0 && a.running_mean) {                     // running_norm.py:26-34
            const double m = tot / n;
            double var = sxx / n - m * m;                                          // biased variance (unbiased=False)
            if (var < 0.0) var = 0.0;
            const float weight = 1.0f / cnt;
            const float mean_b = (float)m, var_b = (float)var;
            a.running_mean[c] = a.running_mean[c] * (1.0f - weight) + mean_b * weight;
            a.running_var[c] = a.running_var[c] * (1.0f - weight) + var_b * weight;
        }
    }
    // ---- 5. count += 1 once every block has read it: the last block to finish does it and re-arms the ticket -----------------
    if (lane == 0) {
        __threadfence();
        const unsigned t = atomicAdd(a.comm.ticket, 1u);
        if (t == (unsigned)nblk - 1) {                 // every block has read count and row_adjust by now
            if (a.count && n > 0.0) a.count[0] = cnt + 1.0f;
            if (a.row_adjust) a.row_adjust[0] = 0.0;
            *a.comm.ticket = 0u;
        }
    }
}

}  // namespace phc

using namespace phc;

extern "C" int64_t phc_stats_comm_bytes(int world, int C) {
    if (world < 1 || C < 1) return 0;
    return (int64_t)(sc_data_doubles(world, C) * sizeof(double) + (size_t)2 * world * sc_blocks(C) * sizeof(unsigned long long));
}

extern "C" int phc_stats_allreduce_finalize(double* moment_partials, int num_partials, int C, int64_t rows, double* row_adjust,
                                            double* metric_partials, int num_metric_partials, double* stats, const phc_stats_comm* comm,
                                            float* running_mean, float* running_var, float* count, phc_stream_t stream) {
    const char* fn = "phc_stats_allreduce_finalize";
    PHC_REQUIRE(stats && comm, PHC_EINVAL, "%s: stats / comm is NULL", fn);
    PHC_REQUIRE(C >= 1 && num_partials >= 0 && num_metric_partials >= 0 && rows >= 0, PHC_EINVAL, "%s: bad size", fn);
    PHC_REQUIRE(moment_partials || num_partials == 0, PHC_EINVAL, "%s: moment_partials is NULL", fn);
    PHC_REQUIRE(metric_partials || num_metric_partials == 0, PHC_EINVAL, "%s: metric_partials is NULL", fn);
    PHC_REQUIRE(comm->world >= 1 && comm->world <= 32 && comm->rank >= 0 && comm->rank < comm->world, PHC_EINVAL, "%s: bad rank / world", fn);
    PHC_REQUIRE(comm->peer_bufs && comm->ticket && comm->epoch >= 1, PHC_EINVAL, "%s: comm.peer_bufs / ticket NULL or epoch < 1", fn);
    PHC_REQUIRE((running_mean == nullptr) == (running_var == nullptr), PHC_EINVAL, "%s: running_mean and running_var go together", fn);
    SCArgs a{moment_partials, num_partials, C, (double)rows, row_adjust, metric_partials, num_metric_partials, stats, *comm,
             running_mean, running_var, count};
    stats_allreduce_finalize_kernel<<<sc_blocks(C), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch(fn);
}
