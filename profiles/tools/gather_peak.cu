// gather_peak.cu -- what does THIS access pattern get out of HBM3e?  A kernel with the fused step's memory traffic and none of its
// arithmetic: per env one sequential 1248-byte record, two RANDOM 1248-byte frame records + two random 192-byte pair-table rows out of
// a 3 GB table, and 7472 bytes of sequential output (two 3736-byte rows).  Its GB/s is the ceiling for phc_step_fused's DRAM traffic
// (profiles/README.md); the streaming-copy peak of MEASURED_PEAKS.json is not reachable with 1.2 KB gathers and a 57 % write share.
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o _build/libgather_peak.so gather_peak.cu
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) gather_mix_kernel(const float4* __restrict__ sim, const float4* __restrict__ frames,
                                                         const float4* __restrict__ aux, const int64_t* __restrict__ f0,
                                                         const int64_t* __restrict__ f1, int64_t N, float4* __restrict__ out, int frames_per_env) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < N; e += nwarps) {
        const int64_t a = f0[e], b = f1[e];
        float4 r[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {                       // 78 x 16 bytes per record
            const int i = lane + 32 * k;
            r[k] = i < 78 ? __ldg(sim + e * 78 + i) : make_float4(0, 0, 0, 0);
            r[3 + k] = i < 78 ? __ldg(frames + a * 78 + i) : make_float4(0, 0, 0, 0);
            r[6 + k] = (frames_per_env > 1 && i < 78) ? __ldg(frames + b * 78 + i) : make_float4(0, 0, 0, 0);
        }
        const float4 x = lane < 12 ? __ldg(aux + a * 12 + lane) : (lane < 24 ? __ldg(aux + b * 12 + (lane - 12)) : make_float4(0, 0, 0, 0));
        float4 s = x;
#pragma unroll
        for (int k = 0; k < 9; ++k) { s.x += r[k].x; s.y += r[k].y; s.z += r[k].z; s.w += r[k].w; }
        float4* o = out + e * 467;                          // 7472 bytes = 467 x 16
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            const int i = lane + 32 * k;
            if (i < 467) o[i] = (k < 9) ? r[k] : s;
        }
    }
}

extern "C" int gather_mix(const void* sim, const void* frames, const void* aux, const int64_t* f0, const int64_t* f1, int64_t N, void* out,
                          int frames_per_env, int blocks, void* stream) {
    gather_mix_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)sim, (const float4*)frames, (const float4*)aux, f0, f1, N,
                                                               (float4*)out, frames_per_env);
    return (int)cudaGetLastError();
}
