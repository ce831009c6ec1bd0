// phc_common.cuh -- shared host/device helpers for libphc_b200 (errors, warp primitives, loads).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/phc_b200.h"
#include "phc_math.cuh"

namespace phc {

constexpr int NB = 24;        // SMPL bodies (reference puffer_phc/body_sets.py:11-36)
constexpr int NDOF = 69;      // 23 joints x 3
constexpr int REC = 13;       // floats per PhysX rigid-body record: pos3 rot4 vel3 angvel3
constexpr int SIM_F = NB * REC;   // 312 floats = 1248 B per env
constexpr int FRAME_F = 312;  // packed reference frame: pos 72 | rot 96 | vel 72 | ang 72
constexpr int OBS_SELF = 358, OBS_TASK = 576, OBS_W = 934;   // humanoid_phc.py:461-467
constexpr unsigned FULL = 0xffffffffu;

// host side -----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

#define PHC_REQUIRE(cond, code, ...) \
    do { if (!(cond)) return ::phc::fail((code), __VA_ARGS__); } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

// device side ---------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ Q4 ld4(const float* p) { return Q4{p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ V3 ldg3(const float* p) { return V3{__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }
__device__ __forceinline__ Q4 ldg4a(const float* p) {   // 16-byte aligned
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return Q4{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ void st4(float* p, Q4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
#endif

}  // namespace phc
