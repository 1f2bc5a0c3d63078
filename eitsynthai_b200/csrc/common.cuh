// Shared helpers for the libeitb200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/eitb200.h"

#define EITB_NUM_SMS 148

// Per-launch CUDA-event timing (api_misc.cu); both calls are no-ops unless eitb_profile_enable(1).
void eitb_prof_begin(const char* kernel_name, cudaStream_t stream);
void eitb_prof_end();

#define EITB_CHECK_LAUNCH()                                    \
    do {                                                       \
        eitb_prof_end();                                       \
        if (cudaGetLastError() != cudaSuccess) return EITB_ERR_LAUNCH; \
    } while (0)

static inline int eitb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid for a grid-stride kernel: enough CTAs to fill the chip, capped by the work
static inline int eitb_grid(long long work_items, int threads, int ctas_per_sm) {
    long long need = (work_items + threads - 1) / threads;
    long long cap = (long long)EITB_NUM_SMS * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// x-extent of a (pixel blocks, image) grid: enough CTAs per image to fill the chip across B images
static inline int eitb_grid_per_image(long long items_per_image, int threads, int B) {
    long long need = (items_per_image + threads - 1) / threads;
    long long cap = ((long long)EITB_NUM_SMS * 8) / (B > 0 ? B : 1);    // floor: never spill into a second, mostly empty wave
    if (cap < 1) cap = 1;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// CTAs that are resident at once for this kernel (SMs x occupancy): the grid of a grid-stride
// streaming kernel, so that it runs as exactly one full wave.
template <typename K>
static inline int eitb_resident_ctas(K kernel, int threads, size_t smem) {
    // one query per (kernel, smem bucket) and thread; the driver call costs tens of microseconds
    static thread_local K last_kernel = nullptr;
    static thread_local size_t last_smem = ~(size_t)0;
    static thread_local int last = 0;
    if (last_kernel == kernel && last_smem == smem && last > 0) return last;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    last_kernel = kernel; last_smem = smem; last = EITB_NUM_SMS * per_sm;
    return last;
}

__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_uint2(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_int4(int4* p, int4 v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_uint2(uint2* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// u8/255 rounded once to the output type, the way torch does x.to(dtype) / 255
// (division evaluated in fp32, then rounded to the storage type).
template <typename T> __device__ __forceinline__ T unit_from_u8(int u);
template <> __device__ __forceinline__ float unit_from_u8<float>(int u) { return __fdiv_rn((float)u, 255.0f); }
template <> __device__ __forceinline__ __half unit_from_u8<__half>(int u) { return __float2half_rn(__fdiv_rn((float)u, 255.0f)); }
template <> __device__ __forceinline__ __nv_bfloat16 unit_from_u8<__nv_bfloat16>(int u) { return __float2bfloat16_rn(__fdiv_rn((float)u, 255.0f)); }
