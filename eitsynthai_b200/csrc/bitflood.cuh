// Frame flood on bit images: which pixels of a set are 4-connected to the image frame through the set?
//
// K2 (body mask) needs it for "outside every external contour" (utils.py:572-582) and K7 (label clean-up) to tell
// external contours from nested ones (cv2.RETR_EXTERNAL, utils.py:792) for each of the three target colours.  Round 1
// answered it with a full connected-component labelling of the background (int32 label per pixel, union-find: ~1 M warp
// instructions per image); here the image is one bit per pixel and a whole row is handled at once:
//
//   * a row of W <= 1024 pixels is W/32 words, one per lane of a warp;
//   * "extend the seeds s along the runs of the allowed set a" is two multi-word additions: in a + s every run that
//     holds a seed carries from the seed to the run's end, so ((a ^ (a + s)) & a) | s fills upwards, and the same on
//     the bit-reversed row fills downwards; the carries between the words come from two ballots (carry look-ahead);
//   * rows propagate into their neighbours by sweeping down and up; NW warps sweep NW bands of rows at the same time
//     and repeat until no band changed.
//
// One CTA per (image, job); `allowed` and `reach` live in shared memory (2 x H x W/8 bytes).
#pragma once
#include "common.cuh"

namespace eitb_flood {

constexpr int kWarps = 8;

enum Src { SRC_U8_NE = 0, SRC_BITS_ZERO = 1, SRC_U8_EQ = 2 };

// multi-word a + s over the lanes [0, wpr); returns this lane's word of the sum
__device__ __forceinline__ uint32_t mw_add(uint32_t a, uint32_t s, int lane, unsigned lane_mask) {
    const uint32_t sum = a + s;
    const unsigned G = __ballot_sync(0xffffffffu, sum < a) & lane_mask;            // word generates a carry
    const unsigned P = __ballot_sync(0xffffffffu, sum == 0xffffffffu) & lane_mask; // word passes a carry on
    const unsigned X = G << 1;
    const unsigned C = ((X + P) ^ P ^ X) | X;                                      // carry INTO each word
    return sum + ((C >> lane) & 1u);
}

// extend the seed bits s (a subset of a) along the runs of a, over a row held one word per lane
__device__ __forceinline__ uint32_t hfill(uint32_t a, uint32_t s, int lane, int wpr, unsigned lane_mask) {
    // a and s are zero in the lanes >= wpr
    const uint32_t up = ((a ^ mw_add(a, s, lane, lane_mask)) & a) | s;
    const int mirror = (wpr - 1 - lane) & 31;                      // the same row, bit-reversed: word order and bit order
    uint32_t ar = __brev(__shfl_sync(0xffffffffu, a, mirror)), sr = __brev(__shfl_sync(0xffffffffu, s, mirror));
    if (lane >= wpr) { ar = 0u; sr = 0u; }
    const uint32_t dr = ((ar ^ mw_add(ar, sr, lane, lane_mask)) & ar) | sr;
    const uint32_t down = __brev(__shfl_sync(0xffffffffu, dr, mirror));
    return lane < wpr ? (up | down) : 0u;
}

// 32 pixels (bytes) -> one word: bit i = (byte i != t)
__device__ __forceinline__ uint32_t ne_word(const uint8_t* p, int t) {
    const int4 v0 = *reinterpret_cast<const int4*>(p), v1 = *reinterpret_cast<const int4*>(p + 16);
    const uint32_t w[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w,
                           (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
    const uint32_t tt = (uint32_t)t * 0x01010101u;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        out |= (((__vcmpne4(w[k], tt) & 0x80808080u) * 0x00204081u) >> 28) << (4 * k);
    return out;
}

// grid (jobs_per_image, B).  SRC_U8_NE: src u8 [B,H,W], the set of job j is {src != targets[j]}; SRC_U8_EQ: {src == targets[j]}.
// SRC_BITS_ZERO: src bit image [B, H*W/32] words, one job, the set is the zero bits.
// reach_out [B, out_jobs, H*W/32], this launch fills the jobs [job_off, job_off + gridDim.x): bit = pixel is in the
// set and frame-connected.
template <int SRC>
__global__ void __launch_bounds__(kWarps * 32)
frame_flood_kernel(const void* __restrict__ src, int H, int W, int t0, int t1, int t2, const int* __restrict__ skip,
                   uint32_t* __restrict__ reach_out, int out_jobs, int job_off) {
    extern __shared__ uint32_t fsm[];
    const int wpr = W >> 5, job = blockIdx.x, b = blockIdx.y, jobs = gridDim.x;
    if (skip && skip[b * jobs + job]) return;
    uint32_t* allowed = fsm;
    uint32_t* reach = fsm + (size_t)H * wpr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lane_mask = wpr >= 32 ? 0xffffffffu : ((1u << wpr) - 1u);
    const int t = job == 0 ? t0 : job == 1 ? t1 : t2;
    const int band = (H + kWarps - 1) / kWarps;
    const int y_lo = warp * band, y_hi = min(y_lo + band, H);

    // ---- the allowed set, and the frame seeds extended along their rows
    for (int y = y_lo; y < y_hi; ++y) {
        uint32_t a = 0;
        if (lane < wpr) {
            if (SRC == SRC_U8_NE) a = ne_word(reinterpret_cast<const uint8_t*>(src) + ((size_t)b * H + y) * W + lane * 32, t);
            else if (SRC == SRC_U8_EQ) a = ~ne_word(reinterpret_cast<const uint8_t*>(src) + ((size_t)b * H + y) * W + lane * 32, t);
            else a = ~reinterpret_cast<const uint32_t*>(src)[((size_t)b * H + y) * wpr + lane];
        }
        uint32_t s = (y == 0 || y == H - 1) ? a : 0u;
        if (lane == 0) s |= a & 1u;
        if (lane == wpr - 1) s |= a & 0x80000000u;
        const uint32_t r = hfill(a, s, lane, wpr, lane_mask);
        if (lane < wpr) { allowed[y * wpr + lane] = a; reach[y * wpr + lane] = r; }
    }
    __syncthreads();
    // ---- sweep the bands down and up until nothing changes anywhere
    for (;;) {
        bool changed = false;
        uint32_t prev = (y_lo > 0 && lane < wpr) ? reach[(y_lo - 1) * wpr + lane] : 0u;     // row above the band (neighbour's)
        for (int y = y_lo; y < y_hi; ++y) {
            const uint32_t a = lane < wpr ? allowed[y * wpr + lane] : 0u;
            const uint32_t cur = lane < wpr ? reach[y * wpr + lane] : 0u;
            const uint32_t s = cur | (prev & a);
            uint32_t r = cur;
            if (__any_sync(0xffffffffu, s != cur)) {               // something new enters the row: extend it
                r = hfill(a, s, lane, wpr, lane_mask);
                if (lane < wpr) reach[y * wpr + lane] = r;
                changed = true;
            }
            prev = r;
        }
        prev = (y_hi < H && lane < wpr) ? reach[y_hi * wpr + lane] : 0u;                    // row below the band
        for (int y = y_hi - 1; y >= y_lo; --y) {
            const uint32_t a = lane < wpr ? allowed[y * wpr + lane] : 0u;
            const uint32_t cur = lane < wpr ? reach[y * wpr + lane] : 0u;
            const uint32_t s = cur | (prev & a);
            uint32_t r = cur;
            if (__any_sync(0xffffffffu, s != cur)) {
                r = hfill(a, s, lane, wpr, lane_mask);
                if (lane < wpr) reach[y * wpr + lane] = r;
                changed = true;
            }
            prev = r;
        }
        if (!__syncthreads_or(changed ? 1 : 0)) break;
    }
    uint32_t* out = reach_out + ((size_t)b * out_jobs + job_off + job) * H * wpr;
    for (int i = threadIdx.x; i < H * wpr; i += kWarps * 32) out[i] = reach[i];
}

inline bool flood_supported(int H, int W) {
    return (W & 31) == 0 && W <= 1024 && (size_t)H * (W >> 5) * 8 <= 200 * 1024;
}

// src: see the kernel; reach_out [B, jobs, H*W/32] words
template <int SRC>
int frame_flood(const void* src, int B, int H, int W, int jobs, int t0, int t1, int t2, const int* skip, uint32_t* reach_out,
                cudaStream_t s, int out_jobs = 0, int job_off = 0) {
    if (!flood_supported(H, W) || jobs < 1 || jobs > 3 || B > 65535) return EITB_ERR_UNSUPPORTED;
    const size_t smem = (size_t)H * (W >> 5) * 8;
    if (cudaFuncSetAttribute(frame_flood_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    eitb_prof_begin("frame_flood_kernel", s);
    frame_flood_kernel<SRC><<<dim3(jobs, B), kWarps * 32, smem, s>>>(src, H, W, t0, t1, t2, skip, reach_out, out_jobs > 0 ? out_jobs : jobs,
                                                                     job_off);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

}  // namespace eitb_flood
