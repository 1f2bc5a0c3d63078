// K12: the two convolution shapes of YOLO11s-seg that are not GEMMs, as direct CUDA-core kernels with the
// same fused epilogue as K11 (folded-BatchNorm bias + SiLU):
//   * the stem  Conv(3 -> 32, k3, s2)  on the 3-channel network input (27 taps per output: no tensor-core shape);
//   * depthwise Conv(C -> C, k3, s1, groups = C) of the class branch of the Segment head and of the
//     attention position encoding (9 taps per channel: bandwidth-bound).
// Reference: the ultralytics modules behind model(...) at kt_service/ai_tools/ai_tools.py:121-122,153.
#include "common.cuh"
#include <cuda.h>        // CUtensorMap and its enums only; the encoder is fetched from the driver at run time
#include <mutex>

namespace {

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// x [N,H,W,3] fp16 (channels-last network input), w [27][COUT] fp32 ((r*3+s)*3+ci major), y [N,Ho,Wo,y_ctot] fp16
template <int COUT>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int N, int H, int W,
                 int Ho, int Wo, int act, __half* __restrict__ y, int y_ctot, int y_coff) {
    __shared__ float sw[27 * COUT];
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const long long total = (long long)N * Ho * Wo;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(o % Wo), oy = (int)((o / Wo) % Ho), n = (int)(o / ((long long)Wo * Ho));
        float in[27];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = 2 * oy - 1 + r;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ix = 2 * ox - 1 + s;
                const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
                const __half* px = x + (((long long)n * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) in[(r * 3 + s) * 3 + c] = ok ? __half2float(px[c]) : 0.f;
            }
        }
        __half* out = y + o * y_ctot + y_coff;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = sb[c8 * 8 + e];
#pragma unroll
            for (int t = 0; t < 27; ++t) {
                const float4 w0 = *reinterpret_cast<const float4*>(&sw[t * COUT + c8 * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&sw[t * COUT + c8 * 8 + 4]);
                acc[0] = fmaf(in[t], w0.x, acc[0]); acc[1] = fmaf(in[t], w0.y, acc[1]);
                acc[2] = fmaf(in[t], w0.z, acc[2]); acc[3] = fmaf(in[t], w0.w, acc[3]);
                acc[4] = fmaf(in[t], w1.x, acc[4]); acc[5] = fmaf(in[t], w1.y, acc[5]);
                acc[6] = fmaf(in[t], w1.z, acc[6]); acc[7] = fmaf(in[t], w1.w, acc[7]);
            }
            int4 v;
            __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = act ? silu_f(acc[2 * e]) : acc[2 * e], b = act ? silu_f(acc[2 * e + 1]) : acc[2 * e + 1];
                h[e] = __floats2half2_rn(a, b);
            }
            *reinterpret_cast<int4*>(out + c8 * 8) = v;
        }
    }
}

// Stem for a gray image replicated to three channels (what K1 and the letterbox produce): the three input
// channels are equal, so the convolution is a 9-tap one with the weights summed over the input channel.
// One thread = one output pixel x 8 output channels; its 72 weights live in registers, four neighbouring threads
// share the pixel's nine inputs (L1 hits) and together store 64 contiguous bytes.
__global__ void __launch_bounds__(256)
stem_gray_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int N, int H, int W,
                 int Ho, int Wo, int act, __half* __restrict__ y, int y_ctot, int y_coff) {
    const int cgp = threadIdx.x & 3;                              // which 8 of the 32 output channels
    float wr[9][8], b8[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e)
            wr[t][e] = w[(t * 3 + 0) * 32 + cgp * 8 + e] + w[(t * 3 + 1) * 32 + cgp * 8 + e] + w[(t * 3 + 2) * 32 + cgp * 8 + e];
#pragma unroll
    for (int e = 0; e < 8; ++e) b8[e] = bias ? bias[cgp * 8 + e] : 0.f;
    const unsigned total = (unsigned)N * Ho * Wo * 4;              // < 2^32: checked by the launcher
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned o = i >> 2;
        const unsigned orow = o / (unsigned)Wo;
        const int ox = (int)(o - orow * Wo);
        const unsigned n = orow / (unsigned)Ho;
        const int oy = (int)(orow - n * Ho);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = b8[e];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = 2 * oy - 1 + r;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ix = 2 * ox - 1 + s;
                const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
                const float v = ok ? __half2float(__ldg(x + ((size_t)(n * H + iy) * W + ix) * 3)) : 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(v, wr[r * 3 + s][e], acc[e]);
            }
        }
        int4 v4;
        __half2* h = reinterpret_cast<__half2*>(&v4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = act ? silu_f(acc[2 * e]) : acc[2 * e], b = act ? silu_f(acc[2 * e + 1]) : acc[2 * e + 1];
            h[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<int4*>(y + (size_t)o * y_ctot + y_coff + cgp * 8) = v4;
    }
}

// Stem on the u8 window image itself (K1's out_u8 / the body-masked classic_norm image): the network input
// x = half(u8 / 255) replicated to three channels never exists in HBM -- the kernel converts through a 256-entry table
// while it stages a (2*TO_H+1) x (2*TO_W+1) input tile in shared memory, and the three equal channels fold into one 9-tap
// filter.  One thread = one output column x 8 output channels x TO_H output rows (its 72 weights stay in registers);
// four neighbouring threads store the 64 contiguous bytes of a pixel, a warp 512.  SiLU through one tanh.approx per
// output: x * sigmoid(x) = h + h * tanh(h), h = x / 2.
constexpr int ST_TO_H = 8, ST_TO_W = 64, ST_IN_H = 2 * ST_TO_H + 1, ST_IN_W = 2 * ST_TO_W + 1, ST_PITCH = ST_IN_W + 3;

__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

__global__ void __launch_bounds__(256, 2)
stem_u8_kernel(const uint8_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int N, int H, int W,
               int Ho, int Wo, int act, __half* __restrict__ y, int y_ctot, int y_coff, int tiles_x, int tiles_y) {
    __shared__ float lut[256];
    __shared__ float tile[ST_IN_H][ST_PITCH];
    const int cgp = threadIdx.x & 3, col = threadIdx.x >> 2;       // 8 of the 32 output channels; output column in the tile
    lut[threadIdx.x] = __half2float(unit_from_u8<__half>((int)threadIdx.x));
    float wr[9][8], b8[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e)
            wr[t][e] = w[(t * 3 + 0) * 32 + cgp * 8 + e] + w[(t * 3 + 1) * 32 + cgp * 8 + e] + w[(t * 3 + 2) * 32 + cgp * 8 + e];
#pragma unroll
    for (int e = 0; e < 8; ++e) b8[e] = bias ? bias[cgp * 8 + e] : 0.f;
    const int total = N * tiles_y * tiles_x;
    // The input tile of one output tile is 17 rows x 129 bytes starting one byte before a 128-byte boundary: every thread
    // fetches (at most) one aligned 16-byte chunk of it -- 17 rows x 10 chunks -- and the fetch for the NEXT tile is in
    // flight while this tile is computed, so the only global-memory latency a CTA ever waits for is the first one.
    const int crow = threadIdx.x / 10, cchunk = threadIdx.x - crow * 10;          // chunk row 0..16, chunk 0..9 (170 threads fetch)
    const bool fetcher = threadIdx.x < ST_IN_H * 10;
    const bool vec_ok = (W & 15) == 0 && !(reinterpret_cast<uintptr_t>(x) & 15);
    auto fetch = [&](int tix, uint4& v, bool& ok) {
        ok = false;
        v = make_uint4(0u, 0u, 0u, 0u);
        if (!fetcher || tix >= total || !vec_ok) return;
        const int tx = tix % tiles_x, ty = (tix / tiles_x) % tiles_y, n = tix / (tiles_x * tiles_y);
        const int iy = 2 * ty * ST_TO_H - 1 + crow, xs = 2 * tx * ST_TO_W - 16 + cchunk * 16;   // first pixel of the chunk
        if (iy < 0 || iy >= H || xs < 0 || xs + 16 > W) return;                                 // stays zero: the padding
        v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)n * H + iy) * W + xs));
        ok = true;
    };
    uint4 pre;
    bool pre_ok;
    fetch(blockIdx.x, pre, pre_ok);
    for (int tix = blockIdx.x; tix < total; tix += gridDim.x) {
        const int tx = tix % tiles_x, ty = (tix / tiles_x) % tiles_y, n = tix / (tiles_x * tiles_y);
        const int ox0 = tx * ST_TO_W, oy0 = ty * ST_TO_H;
        const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
        __syncthreads();                                           // the previous tile is consumed (and the table is written)
        if (vec_ok) {
            if (fetcher) {
                const uint32_t wv[4] = {pre.x, pre.y, pre.z, pre.w};
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int c = cchunk * 16 + k - 15;            // tile column of this byte
                    if (c >= 0 && c < ST_IN_W) tile[crow][c] = pre_ok ? lut[(wv[k >> 2] >> (8 * (k & 3))) & 0xffu] : 0.f;
                }
            }
        } else {
            const uint8_t* img = x + (size_t)n * H * W;
            for (int i = threadIdx.x; i < ST_IN_H * ST_IN_W; i += 256) {
                const int r = i / ST_IN_W, c = i - r * ST_IN_W;
                const int iy = iy0 + r, ix = ix0 + c;
                tile[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? lut[__ldg(img + (size_t)iy * W + ix)] : 0.f;
            }
        }
        __syncthreads();
        fetch(tix + gridDim.x, pre, pre_ok);                       // in flight during the arithmetic below
        const int ox = ox0 + col;
        if (ox >= Wo) continue;
        float r0[3], r1[3], r2[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) r0[s] = tile[0][2 * col + s];
#pragma unroll
        for (int r = 0; r < ST_TO_H; ++r) {
            const int oy = oy0 + r;
#pragma unroll
            for (int s = 0; s < 3; ++s) { r1[s] = tile[2 * r + 1][2 * col + s]; r2[s] = tile[2 * r + 2][2 * col + s]; }
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = b8[e];
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    acc[e] = fmaf(r0[s], wr[s][e], acc[e]);
                    acc[e] = fmaf(r1[s], wr[3 + s][e], acc[e]);
                    acc[e] = fmaf(r2[s], wr[6 + s][e], acc[e]);
                }
            if (oy < Ho) {
                int4 v4;
                __half2* h = reinterpret_cast<__half2*>(&v4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a = act ? silu_tanh(acc[2 * e]) : acc[2 * e], b = act ? silu_tanh(acc[2 * e + 1]) : acc[2 * e + 1];
                    h[e] = __floats2half2_rn(a, b);
                }
                *reinterpret_cast<int4*>(y + ((size_t)(n * Ho + oy) * Wo + ox) * y_ctot + y_coff + cgp * 8) = v4;
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) r0[s] = r2[s];
        }
    }
}

// The same stem with the 9 x 32 contraction on the tensor cores (mma.sync m16n8k16, fp32 accumulate): the CUDA-core kernel
// above spends 9 FMAs per output and runs at ~45 % of the issue rate (0.64 ms per 320 slices for 1.4 GB of traffic).  The
// tile is staged as fp16 (the u8 -> fp16 / 255 table), an M tile is 16 consecutive output pixels of a row, K = the 9 taps
// padded to 16, N = 4 x 8 output channels.  The channel <-> column assignment is chosen so that a thread ends up with 8
// CONSECUTIVE channels of its two pixels (column c of N tile j <-> channel 8 (c / 2) + 2 j + (c & 1)): one 16-byte store
// per pixel, a warp writes 1 KB contiguous.  Weights (sums over the three equal input channels) enter as fp16 hi + lo, so
// nothing is lost against the fp32 sums of the scalar kernel.  Epilogue: bias, SiLU as h + h tanh(h) with one
// tanh.approx.f16x2 per two outputs (as in K11).
constexpr int ST16_PITCH = 136;                                    // halfs per staged row (129 used)

__device__ __forceinline__ uint32_t pack_h2(__half lo, __half hi) {
    return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}
__device__ __forceinline__ void stem_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 3)
stem_u8_mma_kernel(const uint8_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int N, int H, int W,
                   int Ho, int Wo, int act, __half* __restrict__ y, int y_ctot, int y_coff, int tiles_x, int tiles_y) {
    __shared__ __half lut[256];
    __shared__ __align__(16) __half tile[ST_IN_H][ST16_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    lut[threadIdx.x] = unit_from_u8<__half>((int)threadIdx.x);
    // B fragments: this lane's column is gid, its taps 2 tig, 2 tig + 1 (and 8 for tig 0)
    uint32_t bh[4][2], bl[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = 8 * (gid >> 1) + 2 * j + (gid & 1);
        float wt[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int t = q < 2 ? 2 * tig + q : 8;
            wt[q] = (q < 2 || tig == 0) ? w[(t * 3 + 0) * 32 + ch] + w[(t * 3 + 1) * 32 + ch] + w[(t * 3 + 2) * 32 + ch] : 0.f;
        }
        __half hi[3], lo[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { hi[q] = __float2half_rn(wt[q]); lo[q] = __float2half_rn(wt[q] - __half2float(hi[q])); }
        const __half z = __float2half_rn(0.f);
        bh[j][0] = pack_h2(hi[0], hi[1]); bl[j][0] = pack_h2(lo[0], lo[1]);
        bh[j][1] = pack_h2(hi[2], z);     bl[j][1] = pack_h2(lo[2], z);
    }
    float hb[8];                                                   // bias / 2 of this lane's channels 8 tig .. 8 tig + 7
#pragma unroll
    for (int e = 0; e < 8; ++e) hb[e] = bias ? (act ? 0.5f : 1.f) * bias[8 * tig + e] : 0.f;
    // A fragments: taps 2 tig / 2 tig + 1 of pixel (row 2 r, column 2 ox): offsets in halfs from the pixel's top-left input
    const int k0 = 2 * tig, k1 = 2 * tig + 1;
    const int off0 = (k0 / 3) * ST16_PITCH + k0 % 3, off1 = (k1 / 3) * ST16_PITCH + k1 % 3, off8 = 2 * ST16_PITCH + 2;
    const int total = N * tiles_y * tiles_x;
    const int crow = threadIdx.x / 10, cchunk = threadIdx.x - crow * 10;
    const bool fetcher = threadIdx.x < ST_IN_H * 10;
    const bool vec_ok = (W & 15) == 0 && !(reinterpret_cast<uintptr_t>(x) & 15);
    auto fetch = [&](int tix, uint4& v, bool& ok) {
        ok = false;
        v = make_uint4(0u, 0u, 0u, 0u);
        if (!fetcher || tix >= total || !vec_ok) return;
        const int tx = tix % tiles_x, ty = (tix / tiles_x) % tiles_y, n = tix / (tiles_x * tiles_y);
        const int iy = 2 * ty * ST_TO_H - 1 + crow, xs = 2 * tx * ST_TO_W - 16 + cchunk * 16;
        if (iy < 0 || iy >= H || xs < 0 || xs + 16 > W) return;
        v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)n * H + iy) * W + xs));
        ok = true;
    };
    uint4 pre;
    bool pre_ok;
    fetch(blockIdx.x, pre, pre_ok);
    const __half hz = __float2half_rn(0.f);
    for (int tix = blockIdx.x; tix < total; tix += gridDim.x) {
        const int tx = tix % tiles_x, ty = (tix / tiles_x) % tiles_y, n = tix / (tiles_x * tiles_y);
        const int ox0 = tx * ST_TO_W, oy0 = ty * ST_TO_H;
        const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
        __syncthreads();                                           // the previous tile is consumed (and the table is written)
        if (vec_ok) {
            if (fetcher) {
                const uint32_t wv[4] = {pre.x, pre.y, pre.z, pre.w};
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int c = cchunk * 16 + k - 15;
                    if (c >= 0 && c < ST_IN_W) tile[crow][c] = pre_ok ? lut[(wv[k >> 2] >> (8 * (k & 3))) & 0xffu] : hz;
                }
            }
        } else {
            const uint8_t* img = x + (size_t)n * H * W;
            for (int i = threadIdx.x; i < ST_IN_H * ST_IN_W; i += 256) {
                const int r = i / ST_IN_W, c = i - r * ST_IN_W;
                const int iy = iy0 + r, ix = ix0 + c;
                tile[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? lut[__ldg(img + (size_t)iy * W + ix)] : hz;
            }
        }
        __syncthreads();
        fetch(tix + gridDim.x, pre, pre_ok);                       // in flight during the arithmetic below
        const int oy = oy0 + warp;                                 // one output row per warp, four M tiles along it
        if (oy >= Ho) continue;
        const __half* trow = &tile[2 * warp][0];
#pragma unroll
        for (int mt = 0; mt < ST_TO_W / 16; ++mt) {
            const __half* pa = trow + 2 * (mt * 16 + gid);         // pixel gid of the M tile; pixel gid + 8 is 16 halfs on
            const uint32_t a0 = pack_h2(pa[off0], pa[off1]);
            const uint32_t a1 = pack_h2(pa[16 + off0], pa[16 + off1]);
            const uint32_t a2 = tig == 0 ? (uint32_t)__half_as_ushort(pa[off8]) : 0u;
            const uint32_t a3 = tig == 0 ? (uint32_t)__half_as_ushort(pa[16 + off8]) : 0u;
            float c[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
                stem_mma(c[j], a0, a1, a2, a3, bl[j][0], bl[j][1]);
                stem_mma(c[j], a0, a1, a2, a3, bh[j][0], bh[j][1]);
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {                       // pixel gid, then pixel gid + 8
                const int ox = ox0 + mt * 16 + gid + 8 * rr;
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (act) {
                        const __half2 h2 = __floats2half2_rn(fmaf(c[j][2 * rr], 0.5f, hb[2 * j]), fmaf(c[j][2 * rr + 1], 0.5f, hb[2 * j + 1]));
                        uint32_t t;
                        asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&h2)));
                        const __half2 r2 = __hfma2(h2, *reinterpret_cast<const __half2*>(&t), h2);
                        o[j] = *reinterpret_cast<const uint32_t*>(&r2);
                    } else {
                        const __half2 r2 = __floats2half2_rn(c[j][2 * rr] + hb[2 * j], c[j][2 * rr + 1] + hb[2 * j + 1]);
                        o[j] = *reinterpret_cast<const uint32_t*>(&r2);
                    }
                }
                if (ox < Wo)
                    *reinterpret_cast<int4*>(y + ((size_t)(n * Ho + oy) * Wo + ox) * y_ctot + y_coff + tig * 8) =
                        make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
            }
        }
    }
}

// depthwise 3x3, stride 1, pad 1; w [9][C] fp16.  One thread = 8 channels (one 16-byte vector) x DW_PX consecutive
// output pixels of one row: its 3 x (DW_PX + 2) input vectors are 18 independent 16-byte loads issued back to back (288 B
// in flight per thread -- the round-1 kernel had 24 B and sat at 19 % of the DRAM peak), the 72 weights live in
// registers, and every loaded vector feeds up to three outputs.  A CTA covers DW_ROWS rows x a run of pixels x all channel
// groups, channel group fastest (a pixel's C * 2 bytes are contiguous), so the vertical re-use of input rows is served by
// L1 and the DRAM sees every input byte about 1.5 times.
constexpr int DW_PX = 4, DW_ROWS = 4, DW_THREADS = 256;

__global__ void __launch_bounds__(DW_THREADS, 1)
dwconv3x3_kernel(const __half* __restrict__ x, int x_ctot, int x_coff, const __half* __restrict__ w, const float* __restrict__ bias,
                 int N, int H, int W, int C, int act, __half* __restrict__ y, int y_ctot, int y_coff, int xb_per_cta, int tiles_x,
                 int tiles_y) {
    const int cgs = C >> 3;
    const int g = threadIdx.x % cgs;                               // channel group
    const int xb = (threadIdx.x / cgs) % xb_per_cta;               // block of DW_PX pixels inside the CTA's run
    const int rr = threadIdx.x / (cgs * xb_per_cta);               // row inside the CTA's strip
    float wr[9][8], b8[8];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int4 wv = __ldg(reinterpret_cast<const int4*>(w + k * C + g * 8));
        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(wh[e]); wr[k][2 * e] = f.x; wr[k][2 * e + 1] = f.y; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) b8[e] = bias ? __ldg(bias + g * 8 + e) : 0.f;
    const int total = N * tiles_y * tiles_x;
    for (int tix = blockIdx.x; tix < total; tix += gridDim.x) {
        const int tx = tix % tiles_x, ty = (tix / tiles_x) % tiles_y, n = tix / (tiles_x * tiles_y);
        const int oy = ty * DW_ROWS + rr, ox0 = (tx * xb_per_cta + xb) * DW_PX;
        if (rr >= DW_ROWS || oy >= H || ox0 >= W) continue;
        int4 in[3][DW_PX + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = oy - 1 + r;
#pragma unroll
            for (int c = 0; c < DW_PX + 2; ++c) {
                const int ix = ox0 - 1 + c;
                in[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                               ? __ldg(reinterpret_cast<const int4*>(x + ((size_t)(n * H + iy) * W + ix) * x_ctot + x_coff + g * 8))
                               : make_int4(0, 0, 0, 0);
            }
        }
        float acc[DW_PX][8];
#pragma unroll
        for (int j = 0; j < DW_PX; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = b8[e];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < DW_PX + 2; ++c) {
                float f[8];
                const __half2* h = reinterpret_cast<const __half2*>(&in[r][c]);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int j = c - s;                           // input column c is tap s of output pixel c - s
                    if (j >= 0 && j < DW_PX) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(f[e], wr[r * 3 + s][e], acc[j][e]);
                    }
                }
            }
#pragma unroll
        for (int j = 0; j < DW_PX; ++j) {
            if (ox0 + j >= W) break;
            int4 v;
            __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                h[e] = __floats2half2_rn(act ? silu_tanh(acc[j][2 * e]) : acc[j][2 * e], act ? silu_tanh(acc[j][2 * e + 1]) : acc[j][2 * e + 1]);
            *reinterpret_cast<int4*>(y + ((size_t)(n * H + oy) * W + ox0 + j) * y_ctot + y_coff + g * 8) = v;
        }
    }
}

// ---- depthwise 3x3 with the input staged by TMA.
// A CTA owns a 16 x 8 pixel x 64 channel output tile; its (18 x 10 pixel) input halo arrives as one 4-D TMA box
// (out-of-range rows / columns zero-filled by the TMA unit = the padding), double-buffered, so the copy of tile i+1
// runs under the arithmetic of tile i and no thread waits on a global load.  One thread = 8 channels x 4 pixels of a
// row: 18 conflict-free 16-byte shared-memory reads, 288 FMAs, four 16-byte stores.
constexpr int DT_W = 16, DT_H = 8, DT_C = 64, DT_HW = DT_W + 2, DT_HH = DT_H + 2;
constexpr int DT_STAGE = DT_HH * DT_HW * DT_C * 2;               // 23,040 B

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 2)
dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __half* __restrict__ w, const float* __restrict__ bias, int N, int H,
                     int W, int C, int act, __half* __restrict__ y, int y_ctot, int y_coff, int tiles_x, int tiles_y, int cblocks) {
    extern __shared__ __align__(128) unsigned char dsm[];
    unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsm) + 127) & ~(uintptr_t)127);
    __shared__ uint64_t full[2];
    const int tid = threadIdx.x;
    const int g = tid & 7, xb = (tid >> 3) & 3, rr = tid >> 5;     // channel group (8 ch), block of 4 pixels, row
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_u32(&full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_u32(&full[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int total = N * tiles_y * tiles_x * cblocks;
    auto issue = [&](int tix, int stage) {
        const int cb = tix % cblocks, t2 = tix / cblocks;
        const int tx = t2 % tiles_x, ty = (t2 / tiles_x) % tiles_y, n = t2 / (tiles_x * tiles_y);
        const uint32_t bar = dw_smem_u32(&full[stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)DT_STAGE) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(dw_smem_u32(tile + (size_t)stage * DT_STAGE)), "l"(&map_x), "r"(cb * DT_C), "r"(tx * DT_W - 1), "r"(ty * DT_H - 1), "r"(n), "r"(bar)
            : "memory");
    };
    if (tid == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);
    float wr[9][8], b8[8];
    int wcb = -1;
    uint32_t phase[2] = {0u, 0u};
    int it = 0;
    for (int tix = blockIdx.x; tix < total; tix += gridDim.x, ++it) {
        const int stage = it & 1;
        const int nxt = tix + gridDim.x;
        if (tid == 0 && nxt < total) issue(nxt, stage ^ 1);         // the other buffer was released by the barrier that ended tile it-1
        const int cb = tix % cblocks, t2 = tix / cblocks;
        const int tx = t2 % tiles_x, ty = (t2 / tiles_x) % tiles_y, n = t2 / (tiles_x * tiles_y);
        if (cb != wcb) {                                           // this channel block's weights and bias (cblocks is 2..8: rarely changes)
            wcb = cb;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int4 wv = __ldg(reinterpret_cast<const int4*>(w + k * C + cb * DT_C + g * 8));
                const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(wh[e]); wr[k][2 * e] = f.x; wr[k][2 * e + 1] = f.y; }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) b8[e] = bias ? __ldg(bias + cb * DT_C + g * 8 + e) : 0.f;
        }
        {
            const uint32_t bar = dw_smem_u32(&full[stage]);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(phase[stage]) : "memory");
            phase[stage] ^= 1u;
        }
        const unsigned char* src = tile + (size_t)stage * DT_STAGE;
        const int oy = ty * DT_H + rr, ox0 = tx * DT_W + xb * 4;
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = b8[e];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int4 v = *reinterpret_cast<const int4*>(src + ((size_t)((rr + r) * DT_HW + xb * 4 + c) * DT_C + g * 8) * 2);
                float f[8];
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
#pragma unroll
                for (int sft = 0; sft < 3; ++sft) {
                    const int j = c - sft;
                    if (j >= 0 && j < 4) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(f[e], wr[r * 3 + sft][e], acc[j][e]);
                    }
                }
            }
        if (oy < H) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ox0 + j >= W) break;
                int4 o;
                __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    h[e] = __floats2half2_rn(act ? silu_tanh(acc[j][2 * e]) : acc[j][2 * e], act ? silu_tanh(acc[j][2 * e + 1]) : acc[j][2 * e + 1]);
                *reinterpret_cast<int4*>(y + ((size_t)(n * H + oy) * W + ox0 + j) * y_ctot + y_coff + cb * DT_C + g * 8) = o;
            }
        }
        __syncthreads();                                           // everyone is done with this stage: it may be refilled
    }
}

typedef CUresult (*DwEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
DwEncodeFn dw_encoder() {
    static DwEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<DwEncodeFn>(p);
    });
    return fn;
}

int g_stem_scalar = 0;                                              // eitb_stem_debug(1): the CUDA-core u8 stem (A/B runs, tests)

}  // namespace

extern "C" int eitb_stem_debug(int scalar) { g_stem_scalar = scalar != 0; return EITB_OK; }

extern "C" int eitb_stem_conv3x3s2_nhwc(const void* x, int N, int H, int W, const float* w27, const float* bias, int Cout, int act,
                                        int gray, void* y, int y_ctot, int y_coff, eitb_stream_t stream) {
    if (!x || !w27 || !y || N <= 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if (Cout != 32 || y_ctot % 8 || y_coff % 8) return EITB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo;
    if (total * 4 >= (1LL << 32)) return EITB_ERR_UNSUPPORTED;
    eitb_prof_begin("stem_conv_kernel", s);
    if (gray == 2) {                                               // x is the u8 window image [N,H,W]
        const int tiles_x = eitb_div_up(Wo, ST_TO_W), tiles_y = eitb_div_up(Ho, ST_TO_H);
        const long long tiles = (long long)N * tiles_x * tiles_y;
        if (tiles > 0x7fffffffLL) return EITB_ERR_UNSUPPORTED;
        if (g_stem_scalar) {
            const int grid = tiles < 2 * EITB_NUM_SMS ? (int)tiles : 2 * EITB_NUM_SMS;
            stem_u8_kernel<<<grid, 256, 0, s>>>((const uint8_t*)x, w27, bias, N, H, W, Ho, Wo, act, (__half*)y, y_ctot, y_coff, tiles_x, tiles_y);
        } else {
            const int grid = tiles < 3 * EITB_NUM_SMS ? (int)tiles : 3 * EITB_NUM_SMS;
            stem_u8_mma_kernel<<<grid, 256, 0, s>>>((const uint8_t*)x, w27, bias, N, H, W, Ho, Wo, act, (__half*)y, y_ctot, y_coff, tiles_x, tiles_y);
        }
    } else if (gray)
        stem_gray_kernel<<<eitb_grid(total * 4, 256, 8), 256, 0, s>>>((const __half*)x, w27, bias, N, H, W, Ho, Wo, act, (__half*)y, y_ctot,
                                                                      y_coff);
    else
        stem_conv_kernel<32><<<eitb_grid(total, 256, 4), 256, 0, s>>>((const __half*)x, w27, bias, N, H, W, Ho, Wo, act, (__half*)y, y_ctot,
                                                                      y_coff);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_dwconv3x3_nhwc(const void* x, int N, int H, int W, int x_ctot, int x_coff, int C, const void* w9, const float* bias,
                                   int act, void* y, int y_ctot, int y_coff, eitb_stream_t stream) {
    if (!x || !w9 || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0) return EITB_ERR_BAD_ARG;
    if (C % 8 || x_ctot % 8 || x_coff % 8 || y_ctot % 8 || y_coff % 8 || C / 8 > DW_THREADS) return EITB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w9)) & 15) return EITB_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (C % DT_C == 0 && dw_encoder() && !(reinterpret_cast<uintptr_t>(x) & 15)) {
        // TMA-staged kernel: [N,H,W,x_ctot] seen as {C, W, H, N} with a (64 ch, 18 px, 10 rows, 1) box
        alignas(64) CUtensorMap mx;
        const char* xb0 = static_cast<const char*>(x) + (size_t)x_coff * 2;
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t strides[3] = {(cuuint64_t)x_ctot * 2, (cuuint64_t)W * x_ctot * 2, (cuuint64_t)H * W * x_ctot * 2};
        const cuuint32_t box[4] = {DT_C, DT_HW, DT_HH, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        if (dw_encoder()(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<char*>(xb0), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
            const int tiles_x = eitb_div_up(W, DT_W), tiles_y = eitb_div_up(H, DT_H), cblocks = C / DT_C;
            const long long tiles = (long long)N * tiles_x * tiles_y * cblocks;
            if (tiles <= 0x7fffffffLL) {
                const size_t smem = 2 * (size_t)DT_STAGE + 128;
                if (cudaFuncSetAttribute(dwconv3x3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                    return EITB_ERR_LAUNCH;
                const int grid = tiles < 2LL * EITB_NUM_SMS ? (int)tiles : 2 * EITB_NUM_SMS;
                eitb_prof_begin("dwconv3x3_kernel", s);
                dwconv3x3_tma_kernel<<<grid, 256, smem, s>>>(mx, (const __half*)w9, bias, N, H, W, C, act, (__half*)y, y_ctot, y_coff, tiles_x,
                                                             tiles_y, cblocks);
                EITB_CHECK_LAUNCH();
                return EITB_OK;
            }
        }
    }
    const int cgs = C / 8;
    int xb = DW_THREADS / (cgs * DW_ROWS);                        // pixel blocks per CTA row
    if (xb < 1) xb = 1;
    const int tiles_x = eitb_div_up(W, xb * DW_PX), tiles_y = eitb_div_up(H, DW_ROWS);
    const long long tiles = (long long)N * tiles_x * tiles_y;
    if (tiles > 0x7fffffffLL || (long long)N * H * W * x_ctot >= (1LL << 40)) return EITB_ERR_UNSUPPORTED;
    const int resident = eitb_resident_ctas(dwconv3x3_kernel, DW_THREADS, 0);
    const int grid = tiles < resident ? (int)tiles : resident;
    eitb_prof_begin("dwconv3x3_kernel", s);
    dwconv3x3_kernel<<<grid, DW_THREADS, 0, s>>>((const __half*)x, x_ctot, x_coff, (const __half*)w9, bias, N, H, W, C, act, (__half*)y,
                                                  y_ctot, y_coff, xb, tiles_x, tiles_y);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
