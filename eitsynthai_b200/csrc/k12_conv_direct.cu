// K12: the two convolution shapes of YOLO11s-seg that are not GEMMs, as direct CUDA-core kernels with the
// same fused epilogue as K11 (folded-BatchNorm bias + SiLU):
//   * the stem  Conv(3 -> 32, k3, s2)  on the 3-channel network input (27 taps per output: no tensor-core shape);
//   * depthwise Conv(C -> C, k3, s1, groups = C) of the class branch of the Segment head and of the
//     attention position encoding (9 taps per channel: bandwidth-bound).
// Reference: the ultralytics modules behind model(...) at kt_service/ai_tools/ai_tools.py:121-122,153.
#include "common.cuh"

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
    for (int tix = blockIdx.x; tix < total; tix += gridDim.x) {
        const int tx = tix % tiles_x, ty = (tix / tiles_x) % tiles_y, n = tix / (tiles_x * tiles_y);
        const int ox0 = tx * ST_TO_W, oy0 = ty * ST_TO_H;
        const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
        const uint8_t* img = x + (size_t)n * H * W;
        __syncthreads();                                           // the previous tile is consumed (and the table is written)
        for (int i = threadIdx.x; i < ST_IN_H * ST_IN_W; i += 256) {
            const int r = i / ST_IN_W, c = i - r * ST_IN_W;
            const int iy = iy0 + r, ix = ix0 + c;
            tile[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? lut[__ldg(img + (size_t)iy * W + ix)] : 0.f;
        }
        __syncthreads();
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

}  // namespace

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
        const int grid = tiles < 2 * EITB_NUM_SMS ? (int)tiles : 2 * EITB_NUM_SMS;
        stem_u8_kernel<<<grid, 256, 0, s>>>((const uint8_t*)x, w27, bias, N, H, W, Ho, Wo, act, (__half*)y, y_ctot, y_coff, tiles_x, tiles_y);
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
