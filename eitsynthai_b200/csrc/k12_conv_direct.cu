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

// depthwise 3x3, stride 1, pad 1; w [9][C] fp16.  One thread = 4 channels x one column x a strip of DW_ROWS output
// rows: the nine weight vectors stay in registers and every input row is loaded once (3 x 8 B, the next row
// prefetched while this one is used) and feeds the three output rows it touches -- 3.75 loads per output, not 18.
constexpr int DW_ROWS = 8;

__device__ __forceinline__ void dw_fma4(float* acc, const uint2& xv, const float* wf) {
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
    const float2 a = __half22float2(xh[0]), b = __half22float2(xh[1]);
    acc[0] = fmaf(a.x, wf[0], acc[0]); acc[1] = fmaf(a.y, wf[1], acc[1]);
    acc[2] = fmaf(b.x, wf[2], acc[2]); acc[3] = fmaf(b.y, wf[3], acc[3]);
}

__global__ void __launch_bounds__(256, 3)
dwconv3x3_kernel(const __half* __restrict__ x, int x_ctot, int x_coff, const __half* __restrict__ w, const float* __restrict__ bias,
                 int N, int H, int W, int C, int act, __half* __restrict__ y, int y_ctot, int y_coff) {
    const int cg = C >> 2, strips = (H + DW_ROWS - 1) / DW_ROWS;
    const unsigned total = (unsigned)N * strips * W * cg;          // < 2^32: checked by the launcher
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        unsigned t = i / (unsigned)cg;
        const int g = (int)(i - t * cg);
        unsigned t2 = t / (unsigned)W;
        const int px = (int)(t - t2 * W);
        const size_t n = t2 / (unsigned)strips;
        const int st = (int)(t2 - (unsigned)n * strips);
        float wr[9][4], b4[4];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const uint2 wv = __ldg(reinterpret_cast<const uint2*>(w + k * C + g * 4));
            const __half2* wh = reinterpret_cast<const __half2*>(&wv);
            const float2 p = __half22float2(wh[0]), q = __half22float2(wh[1]);
            wr[k][0] = p.x; wr[k][1] = p.y; wr[k][2] = q.x; wr[k][3] = q.y;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) b4[e] = bias ? __ldg(bias + g * 4 + e) : 0.f;
        const int y0 = st * DW_ROWS, y1 = min(y0 + DW_ROWS, H);
        float a0[4], a1[4], a2[4];                              // output rows iy-1, iy, iy+1 while input row iy is read
#pragma unroll
        for (int e = 0; e < 4; ++e) { a0[e] = b4[e]; a1[e] = b4[e]; a2[e] = b4[e]; }
        const bool has_l = px > 0, has_r = px + 1 < W;
        const uint2 z = make_uint2(0u, 0u);
        auto load = [&](int iy, uint2& l, uint2& c, uint2& r) {
            if (iy >= 0 && iy < H) {
                const __half* row = x + ((n * H + iy) * W + px) * x_ctot + x_coff + g * 4;
                l = has_l ? __ldg(reinterpret_cast<const uint2*>(row - x_ctot)) : z;
                c = __ldg(reinterpret_cast<const uint2*>(row));
                r = has_r ? __ldg(reinterpret_cast<const uint2*>(row + x_ctot)) : z;
            } else {
                l = z; c = z; r = z;
            }
        };
        uint2 xl, xc, xr, nl, nc, nr;
        load(y0 - 1, xl, xc, xr);
        for (int iy = y0 - 1; iy <= y1; ++iy) {
            load(iy + 1 <= y1 ? iy + 1 : -1, nl, nc, nr);       // prefetch the next input row
            // input row iy is tap row 2 of output iy-1 (a0), row 1 of output iy (a1), row 0 of output iy+1 (a2)
            dw_fma4(a0, xl, wr[6]); dw_fma4(a0, xc, wr[7]); dw_fma4(a0, xr, wr[8]);
            dw_fma4(a1, xl, wr[3]); dw_fma4(a1, xc, wr[4]); dw_fma4(a1, xr, wr[5]);
            dw_fma4(a2, xl, wr[0]); dw_fma4(a2, xc, wr[1]); dw_fma4(a2, xr, wr[2]);
            const int oy = iy - 1;                              // a0 is complete once input row oy + 1 has been added
            if (oy >= y0 && oy < y1) {
                uint2 v;
                __half2* h = reinterpret_cast<__half2*>(&v);
                h[0] = __floats2half2_rn(act ? silu_f(a0[0]) : a0[0], act ? silu_f(a0[1]) : a0[1]);
                h[1] = __floats2half2_rn(act ? silu_f(a0[2]) : a0[2], act ? silu_f(a0[3]) : a0[3]);
                *reinterpret_cast<uint2*>(y + ((n * H + oy) * W + px) * y_ctot + y_coff + g * 4) = v;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) { a0[e] = a1[e]; a1[e] = a2[e]; a2[e] = b4[e]; }
            xl = nl; xc = nc; xr = nr;
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
    if (gray)
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
    if (C % 8 || x_ctot % 8 || x_coff % 8 || y_ctot % 8 || y_coff % 8) return EITB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (long long)N * ((H + DW_ROWS - 1) / DW_ROWS) * W * (C / 4);
    if (total >= (1LL << 32)) return EITB_ERR_UNSUPPORTED;
    const int grid = eitb_grid(total, 256, 3);
    eitb_prof_begin("dwconv3x3_kernel", s);
    dwconv3x3_kernel<<<grid, 256, 0, s>>>((const __half*)x, x_ctot, x_coff, (const __half*)w9, bias, N, H, W, C, act, (__half*)y,
                                          y_ctot, y_coff);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
