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

// depthwise 3x3, stride 1, pad 1: one thread = one pixel x 8 channels; w [9][C] fp16
__global__ void __launch_bounds__(256)
dwconv3x3_kernel(const __half* __restrict__ x, int x_ctot, int x_coff, const __half* __restrict__ w, const float* __restrict__ bias,
                 int N, int H, int W, int C, int act, __half* __restrict__ y, int y_ctot, int y_coff) {
    const int cg = C >> 3;
    const long long total = (long long)N * H * W * cg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        const long long pix = i / cg;
        const int px = (int)(pix % W), py = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = bias ? __ldg(bias + g * 8 + e) : 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = py - 1 + r;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ix = px - 1 + s;
                if (ix < 0 || ix >= W) continue;
                const int4 xv = __ldg(reinterpret_cast<const int4*>(x + ((n * H + iy) * W + ix) * x_ctot + x_coff + g * 8));
                const int4 wv = __ldg(reinterpret_cast<const int4*>(w + (r * 3 + s) * C + g * 8));
                const __half2* xh = reinterpret_cast<const __half2*>(&xv);
                const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = __half22float2(xh[e]), b = __half22float2(wh[e]);
                    acc[2 * e] = fmaf(a.x, b.x, acc[2 * e]);
                    acc[2 * e + 1] = fmaf(a.y, b.y, acc[2 * e + 1]);
                }
            }
        }
        int4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = act ? silu_f(acc[2 * e]) : acc[2 * e], b = act ? silu_f(acc[2 * e + 1]) : acc[2 * e + 1];
            h[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<int4*>(y + pix * y_ctot + y_coff + g * 8) = v;
    }
}

}  // namespace

extern "C" int eitb_stem_conv3x3s2_nhwc(const void* x, int N, int H, int W, const float* w27, const float* bias, int Cout, int act,
                                        void* y, int y_ctot, int y_coff, eitb_stream_t stream) {
    if (!x || !w27 || !y || N <= 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if (Cout != 32 || y_ctot % 8 || y_coff % 8) return EITB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo;
    const int grid = eitb_grid(total, 256, 4);
    eitb_prof_begin("stem_conv_kernel", s);
    stem_conv_kernel<32><<<grid, 256, 0, s>>>((const __half*)x, w27, bias, N, H, W, Ho, Wo, act, (__half*)y, y_ctot, y_coff);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_dwconv3x3_nhwc(const void* x, int N, int H, int W, int x_ctot, int x_coff, int C, const void* w9, const float* bias,
                                   int act, void* y, int y_ctot, int y_coff, eitb_stream_t stream) {
    if (!x || !w9 || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0) return EITB_ERR_BAD_ARG;
    if (C % 8 || x_ctot % 8 || x_coff % 8 || y_ctot % 8 || y_coff % 8) return EITB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (long long)N * H * W * (C / 8);
    const int grid = eitb_grid(total, 256, 8);
    eitb_prof_begin("dwconv3x3_kernel", s);
    dwconv3x3_kernel<<<grid, 256, 0, s>>>((const __half*)x, x_ctot, x_coff, (const __half*)w9, bias, N, H, W, C, act, (__half*)y,
                                          y_ctot, y_coff);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
