// K10: the two non-convolution tails of YOLO11s-seg that plain PyTorch runs as a dozen small
// kernels each, written as one pass over channels-last activations.
//
//   eitb_yolo_head_decode   Segment/Detect inference tail (ultralytics Detect._inference, SURVEY
//                           Appendix A.1): per anchor DFL softmax-expectation over 16 bins x 4 sides,
//                           dist2bbox(xywh) * stride, sigmoid class scores, mask coefficients ->
//                           head [B, 4+nc+nm, A] (the operand of K5).  Reads the 9 branch outputs
//                           where cuDNN left them (NHWC), one anchor per thread.
//   eitb_sppf_pool_concat   SPPF (yaml layer 9): x, maxpool5(x), maxpool5^2(x), maxpool5^3(x)
//                           (= 5x5, 9x9, 13x13 windows, -inf padding) written side by side into the
//                           concat buffer in one pass, separable running maxima in shared memory.
#include "common.cuh"

namespace {

struct HeadLevel {
    const __half* box;   // [B,h,w,64]
    const __half* cls;   // [B,h,w,nc]
    const __half* mc;    // [B,h,w,nm]
    int h, w, stride, a0;   // a0: first anchor index of the level
    const float* bbox;   // [64] bias of the last box conv, or NULL
    const float* bcls;   // [nc]
    const float* bmc;    // [nm]
};
struct HeadArgs { HeadLevel lv[3]; };

__global__ void __launch_bounds__(128)
head_decode_kernel(HeadArgs args, int B, int nc, int nm, int A, int cls_cstride, __half* __restrict__ head) {
    const int C = 4 + nc + nm;
    const long long total = (long long)B * A;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / A), a = (int)(t - (long long)b * A);
        const int l = a >= args.lv[2].a0 ? 2 : a >= args.lv[1].a0 ? 1 : 0;
        const HeadLevel L = args.lv[l];
        const int ai = a - L.a0;
        const int y = ai / L.w, x = ai - y * L.w;
        const long long pix = ((long long)b * L.h + y) * L.w + x;
        // DFL: expectation of softmax over 16 bins, per side (left, top, right, bottom)
        float d[4];
        const int4* bp = reinterpret_cast<const int4*>(L.box + pix * 64);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float v[16];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int4 raw = __ldg(bp + s * 2 + q);
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h2[k]); v[q * 8 + 2 * k] = f.x; v[q * 8 + 2 * k + 1] = f.y; }
            }
            if (L.bbox) {
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] += __ldg(L.bbox + s * 16 + k);
            }
            float m = v[0];
#pragma unroll
            for (int k = 1; k < 16; ++k) m = fmaxf(m, v[k]);
            float se = 0.f, sw = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) { const float e = __expf(v[k] - m); se += e; sw += e * (float)k; }
            d[s] = sw / se;
        }
        const float ax = (float)x + 0.5f, ay = (float)y + 0.5f, st = (float)L.stride;
        const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
        __half* o = head + (long long)b * C * A + a;
        o[0] = __float2half_rn((x1 + x2) * 0.5f * st);
        o[(long long)A] = __float2half_rn((y1 + y2) * 0.5f * st);
        o[(long long)2 * A] = __float2half_rn((x2 - x1) * st);
        o[(long long)3 * A] = __float2half_rn((y2 - y1) * st);
        const __half* cp = L.cls + pix * cls_cstride;
        for (int c = 0; c < nc; ++c) {
            const float z = __half2float(__ldg(cp + c)) + (L.bcls ? __ldg(L.bcls + c) : 0.f);
            o[(long long)(4 + c) * A] = __float2half_rn(1.f / (1.f + __expf(-z)));
        }
        const __half* mp = L.mc + pix * nm;
        if (L.bmc) {
            for (int c = 0; c < nm; ++c) o[(long long)(4 + nc + c) * A] = __float2half_rn(__half2float(__ldg(mp + c)) + __ldg(L.bmc + c));
        } else {
            for (int c = 0; c < nm; ++c) o[(long long)(4 + nc + c) * A] = __ldg(mp + c);
        }
    }
}

// one CTA per (image, group of 8 channels); h*w <= 1024 pixels
__global__ void __launch_bounds__(256)
sppf_kernel(const __half* __restrict__ x, int h, int w, int C, __half* __restrict__ out) {
    extern __shared__ int4 sm[];                 // [3][h*w] horizontal maxima for windows 5, 9, 13; then [h*w] input
    const int hw = h * w, vpc = C >> 3;
    const int b = blockIdx.x / vpc, cv = blockIdx.x - b * vpc;
    int4* in = sm + 3 * hw;
    const int4* src = reinterpret_cast<const int4*>(x + (long long)b * hw * C) + cv;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) in[p] = __ldg(src + (long long)p * vpc);
    __syncthreads();
    auto vmax = [](int4 a, int4 b4) {
        int4 r;
        const __half2* pa = reinterpret_cast<const __half2*>(&a);
        const __half2* pb = reinterpret_cast<const __half2*>(&b4);
        __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) pr[k] = __hmax2(pa[k], pb[k]);
        return r;
    };
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const int y = p / w, xx = p - y * w;
        int4 m = in[p];
        for (int r = 1; r <= 6; ++r) {
            if (xx - r >= 0) m = vmax(m, in[p - r]);
            if (xx + r < w) m = vmax(m, in[p + r]);
            if (r == 2) sm[p] = m; else if (r == 4) sm[hw + p] = m; else if (r == 6) sm[2 * hw + p] = m;
        }
    }
    __syncthreads();
    int4* dst = reinterpret_cast<int4*>(out + (long long)b * hw * 4 * C);
    const int ovpc = 4 * vpc;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const int y = p / w;
        dst[(long long)p * ovpc + cv] = in[p];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int R = 2 * (k + 1);
            const int4* hm = sm + k * hw;
            int4 m = hm[p];
            for (int r = 1; r <= R; ++r) {
                if (y - r >= 0) m = vmax(m, hm[p - r * w]);
                if (y + r < h) m = vmax(m, hm[p + r * w]);
            }
            dst[(long long)p * ovpc + (k + 1) * vpc + cv] = m;
        }
    }
}

}  // namespace

extern "C" int eitb_yolo_head_decode(const void* const* box, const void* const* cls, const void* const* mc,
                                     const float* const* box_bias, const float* const* cls_bias, const float* const* mc_bias,
                                     const int* hs, const int* ws, const int* strides, int B, int nc, int nm, int cls_cstride,
                                     void* head, eitb_stream_t stream) {
    if (!box || !cls || !mc || !hs || !ws || !strides || !head || B < 0 || nc <= 0 || nm < 0) return EITB_ERR_BAD_ARG;
    if (nm % 8 || (cls_cstride && cls_cstride < nc)) return EITB_ERR_UNSUPPORTED;
    if (!cls_cstride) cls_cstride = nc;
    HeadArgs a;
    int A = 0;
    for (int l = 0; l < 3; ++l) {
        if (!box[l] || !cls[l] || (nm && !mc[l]) || hs[l] <= 0 || ws[l] <= 0) return EITB_ERR_BAD_ARG;
        if (reinterpret_cast<uintptr_t>(box[l]) & 15) return EITB_ERR_UNSUPPORTED;
        a.lv[l] = HeadLevel{(const __half*)box[l], (const __half*)cls[l], (const __half*)mc[l], hs[l], ws[l], strides[l], A,
                            box_bias ? box_bias[l] : nullptr, cls_bias ? cls_bias[l] : nullptr, mc_bias ? mc_bias[l] : nullptr};
        A += hs[l] * ws[l];
    }
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    eitb_prof_begin("head_decode_kernel", s);
    head_decode_kernel<<<eitb_grid((long long)B * A, 128, 12), 128, 0, s>>>(a, B, nc, nm, A, cls_cstride, (__half*)head);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_sppf_pool_concat(const void* x, int B, int h, int w, int C, void* out, eitb_stream_t stream) {
    if (!x || !out || B < 0 || h <= 0 || w <= 0 || C <= 0) return EITB_ERR_BAD_ARG;
    if ((C & 7) || h * w > 2048 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15)) return EITB_ERR_UNSUPPORTED;
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)4 * h * w * sizeof(int4);
    if (cudaFuncSetAttribute(sppf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EITB_ERR_LAUNCH;
    eitb_prof_begin("sppf_kernel", s);
    sppf_kernel<<<B * (C >> 3), 256, smem, s>>>((const __half*)x, h, w, C, (__half*)out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
