// K9: in-place per-channel bias + SiLU on channels-last activations -- the epilogue of every
// Conv(+folded BatchNorm)+SiLU block of the YOLO11s-seg network (ultralytics fuses BN into the
// convolution at predict time, exactly like the host mirror does; SiLU stays a separate pass in
// plain PyTorch).  One read and one write per activation instead of three of each.
// HBM-bound: 16-byte loads/stores, bias staged in shared memory as fp32.
#include "common.cuh"

namespace {

template <typename T> struct Vec8;
template <> struct Vec8<__half> {
    static __device__ __forceinline__ void unpack(const int4& v, float (&f)[8]) {
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ int4 pack(const float (&f)[8]) {
        int4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void unpack(const int4& v, float (&f)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ int4 pack(const float (&f)[8]) {
        int4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
bias_act_kernel(T* __restrict__ x, const float* __restrict__ bias, long long n_vec, int C, int act) {
    extern __shared__ float sb[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) sb[c] = bias ? bias[c] : 0.f;
    __syncthreads();
    const int vpc = C >> 3;                                       // 8-element vectors per pixel
    int4* xv = reinterpret_cast<int4*>(x);
    constexpr int U = 4;                                           // 64 bytes in flight per thread
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i0 + u * stride < n_vec) v[u] = xv[i0 + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= n_vec) break;
            const int c0 = (int)(i % vpc) << 3;
            float f[8];
            Vec8<T>::unpack(v[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = f[j] + sb[c0 + j];
                f[j] = act ? __fdividef(a, 1.f + __expf(-a)) : a;   // SiLU
            }
            xv[i] = Vec8<T>::pack(f);
        }
    }
}

}  // namespace

extern "C" int eitb_bias_act_nhwc(void* x, int dtype, long long n_pixels, int C, const float* bias, int act,
                                  eitb_stream_t stream) {
    if (!x || n_pixels < 0 || C <= 0 || (act != 0 && act != 1)) return EITB_ERR_BAD_ARG;
    if ((C & 7) || (reinterpret_cast<uintptr_t>(x) & 15) || C > 8192) return EITB_ERR_UNSUPPORTED;
    if (n_pixels == 0) return EITB_OK;
    const long long n_vec = n_pixels * (C >> 3);
    const int grid = eitb_grid(n_vec, 256, 8);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)C * sizeof(float);
    eitb_prof_begin("bias_act_kernel", s);
    switch (dtype) {
        case EITB_F16: bias_act_kernel<__half><<<grid, 256, smem, s>>>((__half*)x, bias, n_vec, C, act); break;
        case EITB_BF16: bias_act_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>((__nv_bfloat16*)x, bias, n_vec, C, act); break;
        default: return EITB_ERR_UNSUPPORTED;
    }
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
