// K5: confidence filter + class-aware greedy NMS, one CTA per image.
//
// Restates ultralytics non_max_suppression + torchvision.ops.nms:
//   candidates: max_c score > conf;  box = (xy - wh/2, xy + wh/2) in fp32;
//   order: score descending, ties by ascending anchor index (stable sort);
//   suppression operands: box + cls*max_wh, IoU = inter / (a_i + a_j - inter), strict >;
//   at most max_det survivors.
// All fp32 arithmetic uses the round-to-nearest intrinsics so nvcc cannot contract
// mul+add into FMA: the keep-set must be bit-identical to the CPU kernel's.
//
// Phases (all in shared memory):
//   1 filter   each thread scans anchors, appends (~score_bits, anchor) keys
//   2 sort     bitonic sort of the next power of two >= #candidates
//   3 suppress tiles of 256 sorted candidates: test against the kept list, build the
//              256x256 intra-tile IoU bit matrix, one warp resolves it sequentially
//   4 emit     kept rows -> dets (x1,y1,x2,y2,conf,cls,coef...), keep_idx, n_out
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 256;
constexpr int kMaxDetCap = 1024;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

struct Box { float x1, y1, x2, y2, area; };

__device__ __forceinline__ bool iou_gt(const Box& a, const Box& b, float thr) {
    const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
    const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
    return ovr > thr;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
nms_kernel(const T* __restrict__ head, int nc, int nm, int A, int npad_cap, float conf, float iou, int max_det,
           float max_wh, float* __restrict__ dets, int32_t* __restrict__ keep_idx, int32_t* __restrict__ n_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);             // npad_cap
    Box* kept = reinterpret_cast<Box*>(keys + npad_cap);                                  // max_det
    int* kept_anchor = reinterpret_cast<int*>(kept + max_det);                            // max_det
    float* kept_conf = reinterpret_cast<float*>(kept_anchor + max_det);                   // max_det
    int* kept_cls = reinterpret_cast<int*>(kept_conf + max_det);                          // max_det
    Box* tile = reinterpret_cast<Box*>(kept_cls + max_det);                               // kTile
    unsigned* bits = reinterpret_cast<unsigned*>(tile + kTile);                           // kTile * kTile/32
    unsigned* alive = bits + kTile * (kTile / 32);                                        // kTile/32
    int* rank_of = reinterpret_cast<int*>(alive + kTile / 32);                            // kTile
    __shared__ int s_ncand, s_nkept;

    const int b = blockIdx.x;
    const int C = 4 + nc + nm;
    const T* hd = head + (long long)b * C * A;
    const int tid = threadIdx.x;
    if (tid == 0) { s_ncand = 0; s_nkept = 0; }
    __syncthreads();

    // ---- phase 1: filter
    for (int a = tid; a < A; a += kThreads) {
        float best = to_f32(hd[(long long)4 * A + a]);
        for (int c = 1; c < nc; ++c) best = fmaxf(best, to_f32(hd[(long long)(4 + c) * A + a]));
        if (best > conf) {
            const int slot = atomicAdd(&s_ncand, 1);
            keys[slot] = ((unsigned long long)(0xffffffffu - __float_as_uint(best)) << 32) | (unsigned)a;
        }
    }
    __syncthreads();
    const int ncand = s_ncand;
    int npad = 1;
    while (npad < ncand) npad <<= 1;
    for (int i = ncand + tid; i < npad; i += kThreads) keys[i] = ~0ull;
    __syncthreads();

    // ---- phase 2: bitonic sort ascending
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += kThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = keys[i], y = keys[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    }

    // ---- phase 3: greedy suppression, tile by tile
    for (int t0 = 0; t0 < ncand; t0 += kTile) {
        const int nk0 = s_nkept;
        if (nk0 >= max_det) break;
        const int idx = t0 + tid;
        const bool valid = idx < ncand;
        Box me = {0.f, 0.f, 0.f, 0.f, 0.f};
        int anchor = 0, cls = 0;
        float score = 0.f;
        bool ok = valid;
        if (valid) {
            anchor = (int)(keys[idx] & 0xffffffffu);
            score = __uint_as_float(0xffffffffu - (unsigned)(keys[idx] >> 32));
            // arg-max class: first index attaining the maximum
            float bestv = to_f32(hd[(long long)4 * A + anchor]);
            for (int c = 1; c < nc; ++c) {
                const float v = to_f32(hd[(long long)(4 + c) * A + anchor]);
                if (v > bestv) { bestv = v; cls = c; }
            }
            const float cx = to_f32(hd[anchor]), cy = to_f32(hd[(long long)A + anchor]);
            const float hw = to_f32(hd[(long long)2 * A + anchor]) / 2, hh = to_f32(hd[(long long)3 * A + anchor]) / 2;
            const float off = __fmul_rn((float)cls, max_wh);
            me.x1 = __fadd_rn(__fsub_rn(cx, hw), off);
            me.y1 = __fadd_rn(__fsub_rn(cy, hh), off);
            me.x2 = __fadd_rn(__fadd_rn(cx, hw), off);
            me.y2 = __fadd_rn(__fadd_rn(cy, hh), off);
            me.area = __fmul_rn(__fsub_rn(me.x2, me.x1), __fsub_rn(me.y2, me.y1));
            for (int q = 0; q < nk0 && ok; ++q) ok = !iou_gt(kept[q], me, iou);
        }
        tile[tid] = me;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0) alive[tid >> 5] = bal;
        __syncthreads();
        // row tid of the intra-tile matrix: which later tile members does tid suppress
#pragma unroll
        for (int w = 0; w < kTile / 32; ++w) {
            unsigned m = 0;
            if (ok && (w * 32 + 31 > tid)) {
                const unsigned al = alive[w];
                for (int l = 0; l < 32; ++l) {
                    const int j = w * 32 + l;
                    if (j > tid && ((al >> l) & 1u) && iou_gt(me, tile[j], iou)) m |= 1u << l;
                }
            }
            bits[tid * (kTile / 32) + w] = m;
        }
        __syncthreads();
        // sequential resolve by warp 0: lane w owns word w of the removed mask
        if (tid < 32) {
            unsigned removed = 0;                         // lanes >= kTile/32 stay 0
            int nk = nk0;
            for (int i = 0; i < kTile; ++i) {
                const int w = i >> 5;
                const unsigned rm = __shfl_sync(0xffffffffu, removed, w);
                const bool keep = nk < max_det && ((alive[w] >> (i & 31)) & 1u) && !((rm >> (i & 31)) & 1u);
                if (keep) {
                    if (tid < kTile / 32) removed |= bits[i * (kTile / 32) + tid];
                    ++nk;
                }
                if (tid == 0) rank_of[i] = keep ? nk - 1 : -1;
            }
            if (tid == 0) s_nkept = nk;
        }
        __syncthreads();
        // survivors publish themselves into the kept list
        const int r = rank_of[tid];
        if (r >= 0) {
            kept[r] = me;
            kept_conf[r] = score;
            kept_cls[r] = cls;
            kept_anchor[r] = anchor;
        }
        __syncthreads();
    }

    // ---- phase 4: emit
    const int nk = s_nkept;
    const int D = 6 + nm;
    float* out = dets + (long long)b * max_det * D;
    for (int i = tid; i < nk * D; i += kThreads) {
        const int r = i / D, f = i - r * D;
        const int a = kept_anchor[r];
        float v;
        if (f < 4) {
            const float cx = to_f32(hd[a]), cy = to_f32(hd[(long long)A + a]);
            const float hw = to_f32(hd[(long long)2 * A + a]) / 2, hh = to_f32(hd[(long long)3 * A + a]) / 2;
            v = f == 0 ? __fsub_rn(cx, hw) : f == 1 ? __fsub_rn(cy, hh) : f == 2 ? __fadd_rn(cx, hw) : __fadd_rn(cy, hh);
        } else if (f == 4) {
            v = kept_conf[r];
        } else if (f == 5) {
            v = (float)kept_cls[r];
        } else {
            v = to_f32(hd[(long long)(4 + nc + f - 6) * A + a]);
        }
        out[i] = v;
    }
    if (keep_idx)
        for (int r = tid; r < nk; r += kThreads) keep_idx[(long long)b * max_det + r] = kept_anchor[r];
    if (tid == 0) n_out[b] = nk;
}

size_t nms_smem_bytes(int npad_cap, int max_det) {
    return (size_t)npad_cap * 8 + (size_t)max_det * (sizeof(Box) + 12) + kTile * sizeof(Box) +
           (size_t)kTile * (kTile / 32) * 4 + (kTile / 32) * 4 + kTile * 4 + 16;
}

template <typename T>
int launch_nms(const void* head, int B, int nc, int nm, int A, float conf, float iou, int max_det, float max_wh,
               float* dets, int32_t* keep_idx, int32_t* n_out, cudaStream_t s) {
    int npad_cap = 1;
    while (npad_cap < A) npad_cap <<= 1;
    const size_t smem = nms_smem_bytes(npad_cap, max_det);
    if (smem > 220 * 1024) return EITB_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(nms_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    eitb_prof_begin("nms_kernel", s);
    nms_kernel<T><<<B, kThreads, smem, s>>>((const T*)head, nc, nm, A, npad_cap, conf, iou, max_det, max_wh, dets, keep_idx, n_out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

}  // namespace

extern "C" size_t eitb_nms_workspace_bytes(int B, int A) { (void)B; (void)A; return 0; }

extern "C" int eitb_nms(const void* head, int head_dtype, int B, int nc, int nm, int A, float conf, float iou,
                        int max_det, float max_wh, float* dets, int32_t* keep_idx, int32_t* n_out, void* ws,
                        size_t ws_bytes, eitb_stream_t stream) {
    (void)ws; (void)ws_bytes;
    if (!head || !dets || !n_out || B < 0 || nc <= 0 || nm < 0 || A <= 0 || max_det <= 0 || max_det > kMaxDetCap)
        return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (head_dtype) {
        case EITB_F32: return launch_nms<float>(head, B, nc, nm, A, conf, iou, max_det, max_wh, dets, keep_idx, n_out, s);
        case EITB_F16: return launch_nms<__half>(head, B, nc, nm, A, conf, iou, max_det, max_wh, dets, keep_idx, n_out, s);
        case EITB_BF16: return launch_nms<__nv_bfloat16>(head, B, nc, nm, A, conf, iou, max_det, max_wh, dets, keep_idx, n_out, s);
        default: return EITB_ERR_BAD_ARG;
    }
}
