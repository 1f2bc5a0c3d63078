// Connected-component labelling primitive shared by K2 (body mask) and K7 (label clean-up).
//
// Labels are int32 per pixel, per image:  >= 0  parent / root pixel index (y*W + x) -- after
// flattening, the smallest pixel index of the component (== its first pixel in raster order,
// the order scipy.ndimage.label and cv2.findContours enumerate components in);
// CC_OUT (-1) the component touches the image frame ("outside", only when link_outside);
// CC_NONE (-2) the pixel is not in the set.
//
// Three launches: (1) one CTA per full-width strip of rows resolves the strip in shared
// memory (atomicMin union-find) and writes global labels, (2) strip borders and the frame are
// merged with the same union-find in global memory, (3) every pixel is pointed at its root.
#pragma once
#include "common.cuh"

#define CC_OUT (-1)
#define CC_NONE (-2)

namespace eitb_cc {

constexpr int kStripPixels = 16384;     // 64 KB of int32 labels per CTA
constexpr int kThreads = 512;

enum Pred { PRED_U8_NONZERO = 0, PRED_U8_ZERO = 1, PRED_LABEL_NOT_OUT = 2, PRED_CODE_NE = 3, PRED_CODE_NOT_BG = 4, PRED_BIT_ZERO = 5 };

// is pixel i of this image in the set?
template <int PRED>
__device__ __forceinline__ bool in_set(const void* __restrict__ src, long long i, int arg) {
    if (PRED == PRED_U8_NONZERO) return reinterpret_cast<const uint8_t*>(src)[i] != 0;
    if (PRED == PRED_U8_ZERO) return reinterpret_cast<const uint8_t*>(src)[i] == 0;
    if (PRED == PRED_LABEL_NOT_OUT) return reinterpret_cast<const int32_t*>(src)[i] != CC_OUT;
    if (PRED == PRED_CODE_NE) return reinterpret_cast<const uint8_t*>(src)[i] != (uint8_t)arg;
    if (PRED == PRED_BIT_ZERO) return ((reinterpret_cast<const uint32_t*>(src)[i >> 5] >> (i & 31)) & 1u) == 0;
    const uint8_t c = reinterpret_cast<const uint8_t*>(src)[i];           // PRED_CODE_NOT_BG
    return c != EITB_CODE_BLACK && c != EITB_CODE_MUSCLE;
}

__device__ __forceinline__ int sfind(const volatile int* L, int i) {
    int p;
    while ((p = L[i]) != i) i = p;
    return i;
}
__device__ __forceinline__ void sunion(int* L, int a, int b) {
    for (;;) {
        a = sfind(L, a); b = sfind(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// root of i in global labels: a pixel index, or CC_OUT
__device__ __forceinline__ int gfind(const int* L, int i) {
    while (i >= 0) {
        const int p = __ldcg(L + i);
        if (p == i) return i;
        i = p;
    }
    return CC_OUT;
}
__device__ __forceinline__ void gunion(int* L, int a, int b) {     // a: pixel in set; b: pixel in set or CC_OUT
    for (;;) {
        a = gfind(L, a);
        b = gfind(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }             // a > b, hence a >= 0
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// Strip-local labelling.  Pixels are visited in 32-pixel row segments, one per warp iteration:
// the segment's occupancy word comes from a ballot (or straight from a bit image), every pixel is
// pointed at the first pixel of its run inside the segment, and unions are issued only where runs
// meet -- segment to segment along a row, and at the first pixel of every overlap with the row
// above -- so a uniform region costs a handful of shared-memory atomics per row, not per pixel.
template <int PRED, int CONN>
__global__ void __launch_bounds__(kThreads)
cc_local_kernel(const void* __restrict__ src, size_t src_img_stride_bytes, int arg, int H, int W, int strip_h,
                int strips_per_img, int32_t* __restrict__ labels) {
    extern __shared__ int sl[];
    const int b = blockIdx.x / strips_per_img;
    const int sidx = blockIdx.x - b * strips_per_img;
    const int y0 = sidx * strip_h;
    const int rows = min(strip_h, H - y0);
    const int n = rows * W;
    const int spr = (W + 31) >> 5;                       // segments per row
    const int nseg = rows * spr;
    unsigned* segmask = reinterpret_cast<unsigned*>(sl + kStripPixels);
    const void* img = reinterpret_cast<const char*>(src) + (size_t)b * src_img_stride_bytes;
    const long long base = (long long)y0 * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kThreads >> 5;

    for (int sg = warp; sg < nseg; sg += nwarps) {
        const int y = sg / spr, x = ((sg - y * spr) << 5) + lane;
        const int i = y * W + x;
        const bool in = x < W && in_set<PRED>(img, base + i, arg);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) segmask[sg] = m;
        if (x < W) {
            int l = CC_NONE;
            if (in) l = i - (__clz(~(m << (31 - lane))) - 1);          // first pixel of the run within the segment
            sl[i] = l;
        }
    }
    __syncthreads();
    for (int sg = warp; sg < nseg; sg += nwarps) {
        const unsigned m = segmask[sg];
        if (m == 0) continue;
        const int y = sg / spr, sx = sg - y * spr, x = (sx << 5) + lane;
        const int i = y * W + x;
        const bool in = (m >> lane) & 1u;
        if (lane == 0 && in && sx > 0 && (segmask[sg - 1] >> 31)) sunion(sl, i, i - 1);
        if (y > 0) {
            const unsigned up = segmask[sg - spr];
            const unsigned both = m & up;
            if ((both & ~(both << 1)) >> lane & 1u) sunion(sl, i, i - W);
            if (CONN == 8 && in && !((up >> lane) & 1u)) {
                const unsigned upl = (up << 1) | (sx > 0 ? segmask[sg - spr - 1] >> 31 : 0u);
                const unsigned upr = (up >> 1) | (sx + 1 < spr ? segmask[sg - spr + 1] << 31 : 0u);
                if ((upl >> lane) & 1u) sunion(sl, i, i - W - 1);
                if ((upr >> lane) & 1u) sunion(sl, i, i - W + 1);
            }
        }
    }
    __syncthreads();
    int32_t* out = labels + (long long)b * H * W + base;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int l = sl[i];
        out[i] = l == CC_NONE ? CC_NONE : (int)base + sfind(sl, i);
    }
}

template <int CONN>
__global__ void __launch_bounds__(256)
cc_merge_kernel(int B, int H, int W, int strip_h, int link_outside, int32_t* __restrict__ labels) {
    const int nb = (H - 1) / strip_h;                    // number of strip borders per image
    const long long border_items = (long long)B * nb * W;
    const long long frame_per_img = link_outside ? 2LL * W + 2LL * H : 0;
    const long long total = border_items + (long long)B * frame_per_img;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        if (t < border_items) {
            const int b = (int)(t / ((long long)nb * W));
            const int r = (int)(t - (long long)b * nb * W);
            const int k = r / W, x = r - k * W;
            const int y = (k + 1) * strip_h;
            int* L = labels + (long long)b * H * W;
            const int i = y * W + x;
            if (L[i] == CC_NONE) continue;
            if (L[i - W] != CC_NONE) gunion(L, i, i - W);
            if (CONN == 8) {
                if (x > 0 && L[i - W - 1] != CC_NONE) gunion(L, i, i - W - 1);
                if (x + 1 < W && L[i - W + 1] != CC_NONE) gunion(L, i, i - W + 1);
            }
        } else {
            const long long u = t - border_items;
            const int b = (int)(u / frame_per_img);
            const int r = (int)(u - (long long)b * frame_per_img);
            int y, x;
            if (r < W) { y = 0; x = r; }
            else if (r < 2 * W) { y = H - 1; x = r - W; }
            else if (r < 2 * W + H) { y = r - 2 * W; x = 0; }
            else { y = r - 2 * W - H; x = W - 1; }
            int* L = labels + (long long)b * H * W;
            const int i = y * W + x;
            if (L[i] != CC_NONE) gunion(L, i, CC_OUT);
        }
    }
}

static __global__ void __launch_bounds__(256)
cc_flatten_kernel(long long n_total, int hw, int32_t* __restrict__ labels) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / hw;
        const int i = (int)(t - b * hw);
        const int* L = labels + b * hw;
        const int l = L[i];
        if (l == CC_NONE) continue;
        labels[t] = gfind(L, i);
    }
}

inline int strip_rows(int H, int W) {
    int sh = kStripPixels / W;
    if (sh < 1) sh = 1;
    if (sh > H) sh = H;
    return sh;
}

// labels [B,H,W] int32 out.  src: per-image stride in bytes (u8 images: H*W, int32 labels: 4*H*W).
template <int PRED, int CONN>
int cc_label(const void* src, size_t src_img_stride_bytes, int arg, int B, int H, int W, int link_outside,
             int32_t* labels, cudaStream_t s) {
    if (W > kStripPixels) return EITB_ERR_UNSUPPORTED;
    const int sh = strip_rows(H, W);
    const int spi = eitb_div_up(H, sh);
    const size_t smem = kStripPixels * sizeof(int) + (size_t)sh * ((W + 31) / 32) * sizeof(unsigned);
    if (cudaFuncSetAttribute(cc_local_kernel<PRED, CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    cc_local_kernel<PRED, CONN><<<B * spi, kThreads, smem, s>>>(src, src_img_stride_bytes, arg, H, W, sh, spi, labels);
    EITB_CHECK_LAUNCH();
    const long long items = (long long)B * ((H - 1) / sh) * W + (link_outside ? (long long)B * (2LL * W + 2LL * H) : 0);
    if (items > 0) {
        cc_merge_kernel<CONN><<<eitb_grid(items, 256, 8), 256, 0, s>>>(B, H, W, sh, link_outside, labels);
        EITB_CHECK_LAUNCH();
    }
    const long long n = (long long)B * H * W;
    cc_flatten_kernel<<<eitb_grid(n, 256, 8), 256, 0, s>>>(n, H * W, labels);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

}  // namespace eitb_cc
