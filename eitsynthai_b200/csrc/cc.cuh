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
// the segment's occupancy word comes from a ballot, one thread per row then chains the segments
// so that every pixel points straight at the first pixel of its row run, and unions are issued
// only at the first pixel of every overlap between a run and the row above -- a uniform region
// costs one shared-memory atomic per row, not per pixel.  Run starts are compressed to their
// roots before the labels are written.
template <int PRED, int CONN>
__global__ void __launch_bounds__(kThreads)
cc_local_kernel(const void* __restrict__ src, size_t src_img_stride_bytes, int arg, int H, int W, int strip_h,
                int strips_per_img, int32_t* __restrict__ labels, const int* __restrict__ skip) {
    extern __shared__ int sl[];
    const int b = blockIdx.x / strips_per_img;
    if (skip && skip[b]) return;                         // this image was answered without labelling (K2 fast path)
    const int sidx = blockIdx.x - b * strips_per_img;
    const int y0 = sidx * strip_h;
    const int rows = min(strip_h, H - y0);
    const int spr = (W + 31) >> 5;                       // segments per row
    const int nseg = rows * spr;
    unsigned* segmask = reinterpret_cast<unsigned*>(sl + kStripPixels);
    int* segstart = reinterpret_cast<int*>(segmask + nseg);          // start of the run entering the segment from the left
    const void* img = reinterpret_cast<const char*>(src) + (size_t)b * src_img_stride_bytes;
    const long long base = (long long)y0 * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kThreads >> 5;
    const int wy0 = warp / spr, wsx0 = warp - wy0 * spr;             // the only division: first (row, segment) of the warp
    const int dy = nwarps / spr, dsx = nwarps - dy * spr;             // and its stride

    // ---- occupancy word of every segment
    for (int y = wy0, sx = wsx0; y < rows;) {
        const int x = (sx << 5) + lane;
        const bool in = x < W && in_set<PRED>(img, base + y * W + x, arg);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) segmask[y * spr + sx] = m;
        y += dy; sx += dsx;
        if (sx >= spr) { sx -= spr; ++y; }
    }
    __syncthreads();
    // ---- chain the segments of every row: where does the run that crosses into segment s begin?
    for (int y = threadIdx.x; y < rows; y += kThreads) {
        int start = -1;                                   // first pixel of the run touching the right edge of the previous segment
        for (int sx = 0; sx < spr; ++sx) {
            const unsigned m = segmask[y * spr + sx];
            const int entering = (m & 1u) ? start : -1;
            segstart[y * spr + sx] = entering;
            if (m == 0xffffffffu) { if (entering < 0) start = y * W + (sx << 5); }
            else if (m >> 31) start = y * W + (sx << 5) + (32 - __clz(~m));
            else start = -1;
        }
    }
    __syncthreads();
    // ---- every pixel points at the first pixel of its row run
    for (int y = wy0, sx = wsx0; y < rows;) {
        const int sg = y * spr + sx, x = (sx << 5) + lane;
        const unsigned m = segmask[sg];
        if (x < W) {
            int l = CC_NONE;
            if ((m >> lane) & 1u) {
                const int ones = __clz(~(m << (31 - lane)));            // set pixels ending here, within the segment
                l = y * W + x - (ones - 1);
                if (ones == lane + 1 && segstart[sg] >= 0) l = segstart[sg];
            }
            sl[y * W + x] = l;
        }
        y += dy; sx += dsx;
        if (sx >= spr) { sx -= spr; ++y; }
    }
    __syncthreads();
    // ---- one union per overlap between a run and the row above
    for (int y = wy0, sx = wsx0; y < rows;) {
        const int sg = y * spr + sx;
        const unsigned m = y > 0 ? segmask[sg] : 0u;
        if (m) {
            const int i = y * W + (sx << 5) + lane;
            const unsigned up = segmask[sg - spr];
            const unsigned both = m & up;
            const unsigned carry = sx > 0 ? (segmask[sg - 1] & segmask[sg - spr - 1]) >> 31 : 0u;
            if (((both & ~((both << 1) | carry)) >> lane) & 1u) sunion(sl, i, i - W);
            if (CONN == 8 && ((m >> lane) & 1u) && !((up >> lane) & 1u)) {
                const unsigned upl = (up << 1) | (sx > 0 ? segmask[sg - spr - 1] >> 31 : 0u);
                const unsigned upr = (up >> 1) | (sx + 1 < spr ? segmask[sg - spr + 1] << 31 : 0u);
                if ((upl >> lane) & 1u) sunion(sl, i, i - W - 1);
                if ((upr >> lane) & 1u) sunion(sl, i, i - W + 1);
            }
        }
        y += dy; sx += dsx;
        if (sx >= spr) { sx -= spr; ++y; }
    }
    __syncthreads();
    // ---- compress the run starts; afterwards every pixel is two hops from its root
    for (int y = wy0, sx = wsx0; y < rows;) {
        const int sg = y * spr + sx;
        const unsigned m = segmask[sg];
        const bool start = ((m >> lane) & 1u) && (lane == 0 ? segstart[sg] < 0 : !((m >> (lane - 1)) & 1u));
        if (start) {
            const int i = y * W + (sx << 5) + lane;
            sl[i] = sfind(sl, i);
        }
        y += dy; sx += dsx;
        if (sx >= spr) { sx -= spr; ++y; }
    }
    __syncthreads();
    int32_t* out = labels + (long long)b * H * W + base;
    for (int y = wy0, sx = wsx0; y < rows;) {
        const int x = (sx << 5) + lane;
        if (x < W) {
            const int i = y * W + x, l = sl[i];
            out[i] = l == CC_NONE ? CC_NONE : (int)base + sl[l];
        }
        y += dy; sx += dsx;
        if (sx >= spr) { sx -= spr; ++y; }
    }
}

template <int CONN>
__global__ void __launch_bounds__(256)
cc_merge_kernel(int B, int H, int W, int strip_h, int link_outside, int32_t* __restrict__ labels, const int* __restrict__ skip) {
    const int nb = (H - 1) / strip_h;                    // number of strip borders per image
    const long long border_items = (long long)B * nb * W;
    const long long frame_per_img = link_outside ? 2LL * W + 2LL * H : 0;
    const long long total = border_items + (long long)B * frame_per_img;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        if (t < border_items) {
            const int b = (int)(t / ((long long)nb * W));
            if (skip && skip[b]) continue;
            const int r = (int)(t - (long long)b * nb * W);
            const int k = r / W, x = r - k * W;
            const int y = (k + 1) * strip_h;
            int* L = labels + (long long)b * H * W;
            const int i = y * W + x;
            if (L[i] == CC_NONE) continue;
            if (L[i - W] != CC_NONE) {
                // one union per run of vertically adjacent pairs
                if (x == 0 || L[i - 1] == CC_NONE || L[i - W - 1] == CC_NONE) gunion(L, i, i - W);
            } else if (CONN == 8) {
                if (x > 0 && L[i - W - 1] != CC_NONE) gunion(L, i, i - W - 1);
                if (x + 1 < W && L[i - W + 1] != CC_NONE) gunion(L, i, i - W + 1);
            }
        } else {
            const long long u = t - border_items;
            const int b = (int)(u / frame_per_img);
            if (skip && skip[b]) continue;
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
cc_flatten_kernel(int hw, int32_t* __restrict__ labels, const int* __restrict__ skip) {       // grid (x: pixel blocks, y: image)
    if (skip && skip[blockIdx.y]) return;
    int* L = labels + (long long)blockIdx.y * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const int l = L[i];
        if (l == CC_NONE) continue;
        L[i] = gfind(L, l);                               // l is the strip root (or CC_OUT) already
    }
}

inline int strip_rows(int H, int W) {
    int sh = kStripPixels / W;
    if (sh < 1) sh = 1;
    if (sh > H) sh = H;
    return sh;
}

// labels [B,H,W] int32 out.  src: per-image stride in bytes (u8 images: H*W, int32 labels: 4*H*W).
// flatten == 0 leaves every pixel pointing at its strip root; callers then resolve the few labels
// they need with gfind().
template <int PRED, int CONN>
int cc_label(const void* src, size_t src_img_stride_bytes, int arg, int B, int H, int W, int link_outside,
             int32_t* labels, cudaStream_t s, int flatten = 1, const int* skip = nullptr) {
    if (W > kStripPixels) return EITB_ERR_UNSUPPORTED;
    const int sh = strip_rows(H, W);
    const int spi = eitb_div_up(H, sh);
    const size_t smem = kStripPixels * sizeof(int) + (size_t)sh * ((W + 31) / 32) * 2 * sizeof(unsigned);
    if (cudaFuncSetAttribute(cc_local_kernel<PRED, CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    eitb_prof_begin("cc_local_kernel", s);
    cc_local_kernel<PRED, CONN><<<B * spi, kThreads, smem, s>>>(src, src_img_stride_bytes, arg, H, W, sh, spi, labels, skip);
    EITB_CHECK_LAUNCH();
    const long long items = (long long)B * ((H - 1) / sh) * W + (link_outside ? (long long)B * (2LL * W + 2LL * H) : 0);
    if (items > 0) {
        eitb_prof_begin("cc_merge_kernel", s);
        cc_merge_kernel<CONN><<<eitb_grid(items, 256, 8), 256, 0, s>>>(B, H, W, sh, link_outside, labels, skip);
        EITB_CHECK_LAUNCH();
    }
    if (!flatten) return EITB_OK;
    eitb_prof_begin("cc_flatten_kernel", s);
    cc_flatten_kernel<<<dim3(eitb_grid_per_image((long long)H * W, 256, B), B), 256, 0, s>>>(H * W, labels, skip);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

}  // namespace eitb_cc
