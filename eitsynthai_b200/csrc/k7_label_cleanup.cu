// K7: label-image clean-up on code images, in place.
//
// Reference: create_color_output's tail (kt_service/ai_tools/utils.py:1005-1007):
//   clear_color_output (utils.py:691-755), only when a body mask with any non-zero pixel is given:
//     1. black pixels inside the body become muscle (:704-709);
//     2. 4-connected components (scipy.ndimage.label) of the pixels that are neither black nor
//        muscle (:712-721); those with < 5 pixels, in label order (= raster order of their first
//        pixel), take the most common non-background colour among the 8-neighbours of their
//        pixels -- votes listed direction-major then pixel-major, Counter keeps the first seen on
//        ties, own pixels vote, the image is updated as it goes -- or muscle when nobody votes
//        (:724-752);
//   highlight_small_masks (utils.py:758-843): for bone, muscle, adipose in that order, external
//     contours (RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) of the exact-colour mask of the INPUT image with
//     <= 5 points are filled (holes included) with the most common colour of the 1-pixel ring
//     around the filled contour, read from the progressively updated output in raster order and
//     ignoring the target colour and black; no vote -> the target colour itself (:807-839).
//
// Data-parallel parts (relabel, small-component detection by bounded flood, the frame flood that
// tells external contours from nested ones, candidate border tracing) run one thread per pixel;
// the order-dependent repaints run one warp per image over a bitmap of the few candidates.
//
// cv2's border following is restated from its published algorithm (Suzuki-Abe as implemented in
// OpenCV's contour tracer: clockwise search from the west neighbour, then counter-clockwise
// search from the previous pixel; a point is emitted where the chain direction changes).
#include "cc.cuh"
#include "bitflood.cuh"

namespace {

using namespace eitb_cc;

__device__ __forceinline__ bool not_bg(uint8_t c) { return c != EITB_CODE_BLACK && c != EITB_CODE_MUSCLE; }
__device__ __forceinline__ uint8_t ldv(const uint8_t* p) { return __ldcg(p); }     // L2-coherent read

// ------------------------------------------------------------------ stage A: black in body -> muscle
__global__ void __launch_bounds__(256)
fill_body_kernel(uint8_t* __restrict__ code, const uint8_t* __restrict__ body, int hw, int* __restrict__ anybody) {
    // grid (x: pixel blocks, y: image)
    const long long off = (long long)blockIdx.y * hw;
    bool any = false;
    if ((hw & 3) == 0 && !((reinterpret_cast<uintptr_t>(code) | reinterpret_cast<uintptr_t>(body)) & 3)) {
        for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hw; i += gridDim.x * blockDim.x * 4) {
            const uint32_t m = *reinterpret_cast<const uint32_t*>(body + off + i);
            if (!m) continue;
            any = true;
            uint32_t c = *reinterpret_cast<const uint32_t*>(code + off + i), o = c;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (((m >> (8 * j)) & 0xffu) == 255u && ((c >> (8 * j)) & 0xffu) == EITB_CODE_BLACK) o |= (uint32_t)EITB_CODE_MUSCLE << (8 * j);
            if (o != c) *reinterpret_cast<uint32_t*>(code + off + i) = o;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
            const uint8_t m = body[off + i];
            if (!m) continue;
            any = true;
            if (m == 255 && code[off + i] == EITB_CODE_BLACK) code[off + i] = EITB_CODE_MUSCLE;
        }
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) anybody[blockIdx.y] = 1;   // benign race: everyone writes 1
}

// ------------------------------------------------------------------ stage B: components with < 5 px
// Bounded flood over 4-neighbours; returns the component size (1..4) with its pixels sorted
// ascending in px[], or 5 when the component has at least 5 pixels.
__device__ __forceinline__ int small_component(const uint8_t* img, int H, int W, int p, int (&px)[5]) {
    int n = 1;
    px[0] = p;
    for (int h = 0; h < n && n < 5; ++h) {
        const int y = px[h] / W, x = px[h] - y * W;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int yy = y + (d == 0 ? -1 : d == 1 ? 1 : 0), xx = x + (d == 2 ? -1 : d == 3 ? 1 : 0);
            if (yy < 0 || yy >= H || xx < 0 || xx >= W || n >= 5) continue;
            const int q = yy * W + xx;
            if (!not_bg(ldv(img + q))) continue;
            bool seen = false;
            for (int k = 0; k < n; ++k) seen = seen || px[k] == q;
            if (!seen) px[n++] = q;
        }
    }
    if (n < 5)
        for (int i = 1; i < n; ++i)                                // insertion sort, n <= 4
            for (int j = i; j > 0 && px[j] < px[j - 1]; --j) { const int t = px[j]; px[j] = px[j - 1]; px[j - 1] = t; }
    return n;
}

__global__ void __launch_bounds__(256)
small_first_kernel(const uint8_t* __restrict__ code, const int* __restrict__ anybody, int B, int H, int W,
                   unsigned* __restrict__ bitmap, int words_per_img) {
    // grid (x: row chunks, y: row, z: image)
    const int b = blockIdx.z, y = blockIdx.y;
    if (!anybody[b]) return;
    const uint8_t* img = code + (long long)b * H * W;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        const int p = y * W + x;
        if (!not_bg(img[p])) continue;
        if ((x > 0 && not_bg(img[p - 1])) || (y > 0 && not_bg(img[p - W]))) continue;   // cannot be a first pixel
        int px[5];
        if (small_component(img, H, W, p, px) < 5 && px[0] == p)
            atomicOr(bitmap + (long long)b * words_per_img + (p >> 5), 1u << (p & 31));
    }
}

// The same test on 16 pixels per thread (rows that are multiples of 16 pixels): "not background" of a row chunk and of the
// chunk above it as 16-bit masks from two 16-byte loads; a pixel can only be the raster-first pixel of its component when
// neither its left nor its upper neighbour is set, so dense label images leave (almost) no survivor for the bounded flood.
__device__ __forceinline__ uint32_t nb_mask16(const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)                                     // byte > 1 <=> neither black (0) nor muscle (1)
        m |= ((((__vcmpgtu4(w[i], 0x01010101u) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * i);
    return m;
}
static_assert(EITB_CODE_BLACK == 0 && EITB_CODE_MUSCLE == 1, "nb_mask16 tests code > 1");

__global__ void __launch_bounds__(256)
small_first16_kernel(const uint8_t* __restrict__ code, const int* __restrict__ anybody, int B, int H, int W,
                     unsigned* __restrict__ bitmap, int words_per_img) {
    const long long per_img = (long long)H * W / 16, total = per_img * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per_img);
        if (!anybody[b]) continue;
        const int p0 = (int)(i - (long long)b * per_img) * 16;
        const uint8_t* img = code + (long long)b * H * W;
        const uint32_t m = nb_mask16(*reinterpret_cast<const uint4*>(img + p0));
        if (!m) continue;
        const int y = p0 / W, x0 = p0 - y * W;
        const uint32_t left = (m << 1) | ((x0 > 0 && not_bg(img[p0 - 1])) ? 1u : 0u);
        const uint32_t up = y > 0 ? nb_mask16(*reinterpret_cast<const uint4*>(img + p0 - W)) : 0u;
        uint32_t cand = m & ~left & ~up & 0xffffu;
        while (cand) {
            const int p = p0 + __ffs(cand) - 1;
            cand &= cand - 1;
            int px[5];
            if (small_component(img, H, W, p, px) < 5 && px[0] == p)
                atomicOr(bitmap + (long long)b * words_per_img + (p >> 5), 1u << (p & 31));
        }
    }
}

// Order-dependent repaints, parallel where the order cannot matter: a candidate reads and writes only the bounding box
// of its pixels grown by one pixel.  Every candidate registers that box in a coarse grid of cells (shared memory, one
// counter per cell); a candidate whose cells all count exactly one is isolated -- no other candidate's reads or writes
// come near it -- so the eight warps of the CTA process isolated candidates concurrently and clear their bits; one warp
// then walks the remaining (conflicting) candidates in the reference's order.
constexpr int kRepaintWarps = 8;
constexpr int kMaxCells = 4096;

__device__ __forceinline__ int cell_shift_for(int H, int W) {
    int cs = 3;
    while (((H >> cs) + 1) * ((W >> cs) + 1) > kMaxCells) ++cs;
    return cs;
}
// lanes cooperate: add `delta` to (delta != 0) or test == 1 (delta == 0) every cell of the box; returns "all cells == 1"
__device__ __forceinline__ bool cells_box(int* cells, int cw, int cs, int x0, int y0, int x1, int y1, int delta, int lane, int nlanes) {
    const int cx0 = x0 >> cs, cx1 = x1 >> cs, cy0 = y0 >> cs, cy1 = y1 >> cs;
    const int bw = cx1 - cx0 + 1, total = bw * (cy1 - cy0 + 1);
    bool ok = true;
    for (int c = lane; c < total; c += nlanes) {
        int* cell = cells + (cy0 + c / bw) * cw + cx0 + c % bw;
        if (delta) atomicAdd(cell, delta);
        else ok = ok && *cell == 1;
    }
    return ok;
}

// one candidate (warp-cooperative, uniform control flow): relabel the < 5-pixel component that starts at p
__device__ __forceinline__ void small_repaint_one(uint8_t* img, int H, int W, int p, int lane) {
    const int dY[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, dX[8] = {-1, 0, 1, -1, 1, -1, 0, 1};       // utils.py:734-736
    int px[5];
    const int n = small_component(img, H, W, p, px);               // uniform across the warp
    const int d = lane >> 2, k = lane & 3;
    int v = 0;
    if (k < n) {
        const int y = px[k] / W + dY[d], x = px[k] % W + dX[d];
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const uint8_t c = ldv(img + y * W + x);
            if (not_bg(c)) v = c;
        }
    }
    // Counter.most_common(1): highest count, first seen wins ties
    int best = EITB_CODE_MUSCLE, best_cnt = 0, best_first = 64;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const int val = ci == 0 ? EITB_CODE_ADIPOSE : ci == 1 ? EITB_CODE_LUNG : EITB_CODE_BONE;
        const unsigned m = __ballot_sync(0xffffffffu, v == val);
        const int cnt = __popc(m), first = m ? __ffs(m) - 1 : 64;
        if (cnt > best_cnt || (cnt == best_cnt && cnt > 0 && first < best_first)) { best = val; best_cnt = cnt; best_first = first; }
    }
    // any other non-background code (does not occur in overlay images) is left alone
    __syncwarp();
    if (d == 0 && k < n) __stcg(img + px[k], (uint8_t)best);
    __syncwarp();
}

// the box a small candidate touches: its (at most four) pixels grown by one
__device__ __forceinline__ void small_box(const uint8_t* img, int H, int W, int p, int& x0, int& y0, int& x1, int& y1) {
    int px[5];
    const int n = small_component(img, H, W, p, px);
    x0 = W; y0 = H; x1 = -1; y1 = -1;
    for (int k = 0; k < n && k < 4; ++k) {
        const int y = px[k] / W, x = px[k] - y * W;
        x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y);
    }
    x0 = max(x0 - 1, 0); y0 = max(y0 - 1, 0); x1 = min(x1 + 1, W - 1); y1 = min(y1 + 1, H - 1);
}

// one CTA per image; candidates in ascending raster order
__global__ void __launch_bounds__(kRepaintWarps * 32)
small_repaint_kernel(uint8_t* __restrict__ code, const int* __restrict__ anybody, int H, int W,
                     unsigned* __restrict__ bitmap, int words_per_img) {
    __shared__ int cells[kMaxCells];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!anybody[b]) return;
    uint8_t* img = code + (long long)b * H * W;
    unsigned* bm = bitmap + (long long)b * words_per_img;
    const int cs = cell_shift_for(H, W), cw = (W >> cs) + 1;
    for (int i = threadIdx.x; i < kMaxCells; i += kRepaintWarps * 32) cells[i] = 0;
    __syncthreads();
    // A. every candidate registers its box (one lane per bitmap word)
    for (int w = threadIdx.x; w < words_per_img; w += kRepaintWarps * 32) {
        unsigned word = bm[w];
        while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            int x0, y0, x1, y1;
            small_box(img, H, W, (w << 5) + bit, x0, y0, x1, y1);
            cells_box(cells, cw, cs, x0, y0, x1, y1, 1, 0, 1);
        }
    }
    __syncthreads();
    // B. isolated candidates, all warps (a chunk of 32 words per warp trip)
    for (int w0 = warp * 32; w0 < words_per_img; w0 += kRepaintWarps * 32) {
        const unsigned mine = w0 + lane < words_per_img ? bm[w0 + lane] : 0u;
        unsigned keep = mine;
        unsigned nz = __ballot_sync(0xffffffffu, mine != 0);
        while (nz) {
            const int wl = __ffs(nz) - 1;
            nz &= nz - 1;
            unsigned word = __shfl_sync(0xffffffffu, mine, wl);
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                const int p = ((w0 + wl) << 5) + bit;
                int x0, y0, x1, y1;
                small_box(img, H, W, p, x0, y0, x1, y1);           // uniform
                if (!__all_sync(0xffffffffu, cells_box(cells, cw, cs, x0, y0, x1, y1, 0, lane, 32))) continue;
                small_repaint_one(img, H, W, p, lane);
                if (lane == wl) keep &= ~(1u << bit);
            }
        }
        if (keep != mine) bm[w0 + lane] = keep;
    }
    __syncthreads();
    // C. the rest, in order, one warp
    if (warp != 0) return;
    for (int w0 = 0; w0 < words_per_img; w0 += 32) {
        const unsigned mine = w0 + lane < words_per_img ? bm[w0 + lane] : 0u;
        unsigned nz = __ballot_sync(0xffffffffu, mine != 0);
        while (nz) {
            const int wl = __ffs(nz) - 1;
            nz &= nz - 1;
            unsigned word = __shfl_sync(0xffffffffu, mine, wl);
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                small_repaint_one(img, H, W, ((w0 + wl) << 5) + bit, lane);
            }
        }
    }
}

// ------------------------------------------------------------------ stage C: highlight_small_masks
__device__ __forceinline__ bool is_t(const uint8_t* img, int H, int W, int y, int x, int t) {
    return y >= 0 && y < H && x >= 0 && x < W && img[y * W + x] == (uint8_t)t;
}

// cv2 outer-border following from the raster-first pixel (y0, x0) of a component of {img == t}.
// Emits the CHAIN_APPROX_SIMPLE points; returns their number, or 6 as soon as there are more than 5.
__device__ int trace_simple(const uint8_t* img, int H, int W, int y0, int x0, int t, int (&vx)[5], int (&vy)[5]) {
    const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1}, DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int s = 4;
    do { s = (s - 1) & 7; } while (!is_t(img, H, W, y0 + DY[s], x0 + DX[s], t) && s != 4);
    if (s == 4) { vx[0] = x0; vy[0] = y0; return 1; }                // isolated pixel
    const int y1 = y0 + DY[s], x1 = x0 + DX[s];
    int y3 = y0, x3 = x0, prev_s = s ^ 4, n = 0;
    for (;;) {
        int y4, x4;
        for (;;) {
            s = (s + 1) & 7;
            y4 = y3 + DY[s]; x4 = x3 + DX[s];
            if (is_t(img, H, W, y4, x4, t)) break;
        }
        if (s != prev_s) {
            if (n == 5) return 6;
            vx[n] = x3; vy[n] = y3; ++n;
        }
        prev_s = s;
        if (y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) break;
        y3 = y4; x3 = x4; s = (s + 4) & 7;
    }
    return n;
}

// Eight pixels per thread: the "first pixel of a component" test (no set neighbour earlier in
// raster order) is evaluated on byte masks; the few survivors resolve the label of their left
// neighbour on demand and trace their border.
__global__ void __launch_bounds__(256)
contour_cand_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ lab, int B, int H, int W, int t,
                    unsigned* __restrict__ bitmap, int words_per_img) {
    const int upr = (W + 7) >> 3;
    const long long n = (long long)B * H * upr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long rowid = i / upr;
        const int x0 = (int)(i - rowid * upr) << 3;
        const int b = (int)(rowid / H), y = (int)(rowid - (long long)b * H);
        const uint8_t* img = src + (long long)b * H * W;
        const uint8_t* row = img + (long long)y * W;
        unsigned cur = 0, up = 0;
        if (x0 + 8 <= W && !((reinterpret_cast<uintptr_t>(row + x0)) & 7)) {
            const uint2 v = *reinterpret_cast<const uint2*>(row + x0);
#pragma unroll
            for (int j = 0; j < 8; ++j) cur |= ((((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 0xffu) == (unsigned)t ? 1u : 0u) << j;
        } else {
            for (int j = 0; j < 8 && x0 + j < W; ++j) cur |= (row[x0 + j] == (uint8_t)t ? 1u : 0u) << j;
        }
        if (!cur) continue;
        unsigned left = x0 > 0 && row[x0 - 1] == (uint8_t)t, upl = 0, upright = 0;
        if (y > 0) {
            const uint8_t* prow = row - W;
            for (int j = 0; j < 8 && x0 + j < W; ++j) up |= (prow[x0 + j] == (uint8_t)t ? 1u : 0u) << j;
            upl = x0 > 0 && prow[x0 - 1] == (uint8_t)t;
            upright = x0 + 8 < W && prow[x0 + 8] == (uint8_t)t;
        }
        // a component's first pixel has no set neighbour earlier in raster order ...
        unsigned tips = cur & ~((cur << 1) | left) & ~up & ~((up << 1) | upl) & ~((up >> 1) | (upright << 7)) & 0xffu;
        while (tips) {
            const int j = __ffs(tips) - 1;
            tips &= tips - 1;
            const int x = x0 + j, p = y * W + x;
            // ... and an external contour has the frame-connected background on its left
            if (x > 0 && gfind(lab + (long long)b * H * W, p - 1) != CC_OUT) continue;
            int vx[5], vy[5];
            const int nv = trace_simple(img, H, W, y, x, t, vx, vy);
            if (nv > 5) continue;
            bool first = true;                                      // p must be the raster-first vertex
            for (int k = 0; k < nv; ++k) first = first && (vy[k] > y || (vy[k] == y && vx[k] >= x));
            if (first) atomicOr(bitmap + (long long)b * words_per_img + (p >> 5), 1u << (p & 31));
        }
    }
}

// All three targets in one pass over the snapshot (bit-image path): byte-compare masks of 8 pixels per thread give the
// "first pixel of a component" test for bone, muscle and adipose at once; survivors read one bit of the frame flood
// of their target (is the pixel on my left outside every contour of this colour?) and trace their border.
__device__ __forceinline__ unsigned eq_mask8(uint2 v, int t) {
    const uint32_t tt = (uint32_t)t * 0x01010101u;
    return ((((__vcmpeq4(v.x, tt) & 0x80808080u) * 0x00204081u) >> 28) | ((((__vcmpeq4(v.y, tt) & 0x80808080u) * 0x00204081u) >> 28) << 4));
}

// trace_simple without storing the vertices: does the border from (y0, x0) have at most five CHAIN_APPROX_SIMPLE points,
// none of them before (y0, x0) in raster order?  Chain-code steps come from two packed constants (no local arrays).
__device__ __forceinline__ int cdx(int s) { return (int)((0x21000122u >> (4 * s)) & 0xfu) - 1; }
__device__ __forceinline__ int cdy(int s) { return (int)((0x22210001u >> (4 * s)) & 0xfu) - 1; }

__device__ bool trace_is_small_first(const uint8_t* img, int H, int W, int y0, int x0, int t) {
    int s = 4;
    do { s = (s - 1) & 7; } while (!is_t(img, H, W, y0 + cdy(s), x0 + cdx(s), t) && s != 4);
    if (s == 4) return true;                                       // isolated pixel: one point
    const int y1 = y0 + cdy(s), x1 = x0 + cdx(s);
    int y3 = y0, x3 = x0, prev_s = s ^ 4, n = 0;
    for (;;) {
        int y4, x4;
        for (;;) {
            s = (s + 1) & 7;
            y4 = y3 + cdy(s); x4 = x3 + cdx(s);
            if (is_t(img, H, W, y4, x4, t)) break;
        }
        if (s != prev_s) {
            if (n == 5) return false;
            if (y3 < y0 || (y3 == y0 && x3 < x0)) return false;
            ++n;
        }
        prev_s = s;
        if (y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) break;
        y3 = y4; x3 = x4; s = (s + 4) & 7;
    }
    return true;
}

// Tracing is the expensive, divergent part (a few percent of the threads find a candidate), so the warp pools its
// candidates: every lane pushes the first pixels it found (after the one-bit frame-flood test) into a 64-entry queue in
// shared memory, and whenever 32 are waiting all 32 lanes trace one each.
__device__ __forceinline__ void trace_one(const uint8_t* __restrict__ src, int H, int W, unsigned item, unsigned* __restrict__ bitmap,
                                          int words_per_img) {
    const unsigned hw = (unsigned)H * W;
    const unsigned bk = item / hw;                                 // b * 3 + k
    const int p = (int)(item - bk * hw);
    const int k = (int)(bk % 3u), y = p / W, x = p - y * W;
    const int t = k == 0 ? EITB_CODE_BONE : k == 1 ? EITB_CODE_MUSCLE : EITB_CODE_ADIPOSE;
    if (trace_is_small_first(src + (size_t)(bk / 3u) * hw, H, W, y, x, t))
        atomicOr(bitmap + (size_t)bk * words_per_img + (p >> 5), 1u << (p & 31));
}

__global__ void __launch_bounds__(256)
contour_cand3_kernel(const uint8_t* __restrict__ src, const uint32_t* __restrict__ reach, int B, int H, int W,
                     unsigned* __restrict__ bitmap, int words_per_img) {
    __shared__ unsigned queue[8][64];
    unsigned* q = queue[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int qn = 0;                                                    // warp-uniform
    const unsigned upr = (unsigned)W >> 3;                         // W % 32 == 0 on this path
    const unsigned n = (unsigned)B * H * upr;                      // < 2^32 (checked by the launcher)
    const unsigned hw = (unsigned)H * W;
    const int T[3] = {EITB_CODE_BONE, EITB_CODE_MUSCLE, EITB_CODE_ADIPOSE};
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned n_round = (n + stride - 1) / stride * stride;   // every lane of a warp runs the same number of trips
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        unsigned tips[3] = {0u, 0u, 0u};
        unsigned rowbase = 0;                                      // (b * 3) * hw + y * W + x0
        unsigned bb = 0;
        if (i < n) {
            const unsigned rowid = i / upr;
            const int x0 = (int)(i - rowid * upr) << 3;
            const int b = (int)(rowid / (unsigned)H), y = (int)(rowid - (unsigned)b * H);
            const uint8_t* row = src + ((size_t)b * H + y) * W;
            const uint2 cv = *reinterpret_cast<const uint2*>(row + x0);
            bb = (unsigned)b;
            rowbase = (unsigned)(y * W + x0);
            if (cv.x | cv.y) {
                const int lb = x0 > 0 ? row[x0 - 1] : -1;
                uint2 uv = make_uint2(0u, 0u);
                int ulb = -1, urb = -1;
                if (y > 0) {
                    uv = *reinterpret_cast<const uint2*>(row - W + x0);
                    ulb = x0 > 0 ? row[x0 - 1 - W] : -1;
                    urb = x0 + 8 < W ? row[x0 + 8 - W] : -1;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int t = T[k];
                    const unsigned cur = eq_mask8(cv, t);
                    if (!cur) continue;
                    const unsigned up = y > 0 ? eq_mask8(uv, t) : 0u;
                    const unsigned left = lb == t, upl = ulb == t, upright = urb == t;
                    // a component's first pixel has no set neighbour earlier in raster order ...
                    tips[k] = cur & ~((cur << 1) | left) & ~up & ~((up << 1) | upl) & ~((up >> 1) | (upright << 7)) & 0xffu;
                }
            }
        }
        unsigned all = tips[0] | (tips[1] << 8) | (tips[2] << 16);
        while (__any_sync(0xffffffffu, all != 0u)) {
            unsigned item = 0xffffffffu;
            if (all) {
                const int bit = __ffs(all) - 1;
                all &= all - 1;
                const int k = bit >> 3, j = bit & 7;
                const unsigned p = rowbase + (unsigned)j;
                // ... and an external contour has the frame-connected background of its colour on its left
                bool ok = true;
                if ((p % (unsigned)W) != 0u) {
                    const unsigned qq = p - 1;
                    ok = (reach[((size_t)bb * 3 + k) * words_per_img + (qq >> 5)] >> (qq & 31)) & 1u;
                }
                if (ok) item = (bb * 3u + (unsigned)k) * hw + p;
            }
            const unsigned m = __ballot_sync(0xffffffffu, item != 0xffffffffu);
            if (item != 0xffffffffu) q[qn + __popc(m & lt_mask)] = item;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                const unsigned it = q[qn + lane];
                __syncwarp();
                trace_one(src, H, W, it, bitmap, words_per_img);
            }
        }
    }
    __syncwarp();
    if (lane < qn) trace_one(src, H, W, q[lane], bitmap, words_per_img);
}

// inside-or-on-boundary test against a closed polygon with integer vertices (exact)
__device__ __forceinline__ bool in_poly(int x, int y, const int (&vx)[5], const int (&vy)[5], int nv) {
    bool in = false;
    for (int i = 0; i < nv; ++i) {
        const int ax = vx[i], ay = vy[i], bx = vx[i + 1 == nv ? 0 : i + 1], by = vy[i + 1 == nv ? 0 : i + 1];
        if ((bx - ax) * (y - ay) - (by - ay) * (x - ax) == 0 && x >= min(ax, bx) && x <= max(ax, bx) &&
            y >= min(ay, by) && y <= max(ay, by))
            return true;
        if ((ay > y) != (by > y)) {
            const int num = (bx - ax) * (y - ay), den = by - ay, lhs = (x - ax) * den;
            if (den > 0 ? lhs < num : lhs > num) in = !in;
        }
    }
    return in;
}

// one candidate (warp-cooperative, uniform control flow): fill the contour traced from p with the ring's majority colour
__device__ __forceinline__ void contour_repaint_one(const uint8_t* simg, uint8_t* out, int H, int W, int t, int p, int lane) {
    int vx[5], vy[5];
    const int nv = trace_simple(simg, H, W, p / W, p % W, t, vx, vy);   // uniform across the warp
    int x0 = vx[0], x1 = vx[0], y0 = vy[0], y1 = vy[0];
    for (int k = 1; k < nv; ++k) { x0 = min(x0, vx[k]); x1 = max(x1, vx[k]); y0 = min(y0, vy[k]); y1 = max(y1, vy[k]); }
    // ring votes over the bounding box grown by one pixel, raster order
    const int rx0 = max(x0 - 1, 0), rx1 = min(x1 + 1, W - 1), ry0 = max(y0 - 1, 0), ry1 = min(y1 + 1, H - 1);
    const int bw = rx1 - rx0 + 1;
    const long long cells = (long long)bw * (ry1 - ry0 + 1);
    int cnt[4] = {0, 0, 0, 0};
    long long first[4] = {-1, -1, -1, -1};               // votes: muscle, adipose, lung, bone
    for (long long c0 = 0; c0 < cells; c0 += 32) {
        const long long c = c0 + lane;
        int v = 0;
        if (c < cells) {
            const int y = ry0 + (int)(c / bw), x = rx0 + (int)(c % bw);
            if (!in_poly(x, y, vx, vy, nv)) {
                bool near = false;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx)
                        if ((dy || dx) && x + dx >= x0 && x + dx <= x1 && y + dy >= y0 && y + dy <= y1)
                            near = near || in_poly(x + dx, y + dy, vx, vy, nv);
                if (near) {
                    const int cc = ldv(out + y * W + x);
                    if (cc != t && cc != EITB_CODE_BLACK) v = cc;
                }
            }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            const int val = ci == 0 ? EITB_CODE_MUSCLE : ci == 1 ? EITB_CODE_ADIPOSE : ci == 2 ? EITB_CODE_LUNG : EITB_CODE_BONE;
            const unsigned m = __ballot_sync(0xffffffffu, v == val);
            if (m) {
                cnt[ci] += __popc(m);
                if (first[ci] < 0) first[ci] = c0 + __ffs(m) - 1;
            }
        }
    }
    int fill = t, best_cnt = 0;
    long long best_first = 0;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
        const int val = ci == 0 ? EITB_CODE_MUSCLE : ci == 1 ? EITB_CODE_ADIPOSE : ci == 2 ? EITB_CODE_LUNG : EITB_CODE_BONE;
        if (cnt[ci] > best_cnt || (cnt[ci] == best_cnt && cnt[ci] > 0 && first[ci] < best_first)) {
            fill = val; best_cnt = cnt[ci]; best_first = first[ci];
        }
    }
    __syncwarp();
    const int fw = x1 - x0 + 1;
    const long long fcells = (long long)fw * (y1 - y0 + 1);
    for (long long c = lane; c < fcells; c += 32) {
        const int y = y0 + (int)(c / fw), x = x0 + (int)(c % fw);
        if (x >= 0 && x < W && y >= 0 && y < H && in_poly(x, y, vx, vy, nv)) __stcg(out + y * W + x, (uint8_t)fill);
    }
    __syncwarp();
}

// the box a contour candidate touches: the bounding box of its (at most five) vertices grown by one
__device__ __forceinline__ void contour_box(const uint8_t* simg, int H, int W, int t, int p, int& x0, int& y0, int& x1, int& y1) {
    int vx[5], vy[5];
    const int nv = trace_simple(simg, H, W, p / W, p % W, t, vx, vy);
    x0 = vx[0]; x1 = vx[0]; y0 = vy[0]; y1 = vy[0];
    for (int k = 1; k < nv && k < 5; ++k) { x0 = min(x0, vx[k]); x1 = max(x1, vx[k]); y0 = min(y0, vy[k]); y1 = max(y1, vy[k]); }
    x0 = max(x0 - 1, 0); y0 = max(y0 - 1, 0); x1 = min(x1 + 1, W - 1); y1 = min(y1 + 1, H - 1);
}

// one CTA per image, candidates in DESCENDING raster order (cv2 returns contours last-found first)
// nt == 1: the candidates of target t0 in bitmap [B, words]; nt == 3: bone, muscle, adipose one after the other
// (dict order, utils.py:782-787) from bitmap [B, 3, words].  Isolated candidates first, in parallel (see small_repaint).
__global__ void __launch_bounds__(kRepaintWarps * 32)
contour_repaint_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ code, int H, int W, int t0, int nt,
                       unsigned* __restrict__ bitmap, int words_per_img) {
    __shared__ int cells[kMaxCells];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t* simg = src + (long long)b * H * W;
    uint8_t* out = code + (long long)b * H * W;
    const int nchunks = (words_per_img + 31) / 32;
    const int cs = cell_shift_for(H, W), cw = (W >> cs) + 1;
    for (int i = threadIdx.x; i < kMaxCells; i += kRepaintWarps * 32) cells[i] = 0;
    __syncthreads();
    // A. every candidate of every target registers its box (one lane per bitmap word)
    for (int w = threadIdx.x; w < nt * words_per_img; w += kRepaintWarps * 32) {
        unsigned word = bitmap[(long long)b * nt * words_per_img + w];
        if (!word) continue;
        const int ti = w / words_per_img, wi = w - ti * words_per_img;
        const int t = nt == 1 ? t0 : ti == 0 ? EITB_CODE_BONE : ti == 1 ? EITB_CODE_MUSCLE : EITB_CODE_ADIPOSE;
        while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            int x0, y0, x1, y1;
            contour_box(simg, H, W, t, (wi << 5) + bit, x0, y0, x1, y1);
            cells_box(cells, cw, cs, x0, y0, x1, y1, 1, 0, 1);
        }
    }
    __syncthreads();
    // B. isolated candidates, all warps
    for (int idx = warp; idx < nt * nchunks; idx += kRepaintWarps) {
        const int ti = idx / nchunks, w0 = (idx - ti * nchunks) * 32;
        const int t = nt == 1 ? t0 : ti == 0 ? EITB_CODE_BONE : ti == 1 ? EITB_CODE_MUSCLE : EITB_CODE_ADIPOSE;
        unsigned* bm = bitmap + ((long long)b * nt + ti) * words_per_img;
        const unsigned mine = w0 + lane < words_per_img ? bm[w0 + lane] : 0u;
        unsigned keep = mine;
        unsigned nz = __ballot_sync(0xffffffffu, mine != 0);
        while (nz) {
            const int wl = __ffs(nz) - 1;
            nz &= nz - 1;
            unsigned word = __shfl_sync(0xffffffffu, mine, wl);
            while (word) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                const int p = ((w0 + wl) << 5) + bit;
                int x0, y0, x1, y1;
                contour_box(simg, H, W, t, p, x0, y0, x1, y1);     // uniform
                if (!__all_sync(0xffffffffu, cells_box(cells, cw, cs, x0, y0, x1, y1, 0, lane, 32))) continue;
                contour_repaint_one(simg, out, H, W, t, p, lane);
                if (lane == wl) keep &= ~(1u << bit);
            }
        }
        if (keep != mine) bm[w0 + lane] = keep;
    }
    __syncthreads();
    // C. the rest in the reference's order, one warp
    if (warp != 0) return;
    for (int ti = 0; ti < nt; ++ti) {
        const int t = nt == 1 ? t0 : ti == 0 ? EITB_CODE_BONE : ti == 1 ? EITB_CODE_MUSCLE : EITB_CODE_ADIPOSE;
        const unsigned* bm = bitmap + ((long long)b * nt + ti) * words_per_img;
        for (int ch = nchunks - 1; ch >= 0; --ch) {
            const int w0 = ch * 32;
            const unsigned mine = w0 + lane < words_per_img ? bm[w0 + lane] : 0u;
            unsigned nz = __ballot_sync(0xffffffffu, mine != 0);
            while (nz) {
                const int wl = 31 - __clz(nz);
                nz &= ~(1u << wl);
                unsigned word = __shfl_sync(0xffffffffu, mine, wl);
                while (word) {
                    const int bit = 31 - __clz(word);
                    word &= ~(1u << bit);
                    contour_repaint_one(simg, out, H, W, t, ((w0 + wl) << 5) + bit, lane);
                }
            }
        }
        __syncwarp();
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t eitb_label_cleanup_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t n = (size_t)B * H * W;
    const size_t words = ((size_t)H * W + 31) / 32;
    // labels int32, snapshot u8, bitmap, per-image flags
    return align256(n * 4) + align256(n) + align256((size_t)B * words * 4) + align256((size_t)B * 4);
}

extern "C" int eitb_label_cleanup(uint8_t* code, const uint8_t* body, int B, int H, int W, void* ws, size_t ws_bytes,
                                  eitb_stream_t stream) {
    if (!code || B < 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    if ((long long)H * W >= (1LL << 30)) return EITB_ERR_UNSUPPORTED;
    if (!ws || ws_bytes < eitb_label_cleanup_workspace_bytes(B, H, W)) return EITB_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)B * H * W;
    const int words = (int)(((size_t)H * W + 31) / 32);
    char* p = reinterpret_cast<char*>(ws);
    int32_t* lab = reinterpret_cast<int32_t*>(p); p += align256(n * 4);
    uint8_t* snap = reinterpret_cast<uint8_t*>(p); p += align256(n);
    unsigned* bitmap = reinterpret_cast<unsigned*>(p); p += align256((size_t)B * words * 4);
    int* anybody = reinterpret_cast<int*>(p);
    const int grid = eitb_grid((long long)n, 256, 8);

    if (body) {
        if (cudaMemsetAsync(anybody, 0, (size_t)B * 4, s) != cudaSuccess) return EITB_ERR_LAUNCH;
        if (cudaMemsetAsync(bitmap, 0, (size_t)B * words * 4, s) != cudaSuccess) return EITB_ERR_LAUNCH;
        eitb_prof_begin("fill_body_kernel", s);
        fill_body_kernel<<<dim3(eitb_grid_per_image((long long)H * W / 4 + 1, 256, B), B), 256, 0, s>>>(code, body, H * W, anybody);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("small_first_kernel", s);
        if (H > 65535 || B > 65535) return EITB_ERR_UNSUPPORTED;
        if (W % 16 == 0 && !(reinterpret_cast<uintptr_t>(code) & 15))
            small_first16_kernel<<<eitb_grid((long long)n / 16, 256, 8), 256, 0, s>>>(code, anybody, B, H, W, bitmap, words);
        else
            small_first_kernel<<<dim3(eitb_div_up(W, 256), H, B), 256, 0, s>>>(code, anybody, B, H, W, bitmap, words);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("small_repaint_kernel", s);
        small_repaint_kernel<<<B, kRepaintWarps * 32, 0, s>>>(code, anybody, H, W, bitmap, words);
        EITB_CHECK_LAUNCH();
    }
    if (cudaMemcpyAsync(snap, code, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return EITB_ERR_LAUNCH;
    const int targets[3] = {EITB_CODE_BONE, EITB_CODE_MUSCLE, EITB_CODE_ADIPOSE};            // dict order, utils.py:782-787
    if (eitb_flood::flood_supported(H, W) && !(reinterpret_cast<uintptr_t>(snap) & 15) && (size_t)B * 3 * words * 4 * 2 <= align256(n * 4) &&
        n / 8 < (1ull << 32)) {
        // bit-image path: one frame flood per (image, colour), one candidate pass and one repaint pass for all three
        uint32_t* reach = reinterpret_cast<uint32_t*>(lab);                     // [B, 3, words] in the label area
        unsigned* bitmap3 = reinterpret_cast<unsigned*>(lab) + (size_t)B * 3 * words;
        int rc = eitb_flood::frame_flood<eitb_flood::SRC_U8_NE>(snap, B, H, W, 3, targets[0], targets[1], targets[2], nullptr, reach, s);
        if (rc != EITB_OK) return rc;
        if (cudaMemsetAsync(bitmap3, 0, (size_t)B * 3 * words * 4, s) != cudaSuccess) return EITB_ERR_LAUNCH;
        eitb_prof_begin("contour_cand_kernel", s);
        contour_cand3_kernel<<<eitb_grid((long long)n / 8, 256, 8), 256, 0, s>>>(snap, reach, B, H, W, bitmap3, words);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("contour_repaint_kernel", s);
        contour_repaint_kernel<<<B, kRepaintWarps * 32, 0, s>>>(snap, code, H, W, 0, 3, bitmap3, words);
        EITB_CHECK_LAUNCH();
        return EITB_OK;
    }
    for (int k = 0; k < 3; ++k) {
        const int t = targets[k];
        const int rc = cc_label<PRED_CODE_NE, 4>(snap, (size_t)H * W, t, B, H, W, 1, lab, s, /*flatten=*/0);
        if (rc != EITB_OK) return rc;
        if (cudaMemsetAsync(bitmap, 0, (size_t)B * words * 4, s) != cudaSuccess) return EITB_ERR_LAUNCH;
        eitb_prof_begin("contour_cand_kernel", s);
        contour_cand_kernel<<<grid, 256, 0, s>>>(snap, lab, B, H, W, t, bitmap, words);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("contour_repaint_kernel", s);
        contour_repaint_kernel<<<B, kRepaintWarps * 32, 0, s>>>(snap, code, H, W, t, 1, bitmap, words);
        EITB_CHECK_LAUNCH();
    }
    return EITB_OK;
}
