// K13: polygon extraction from the cleaned label image, on the device.
//
// Reference: create_list_crd_from_color_output (kt_service/ai_tools/utils.py:1191-1279) and
// get_only_body_mask_contours (utils.py:1157-1188): per tissue colour, in the dict order of utils.py:1228-1233,
//   cv2.inRange -> cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)            (utils.py:1244-1251)
//   eps = 0.001 * cv2.arcLength(cnt, True); cv2.approxPolyDP(cnt, eps, True)       (utils.py:1256-1257)
//   polygons with more than two points are closed by repeating the first point     (utils.py:1260-1266)
// then the body outline: findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) of the body mask, the last contour (in
// OpenCV's order) with at least five points, every border pixel (utils.py:1173-1184).
//
// OpenCV's algorithms are restated from their published form (the same statements as oracle/contours.py, which the
// CPU tests pin against cv2 itself): Suzuki-Abe border following as OpenCV's tracer walks it (clockwise search from the
// west neighbour, then counter-clockwise from the previous pixel; CHAIN_APPROX_SIMPLE keeps a point where the chain
// code changes), contours returned last-found first, arcLength = float32 segment lengths summed in float64,
// approxPolyDP = three farthest-point sweeps for the start, an explicit slice stack, distance to the chord *segment*,
// first maximum wins, and the final pass that drops points on almost straight lines.  All decisions are made on exact
// integers (squared distances times squared chord length fit 64 bits for images up to 1024 x 1024); only the two
// comparisons against eps^2 are in fp64.
//
// Work decomposition: the external/nested decision is the bit-parallel frame flood of bitflood.cuh (one CTA per image
// and colour); candidate first pixels are found eight pixels per thread on byte-compare masks; one CTA per image then
// validates, orders (descending raster order per colour), traces and simplifies the contours, one thread per contour
// -- this stage serves the 1-3 selected slices of a request, not the 320-slice stream.
#include "bitflood.cuh"

namespace {

constexpr int kJobs = 5;                         // adipose "3", bone "0", muscle "1", lung "2", body "4"
constexpr int kThreads = 256;

__device__ __forceinline__ int job_target(int j) {
    return j == 0 ? EITB_CODE_ADIPOSE : j == 1 ? EITB_CODE_BONE : j == 2 ? EITB_CODE_MUSCLE : EITB_CODE_LUNG;
}
__device__ __forceinline__ int job_class(int j) { return j == 0 ? 3 : j == 1 ? 0 : j == 2 ? 1 : j == 3 ? 2 : 4; }

// chain code s -> step: 0 = east, counter-clockwise on the screen (y grows downwards)
__device__ __forceinline__ int step_dx(int s) { return (int)((0x21000122u >> (4 * s)) & 0xfu) - 1; }
__device__ __forceinline__ int step_dy(int s) { return (int)((0x22210001u >> (4 * s)) & 0xfu) - 1; }

// Where the tracer reads "is this pixel in the set": a one-bit-per-pixel plane in shared memory (the usual case: the
// five planes of a 512 x 512 image are 160 KB), or the byte image in global memory when the planes do not fit.
struct BitPlane {
    const uint32_t* plane; int H, W, wpr;
    __device__ __forceinline__ bool operator()(int y, int x) const {
        if (y < 0 || y >= H || x < 0 || x >= W) return false;
        return (plane[y * wpr + (x >> 5)] >> (x & 31)) & 1u;
    }
};
struct ByteImage {                               // t >= 0: pixel == t (tissue colour); t < 0: pixel != 0 (body mask)
    const uint8_t* img; int H, W, t;
    __device__ __forceinline__ bool operator()(int y, int x) const {
        if (y < 0 || y >= H || x < 0 || x >= W) return false;
        const int v = img[y * W + x];
        return t >= 0 ? v == t : v != 0;
    }
};

// Outer border from the candidate first pixel (y0, x0).  emit(i, x, y) receives the points in OpenCV's order; returns
// their number, or -1 as soon as the border reaches a pixel that precedes (y0, x0) in raster order (then (y0, x0) is
// not the first pixel of its component and some other candidate owns this border).
template <class On, class Emit>
__device__ int trace_border(const On& on, int y0, int x0, bool simple, Emit emit) {
    int s = 4;
    do { s = (s - 1) & 7; } while (!on(y0 + step_dy(s), x0 + step_dx(s)) && s != 4);
    if (s == 4) { emit(0, x0, y0); return 1; }                       // isolated pixel (the west neighbour is never set here)
    const int y1 = y0 + step_dy(s), x1 = x0 + step_dx(s);
    int y3 = y0, x3 = x0, prev_s = s ^ 4, n = 0;
    for (;;) {
        int y4, x4;
        for (;;) {
            s = (s + 1) & 7;
            y4 = y3 + step_dy(s); x4 = x3 + step_dx(s);
            if (on(y4, x4)) break;
        }
        if (!simple || s != prev_s) { emit(n, x3, y3); ++n; }
        prev_s = s;
        if (y4 < y0 || (y4 == y0 && x4 < x0)) return -1;
        if (y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) break;
        y3 = y4; x3 = x4; s = (s + 4) & 7;
    }
    return n;
}

struct NoEmit { __device__ void operator()(int, int, int) const {} };
struct StoreEmit {
    int32_t* dst;
    __device__ void operator()(int i, int x, int y) const { dst[i] = (x & 0xffff) | (y << 16); }
};
__device__ __forceinline__ int pt_x(int32_t p) { return p & 0xffff; }
__device__ __forceinline__ int pt_y(int32_t p) { return (int)((uint32_t)p >> 16); }

// ---- candidates: pixels of the set with no set neighbour earlier in raster order and the frame-connected outside on
// their left.  Eight pixels per thread on byte-compare masks.
__device__ __forceinline__ unsigned eq_mask8(uint2 v, int t) {
    const uint32_t tt = (uint32_t)t * 0x01010101u;
    return ((((__vcmpeq4(v.x, tt) & 0x80808080u) * 0x00204081u) >> 28) | ((((__vcmpeq4(v.y, tt) & 0x80808080u) * 0x00204081u) >> 28) << 4));
}
__device__ __forceinline__ unsigned nz_mask8(uint2 v) { return eq_mask8(v, 0) ^ 0xffu; }

__global__ void __launch_bounds__(256)
poly_tips_kernel(const uint8_t* __restrict__ code, const uint8_t* __restrict__ body, const uint32_t* __restrict__ reach,
                 int B, int H, int W, unsigned* __restrict__ tips, int words) {
    const unsigned upr = (unsigned)W >> 3;                         // W % 32 == 0
    const unsigned n = (unsigned)B * H * upr;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned rowid = i / upr;
        const int x0 = (int)(i - rowid * upr) << 3;
        const int b = (int)(rowid / (unsigned)H), y = (int)(rowid - (unsigned)b * H);
#pragma unroll 1
        for (int src = 0; src < 2; ++src) {
            const uint8_t* base = src == 0 ? code : body;
            if (!base) continue;
            const uint8_t* row = base + ((size_t)b * H + y) * W;
            const uint2 cv = *reinterpret_cast<const uint2*>(row + x0);
            if (!(cv.x | cv.y)) continue;                          // black / outside the body: no candidates
            const int lb = x0 > 0 ? row[x0 - 1] : -1;
            uint2 uv = make_uint2(0u, 0u);
            int ulb = -1, urb = -1;
            if (y > 0) {
                uv = *reinterpret_cast<const uint2*>(row - W + x0);
                ulb = x0 > 0 ? row[x0 - 1 - W] : -1;
                urb = x0 + 8 < W ? row[x0 + 8 - W] : -1;
            }
            const int j0 = src == 0 ? 0 : 4, j1 = src == 0 ? 4 : 5;
            for (int j = j0; j < j1; ++j) {
                unsigned cur, up, left, upl, upright;
                if (j < 4) {
                    const int t = job_target(j);
                    cur = eq_mask8(cv, t);
                    if (!cur) continue;
                    up = y > 0 ? eq_mask8(uv, t) : 0u;
                    left = lb == t; upl = ulb == t; upright = urb == t;
                } else {
                    cur = nz_mask8(cv);
                    up = y > 0 ? nz_mask8(uv) : 0u;
                    left = lb > 0; upl = ulb > 0; upright = urb > 0;
                }
                unsigned t8 = cur & ~((cur << 1) | left) & ~up & ~((up << 1) | upl) & ~((up >> 1) | (upright << 7)) & 0xffu;
                const uint32_t* rj = reach + ((size_t)b * kJobs + j) * words;
                unsigned keep = 0;
                while (t8) {
                    const int k = __ffs(t8) - 1;
                    t8 &= t8 - 1;
                    const int x = x0 + k;
                    if (x > 0) {
                        const int q = y * W + x - 1;
                        if (!((rj[q >> 5] >> (q & 31)) & 1u)) continue;      // nested inside another contour of this set
                    }
                    keep |= 1u << k;
                }
                if (keep) {
                    const int p = y * W + x0;                      // x0 % 8 == 0: the 8 bits stay inside one word
                    atomicOr(tips + ((size_t)b * kJobs + j) * words + (p >> 5), keep << (p & 31));
                }
            }
        }
    }
}

// ---- cv2.approxPolyDP(cnt, eps, True) on the contour src[0..count), result to dst; returns the new count
__device__ int approx_closed(const int32_t* __restrict__ src, int count, double eps2, int32_t* __restrict__ dst,
                             int2* __restrict__ stack) {
    int new_count = 0, top = 0;
    // 1. approximately the two farthest points
    int pos = 0, right_start = 0;
    bool le_eps = false;
    int32_t start_pt = 0;
    for (int it = 0; it < 3; ++it) {
        long long max_dist = 0;
        pos = (pos + right_start) % count;
        start_pt = src[pos];
        if (++pos >= count) pos = 0;
        for (int j = 1; j < count; ++j) {
            const int32_t pt = src[pos];
            if (++pos >= count) pos = 0;
            const long long dx = pt_x(pt) - pt_x(start_pt), dy = pt_y(pt) - pt_y(start_pt);
            const long long dist = dx * dx + dy * dy;
            if (dist > max_dist) { max_dist = dist; right_start = j; }
        }
        le_eps = (double)max_dist <= eps2;
    }
    // 2. the two initial slices
    if (!le_eps) {
        const int slice_start = pos % count;
        const int slice_end = (right_start + slice_start) % count;
        stack[top++] = make_int2(slice_end, slice_start);          // right slice
        stack[top++] = make_int2(slice_start, slice_end);
    } else {
        dst[new_count++] = start_pt;
    }
    // 3. split until every slice is within eps of its chord segment
    while (top > 0) {
        const int2 sl = stack[--top];
        const int32_t end_pt = src[sl.y];
        pos = sl.x;
        start_pt = src[pos];
        if (++pos >= count) pos = 0;
        bool le = true;
        int r_start = 0;
        if (pos != sl.y) {
            const long long sx = pt_x(start_pt), sy = pt_y(start_pt), ex = pt_x(end_pt), ey = pt_y(end_pt);
            const long long dx = ex - sx, dy = ey - sy, L = dx * dx + dy * dy;
            long long max_key = 0;                                 // squared distance x L (x 1 when the chord is a point)
            while (pos != sl.y) {
                const int32_t pt = src[pos];
                if (++pos >= count) pos = 0;
                const long long px = pt_x(pt) - sx, py = pt_y(pt) - sy;
                const long long t = px * dx + py * dy;
                long long key;
                if (t <= 0) key = (px * px + py * py) * (L ? L : 1);
                else if (t >= L) { const long long qx = pt_x(pt) - ex, qy = pt_y(pt) - ey; key = (qx * qx + qy * qy) * L; }
                else { const long long c = py * dx - px * dy; key = c * c; }
                if (key > max_key) { max_key = key; r_start = (pos + count - 1) % count; }
            }
            le = (double)max_key <= __dmul_rn(eps2, (double)(L ? L : 1));
        }
        if (le) {
            dst[new_count++] = start_pt;
        } else {
            stack[top++] = make_int2(r_start, sl.y);
            stack[top++] = make_int2(sl.x, r_start);
        }
    }
    // 4. drop the points that lie on [almost] straight lines
    count = new_count;
    pos = count - 1;
    start_pt = dst[pos]; if (++pos >= count) pos = 0;
    int wpos = pos;
    int32_t pt = dst[pos]; if (++pos >= count) pos = 0;
    for (int i = 0; i < count && new_count > 2; ++i) {
        const int32_t end_pt = dst[pos]; if (++pos >= count) pos = 0;
        const long long dx = pt_x(end_pt) - pt_x(start_pt), dy = pt_y(end_pt) - pt_y(start_pt);
        const long long cr = (long long)(pt_x(pt) - pt_x(start_pt)) * dy - (long long)(pt_y(pt) - pt_y(start_pt)) * dx;
        const long long inner = (long long)(pt_x(pt) - pt_x(start_pt)) * (pt_x(end_pt) - pt_x(pt)) +
                                (long long)(pt_y(pt) - pt_y(start_pt)) * (pt_y(end_pt) - pt_y(pt));
        if ((double)(cr * cr) <= __dmul_rn(0.5 * eps2, (double)(dx * dx + dy * dy)) && dx != 0 && dy != 0 && inner >= 0) {
            --new_count;
            dst[wpos] = start_pt = end_pt;
            if (++wpos >= count) wpos = 0;
            pt = dst[pos]; if (++pos >= count) pos = 0;
            ++i;
            continue;
        }
        dst[wpos] = start_pt = pt;
        if (++wpos >= count) wpos = 0;
        pt = end_pt;
    }
    return new_count;
}

struct PolyWs {
    uint32_t* reach;       // [B, 5, words]
    unsigned* tips;        // [B, 5, words]
    int32_t* slot_p;       // [B, max_slots]   first pixel
    int32_t* slot_job;     // [B, max_slots]
    int32_t* slot_cnt;     // [B, max_slots]   raw points, then output points
    int32_t* slot_off;     // [B, max_slots]   offset into raw / dst
    int32_t* raw;          // [B, raw_cap]
    int32_t* dst;          // [B, raw_cap + max_slots]
    int2* stack;           // [B, raw_cap + 2 * max_slots]
    int max_slots, raw_cap;
};

__global__ void __launch_bounds__(kThreads)
poly_build_kernel(const uint8_t* __restrict__ code, const uint8_t* __restrict__ body, int H, int W, int words, PolyWs ws,
                  int max_polys, int max_points, int32_t* __restrict__ n_polys, int32_t* __restrict__ poly_cls,
                  int32_t* __restrict__ poly_off, int32_t* __restrict__ points_xy, int32_t* __restrict__ status, int use_planes) {
    __shared__ int s_scan[kThreads];
    __shared__ int s_base, s_nslots, s_status, s_body_slot;
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint8_t* cimg = code + (size_t)b * H * W;
    const uint8_t* bimg = body ? body + (size_t)b * H * W : nullptr;
    unsigned* tips = ws.tips + (size_t)b * kJobs * words;
    int32_t* slot_p = ws.slot_p + (size_t)b * ws.max_slots;
    int32_t* slot_job = ws.slot_job + (size_t)b * ws.max_slots;
    int32_t* slot_cnt = ws.slot_cnt + (size_t)b * ws.max_slots;
    int32_t* slot_off = ws.slot_off + (size_t)b * ws.max_slots;
    int32_t* raw = ws.raw + (size_t)b * ws.raw_cap;
    int32_t* dst = ws.dst + (size_t)b * (ws.raw_cap + ws.max_slots);
    int2* stack = ws.stack + (size_t)b * (ws.raw_cap + 2 * ws.max_slots);
    if (tid == 0) { s_base = 0; s_status = 0; s_body_slot = -1; }
    const int njobs = bimg ? kJobs : kJobs - 1;
    const int wpr = W >> 5, hw = H * W;
    // the five sets as bit planes in shared memory (when they fit): border following then probes shared memory
    extern __shared__ uint32_t planes[];
    if (use_planes) {
        for (int w = tid; w < words; w += kThreads) {
            const int4 v0 = *reinterpret_cast<const int4*>(cimg + (size_t)w * 32), v1 = *reinterpret_cast<const int4*>(cimg + (size_t)w * 32 + 16);
            const uint2 q[4] = {make_uint2((unsigned)v0.x, (unsigned)v0.y), make_uint2((unsigned)v0.z, (unsigned)v0.w),
                                make_uint2((unsigned)v1.x, (unsigned)v1.y), make_uint2((unsigned)v1.z, (unsigned)v1.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = job_target(j);
                planes[j * words + w] = eq_mask8(q[0], t) | (eq_mask8(q[1], t) << 8) | (eq_mask8(q[2], t) << 16) | (eq_mask8(q[3], t) << 24);
            }
            if (bimg) {
                const int4 b0 = *reinterpret_cast<const int4*>(bimg + (size_t)w * 32), b1 = *reinterpret_cast<const int4*>(bimg + (size_t)w * 32 + 16);
                planes[4 * words + w] = nz_mask8(make_uint2((unsigned)b0.x, (unsigned)b0.y)) | (nz_mask8(make_uint2((unsigned)b0.z, (unsigned)b0.w)) << 8) |
                                        (nz_mask8(make_uint2((unsigned)b1.x, (unsigned)b1.y)) << 16) | (nz_mask8(make_uint2((unsigned)b1.z, (unsigned)b1.w)) << 24);
            }
        }
    }
    __syncthreads();
    auto trace = [&](int j, int p, auto emit) -> int {
        const int y = p / W, x = p - y * W;
        if (use_planes) return trace_border(BitPlane{planes + j * words, H, W, wpr}, y, x, j < 4, emit);
        return trace_border(ByteImage{j < 4 ? cimg : bimg, H, W, j < 4 ? job_target(j) : -1}, y, x, j < 4, emit);
    };

    // ---- A. keep the candidates that are the first pixel of their component (their border never runs above them);
    // the point count of the survivors is parked per pixel (in the not yet used `dst` area: tissue at p, body at hw + p)
    int32_t* cnt_px = dst;
    for (int w = tid; w < njobs * words; w += kThreads) {
        unsigned word = tips[w], keep = word;
        if (!word) continue;
        const int j = w / words, wi = w - j * words;
        while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            const int p = (wi << 5) + bit;
            const int n = trace(j, p, NoEmit());
            if (n < 0) keep &= ~(1u << bit);
            else cnt_px[(j == 4 ? hw : 0) + p] = n;
        }
        tips[w] = keep;
    }
    __syncthreads();

    // ---- B. slots in the reference's order: colour by colour, last-found (highest raster index) first
    for (int j = 0; j < njobs; ++j) {
        for (int c0 = 0; c0 < words; c0 += kThreads) {
            const int wi = words - 1 - (c0 + tid);
            const unsigned word = wi >= 0 ? tips[j * words + wi] : 0u;
            const int cnt = __popc(word);
            s_scan[tid] = cnt;
            __syncthreads();
            for (int o = 1; o < kThreads; o <<= 1) {               // inclusive scan
                const int v = tid >= o ? s_scan[tid - o] : 0;
                __syncthreads();
                s_scan[tid] += v;
                __syncthreads();
            }
            int at = s_base + s_scan[tid] - cnt;
            unsigned wv = word;
            while (wv) {
                const int bit = 31 - __clz(wv);
                wv &= ~(1u << bit);
                if (at < ws.max_slots) { slot_p[at] = (wi << 5) + bit; slot_job[at] = j; }
                ++at;
            }
            __syncthreads();
            if (tid == kThreads - 1) s_base += s_scan[tid];
            __syncthreads();
        }
    }
    if (tid == 0) {
        s_nslots = s_base;
        if (s_base > ws.max_slots) { s_nslots = ws.max_slots; s_status |= 1; }
    }
    __syncthreads();
    const int nslots = s_nslots;

    // ---- C. raw point counts (parked in phase A)
    for (int sl = tid; sl < nslots; sl += kThreads) slot_cnt[sl] = cnt_px[(slot_job[sl] == 4 ? hw : 0) + slot_p[sl]];
    __syncthreads();
    // ---- D. offsets (a few hundred slots: one thread)
    if (tid == 0) {
        int off = 0;
        for (int s = 0; s < nslots; ++s) {
            if (off + slot_cnt[s] > ws.raw_cap) { s_status |= 4; s_nslots = s; break; }
            slot_off[s] = off;
            off += slot_cnt[s];
        }
    }
    __syncthreads();
    const int ns = s_nslots;
    // ---- E/F. raw points; tissue contours: arc length, approxPolyDP, closing point.  Body: every border pixel.
    for (int s = tid; s < ns; s += kThreads) {
        const int j = slot_job[s], p = slot_p[s], off = slot_off[s];
        int32_t* r = raw + off;
        const int cnt = trace(j, p, StoreEmit{r});
        if (j < 4) {
            double perimeter = 0.0;                                // cv2.arcLength(cnt, True)
            if (cnt > 1) {
                int32_t prev = r[cnt - 1];
                for (int i = 0; i < cnt; ++i) {
                    const float dx = (float)(pt_x(r[i]) - pt_x(prev)), dy = (float)(pt_y(r[i]) - pt_y(prev));
                    perimeter += (double)__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                    prev = r[i];
                }
            }
            const double eps = __dmul_rn(0.001, perimeter);
            int32_t* d = dst + off + s;
            int n = approx_closed(r, cnt, __dmul_rn(eps, eps), d, stack + off + 2 * s);
            if (n > 2 && d[0] != d[n - 1]) { d[n] = d[0]; ++n; }
            slot_cnt[s] = n;
        } else {
            // get_only_body_mask_contours: contours with < 5 points are skipped, the last of the others wins
            slot_cnt[s] = cnt >= 5 ? (r[0] == r[cnt - 1] ? cnt - 1 : cnt) : 0;
            if (cnt >= 5) atomicMax(&s_body_slot, s);
        }
    }
    __syncthreads();
    // ---- G. polygon table
    if (tid == 0) {
        int np = 0, off = 0;
        int32_t* pc = poly_cls + (size_t)b * max_polys;
        int32_t* po = poly_off + (size_t)b * (max_polys + 1);
        po[0] = 0;
        for (int s = 0; s < ns; ++s) {
            const int j = slot_job[s];
            if (j == 4 && s != s_body_slot) { slot_off[s] = -1; continue; }
            const int n = slot_cnt[s];
            if (np >= max_polys) { s_status |= 1; slot_off[s] = -1; continue; }
            if (off + n > max_points) { s_status |= 2; slot_off[s] = -1; continue; }
            pc[np] = job_class(j);
            slot_p[s] = off;                                       // output offset (the first pixel is no longer needed)
            off += n;
            po[++np] = off;
        }
        n_polys[b] = np;
        if (bimg && s_body_slot < 0) s_status |= 8;                // a body mask was given but has no outline (utils.py:1165)
        status[b] = s_status;
    }
    __syncthreads();
    // ---- H. points out: (x, y) int32
    int32_t* out = points_xy + (size_t)b * max_points * 2;
    for (int s = 0; s < ns; ++s) {
        if (slot_off[s] < 0) continue;
        const int j = slot_job[s], n = slot_cnt[s], o = slot_p[s];
        const int32_t* srcp = j < 4 ? dst + slot_off[s] + s : raw + slot_off[s];
        for (int i = tid; i < n; i += kThreads) {
            out[2 * (o + i)] = pt_x(srcp[i]);
            out[2 * (o + i) + 1] = pt_y(srcp[i]);
        }
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct PolyLayout {
    size_t reach, tips, slot_p, slot_job, slot_cnt, slot_off, raw, dst, stack, total;
    int words, max_slots, raw_cap;
};

PolyLayout poly_layout(int B, int H, int W, int max_polys) {
    PolyLayout l{};
    l.words = (int)(((size_t)H * W + 31) / 32);
    l.max_slots = max_polys + 64;                                  // the body mask's extra outlines are slots, not polygons
    l.raw_cap = 3 * H * W;                                         // pure-noise images reach ~1.5 border points per pixel
    size_t o = 0;
    l.reach = o; o += align256((size_t)B * kJobs * l.words * 4);
    l.tips = o; o += align256((size_t)B * kJobs * l.words * 4);
    l.slot_p = o; o += align256((size_t)B * l.max_slots * 4);
    l.slot_job = o; o += align256((size_t)B * l.max_slots * 4);
    l.slot_cnt = o; o += align256((size_t)B * l.max_slots * 4);
    l.slot_off = o; o += align256((size_t)B * l.max_slots * 4);
    l.raw = o; o += align256((size_t)B * l.raw_cap * 4);
    l.dst = o; o += align256((size_t)B * ((size_t)l.raw_cap + l.max_slots) * 4);
    l.stack = o; o += align256((size_t)B * ((size_t)l.raw_cap + 2 * (size_t)l.max_slots) * 8);
    l.total = o;
    return l;
}

// ---- K8 hand-over: what divide_triangles_into_groups / build_polygons_with_area do to the polygon list
// (femm_generator.py:49-60, 88-115): drop polygons with fewer than four points, close the rings, sort (stable) by
// ascending shoelace area; the class-4 outline is the outer contour and not part of the list (femm_generator.py:454-459).
__global__ void __launch_bounds__(kThreads)
poly_for_mesh_kernel(const int32_t* __restrict__ n_polys, const int32_t* __restrict__ poly_cls, const int32_t* __restrict__ poly_off,
                     const int32_t* __restrict__ points_xy, int max_polys, int max_points, double* __restrict__ out_xy,
                     int32_t* __restrict__ out_off, int32_t* __restrict__ out_cls, int32_t* __restrict__ out_n,
                     double* __restrict__ area_ws, int32_t* __restrict__ order_ws) {
    // one CTA per image
    const int b = blockIdx.x, tid = threadIdx.x;
    const int np = n_polys[b];
    const int32_t* pc = poly_cls + (size_t)b * max_polys;
    const int32_t* po = poly_off + (size_t)b * (max_polys + 1);
    const int32_t* pts = points_xy + (size_t)b * max_points * 2;
    double* area = area_ws + (size_t)b * max_polys;
    int32_t* order = order_ws + (size_t)b * max_polys;
    __shared__ int s_kept;
    for (int i = tid; i < np; i += kThreads) {
        const int o = po[i], n = po[i + 1] - o;
        double a = -1.0;                                           // dropped
        if (pc[i] != 4 && n >= 4) {
            long long acc = 0;                                     // integer coordinates: the shoelace sum is exact
            const bool closed = pts[2 * o] == pts[2 * (o + n - 1)] && pts[2 * o + 1] == pts[2 * (o + n - 1) + 1];
            const int m = closed ? n - 1 : n;                      // distinct vertices
            for (int k = 0; k < m; ++k) {
                const int k2 = k + 1 == m ? 0 : k + 1;
                acc += (long long)pts[2 * (o + k)] * pts[2 * (o + k2) + 1] - (long long)pts[2 * (o + k2)] * pts[2 * (o + k) + 1];
            }
            a = fabs(0.5 * (double)acc);
        }
        area[i] = a;
    }
    __syncthreads();
    // stable rank by (area, index): np is a few hundred at most
    for (int i = tid; i < np; i += kThreads) {
        if (area[i] < 0.0) continue;
        int r = 0;
        for (int k = 0; k < np; ++k)
            if (area[k] >= 0.0 && (area[k] < area[i] || (area[k] == area[i] && k < i))) ++r;
        order[r] = i;
    }
    if (tid == 0) {
        int kept = 0;
        for (int i = 0; i < np; ++i) kept += area[i] >= 0.0;
        s_kept = kept;
        out_n[b] = kept;
    }
    __syncthreads();
    const int kept = s_kept;
    int32_t* oo = out_off + (size_t)b * (max_polys + 1);
    int32_t* oc = out_cls + (size_t)b * max_polys;
    double* oxy = out_xy + (size_t)b * ((size_t)max_points + max_polys) * 2;
    if (tid == 0) {
        int off = 0;
        oo[0] = 0;
        for (int r = 0; r < kept; ++r) {
            const int i = order[r], o = po[i], n = po[i + 1] - o;
            const bool closed = pts[2 * o] == pts[2 * (o + n - 1)] && pts[2 * o + 1] == pts[2 * (o + n - 1) + 1];
            off += closed ? n : n + 1;
            oo[r + 1] = off;
            oc[r] = pc[i];
        }
    }
    __syncthreads();
    for (int r = 0; r < kept; ++r) {
        const int i = order[r], o = po[i], n = po[i + 1] - o, base = oo[r], m = oo[r + 1] - base;
        for (int k = tid; k < m; k += kThreads) {
            const int kk = k < n ? k : 0;                          // the closing point
            oxy[2 * (base + k)] = (double)pts[2 * (o + kk)];
            oxy[2 * (base + k) + 1] = (double)pts[2 * (o + kk) + 1];
        }
    }
}

}  // namespace

extern "C" size_t eitb_label_polygons_workspace_bytes(int B, int H, int W, int max_polys) {
    if (B <= 0 || H <= 0 || W <= 0 || max_polys <= 0) return 0;
    return poly_layout(B, H, W, max_polys).total;
}

extern "C" int eitb_label_polygons(const uint8_t* code, const uint8_t* body, int B, int H, int W, int max_polys, int max_points,
                                   int32_t* n_polys, int32_t* poly_cls, int32_t* poly_off, int32_t* points_xy, int32_t* status,
                                   void* ws, size_t ws_bytes, eitb_stream_t stream) {
    if (!code || !n_polys || !poly_cls || !poly_off || !points_xy || !status || B < 0 || H <= 0 || W <= 0 || max_polys <= 0 ||
        max_points <= 0)
        return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    if (!eitb_flood::flood_supported(H, W) || H > 32767 || W > 32767 || (size_t)B * H * W / 8 >= (1ull << 32)) return EITB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(code) | (body ? reinterpret_cast<uintptr_t>(body) : 0)) & 15) return EITB_ERR_BAD_ARG;
    const PolyLayout l = poly_layout(B, H, W, max_polys);
    if (!ws || ws_bytes < l.total) return EITB_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    PolyWs w{};
    w.reach = reinterpret_cast<uint32_t*>(base + l.reach);
    w.tips = reinterpret_cast<unsigned*>(base + l.tips);
    w.slot_p = reinterpret_cast<int32_t*>(base + l.slot_p);
    w.slot_job = reinterpret_cast<int32_t*>(base + l.slot_job);
    w.slot_cnt = reinterpret_cast<int32_t*>(base + l.slot_cnt);
    w.slot_off = reinterpret_cast<int32_t*>(base + l.slot_off);
    w.raw = reinterpret_cast<int32_t*>(base + l.raw);
    w.dst = reinterpret_cast<int32_t*>(base + l.dst);
    w.stack = reinterpret_cast<int2*>(base + l.stack);
    w.max_slots = l.max_slots; w.raw_cap = l.raw_cap;
    // outside of every external contour, per colour (background = everything that is not the colour) and for the body
    int rc = eitb_flood::frame_flood<eitb_flood::SRC_U8_NE>(code, B, H, W, 3, EITB_CODE_ADIPOSE, EITB_CODE_BONE, EITB_CODE_MUSCLE,
                                                            nullptr, w.reach, s, kJobs, 0);
    if (rc != EITB_OK) return rc;
    rc = eitb_flood::frame_flood<eitb_flood::SRC_U8_NE>(code, B, H, W, 1, EITB_CODE_LUNG, 0, 0, nullptr, w.reach, s, kJobs, 3);
    if (rc != EITB_OK) return rc;
    if (body) {
        rc = eitb_flood::frame_flood<eitb_flood::SRC_U8_EQ>(body, B, H, W, 1, 0, 0, 0, nullptr, w.reach, s, kJobs, 4);
        if (rc != EITB_OK) return rc;
    }
    if (cudaMemsetAsync(w.tips, 0, (size_t)B * kJobs * l.words * 4, s) != cudaSuccess) return EITB_ERR_LAUNCH;
    eitb_prof_begin("poly_tips_kernel", s);
    poly_tips_kernel<<<eitb_grid((long long)B * H * W / 8, 256, 8), 256, 0, s>>>(code, body, w.reach, B, H, W, w.tips, l.words);
    EITB_CHECK_LAUNCH();
    const size_t plane_bytes = (size_t)kJobs * l.words * 4;
    const int use_planes = plane_bytes <= 200 * 1024 && (H * W) % 32 == 0;
    if (use_planes && cudaFuncSetAttribute(poly_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    eitb_prof_begin("poly_build_kernel", s);
    poly_build_kernel<<<B, kThreads, use_planes ? plane_bytes : 0, s>>>(code, body, H, W, l.words, w, max_polys, max_points, n_polys,
                                                                        poly_cls, poly_off, points_xy, status, use_planes);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" size_t eitb_polygons_for_mesh_workspace_bytes(int B, int max_polys) {
    if (B <= 0 || max_polys <= 0) return 0;
    return align256((size_t)B * max_polys * 8) + align256((size_t)B * max_polys * 4);
}

extern "C" int eitb_polygons_for_mesh(const int32_t* n_polys, const int32_t* poly_cls, const int32_t* poly_off,
                                      const int32_t* points_xy, int B, int max_polys, int max_points, double* out_xy,
                                      int32_t* out_off, int32_t* out_cls, int32_t* out_n, void* ws, size_t ws_bytes,
                                      eitb_stream_t stream) {
    if (!n_polys || !poly_cls || !poly_off || !points_xy || !out_xy || !out_off || !out_cls || !out_n || B < 0 || max_polys <= 0 ||
        max_points <= 0)
        return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    if (!ws || ws_bytes < eitb_polygons_for_mesh_workspace_bytes(B, max_polys)) return EITB_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    double* area = reinterpret_cast<double*>(ws);
    int32_t* order = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + align256((size_t)B * max_polys * 8));
    eitb_prof_begin("poly_for_mesh_kernel", s);
    poly_for_mesh_kernel<<<B, kThreads, 0, s>>>(n_polys, poly_cls, poly_off, points_xy, max_polys, max_points, out_xy, out_off,
                                                out_cls, out_n, area, order);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
