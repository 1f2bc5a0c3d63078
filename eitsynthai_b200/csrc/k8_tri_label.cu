// K8: per-triangle tissue labelling with the reference's polygon semantics, in fp64.
//
// Reference: process_triangle (kt_service/ai_tools/mesh_tools/femm_generator.py:118-184) over
// polygons sorted by ascending area (femm_generator.py:59-60), skipping the outer class:
//   centroid strictly inside the polygon      -> that class, stop            (:172-174)
//   area(tri ∩ poly) / area(tri) > 0.5        -> that class, stop            (:175-178)
//   area(tri ∩ poly) > best so far (> 0)      -> remember the class          (:179-181)
//   nothing                                   -> outer class                 (:162)
//
// One thread per triangle over y-slab lists of every polygon's edges (see below).  The intersection area needs no
// clipped-polygon storage: by
// Green's theorem about the centroid O the boundary of (T ∩ P) splits into
//   * the pieces of P's edges inside T  (Cyrus-Beck parameter interval [t0,t1] per edge):
//         (t1 - t0) * cross(u - O, v - O), signed by P's orientation, and
//   * the pieces of T's edges inside P:  cross(a - O, b - O) * (winding-weighted fraction of a->b),
//         fraction = wn(b) + sum(s at leaving crossings) - sum(s at entering crossings),
//     with wn the signed winding number of P about the edge's end point -- so rings that touch
//     or retrace themselves (1-px whiskers from findContours) integrate exactly like the
//     signed area of a clipped ring,
// all of which are sums over P's edges, accumulated in a fixed order (deterministic results).
#include "common.cuh"

namespace {

constexpr double kAreaNoiseFloor = 1e-9;   // relative to the triangle's area
// Every polygon's edges are binned by y into kBuckets equal slabs of its bounding box (an edge is listed in every slab
// its y-span touches, in edge order).  Only edges whose y-span meets the triangle's can contribute to any of the sums
// -- the crossing tests need an edge that straddles the query point's y, Cyrus-Beck an edge that meets the triangle's
// box -- so a triangle walks the few slabs it touches instead of the whole ring: the body-sized rings of the reference's
// polygon sets (5,000 vertices) cost a 1-2 pixel triangle tens of edges, not thousands.
constexpr int kBuckets = 64;

struct D2 { double x, y; };
__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) {
    return __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ D2 ld_pt(const double* xy, long long i) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(xy) + i);
    return D2{v.x, v.y};
}

// ws layout: [P][4] bbox (minx, miny, maxx, maxy), then [P] orientation (+1 ccw, -1 cw, 0 degenerate)
__global__ void poly_prep_kernel(const double* __restrict__ poly_xy, const int32_t* __restrict__ poly_off, int P,
                                 double* __restrict__ bbox, double* __restrict__ orient) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int o0 = poly_off[p], o1 = poly_off[p + 1];
    double mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300, a2 = 0.0;
    for (int i = o0 + lane; i < o1; i += 32) {
        const D2 u = ld_pt(poly_xy, i);
        mnx = fmin(mnx, u.x); mny = fmin(mny, u.y); mxx = fmax(mxx, u.x); mxy = fmax(mxy, u.y);
        if (i + 1 < o1) {
            const D2 v = ld_pt(poly_xy, i + 1);
            a2 = __dadd_rn(a2, cross2(u.x, u.y, v.x, v.y));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    a2 = warp_sum_d(a2);
    if (lane == 0) {
        bbox[p * 4 + 0] = mnx; bbox[p * 4 + 1] = mny; bbox[p * 4 + 2] = mxx; bbox[p * 4 + 3] = mxy;
        orient[p] = a2 > 0.0 ? 1.0 : a2 < 0.0 ? -1.0 : 0.0;
    }
}

__device__ __forceinline__ int y_bucket(double y, double miny, double inv_h) {
    const double f = __dmul_rn(__dsub_rn(y, miny), inv_h);                // monotone in y: overlapping spans share a bucket
    return f <= 0.0 ? 0 : f >= (double)(kBuckets - 1) ? kBuckets - 1 : (int)f;
}
__device__ __forceinline__ double inv_bucket_height(double miny, double maxy) {
    return maxy > miny ? __ddiv_rn((double)kBuckets, __dsub_rn(maxy, miny)) : 0.0;
}

// One CTA per polygon.  Every edge's slab span [lo, hi] is computed once (phase 1, parked in the two spare ints per vertex
// at the end of the polygon's region); thread (b, s) then owns slab b and the s-th of kSeg consecutive runs of edges: it
// counts (phase 2) and, after a prefix over (slab, run), lists (phase 4) its run's edges that touch slab b -- so every slab
// list is in edge order, exactly the order a single thread per slab would produce (the triangle kernel sums in list order).
// The body-sized ring of a real label image has ~5,000 edges: with one thread per slab this kernel took 1.05 ms on the
// reference's largest polygon set, longer than the labelling itself.
// boff [P][kBuckets + 1] offsets relative to the polygon's region, entries: region of polygon p starts at (kBuckets + 2) * poly_off[p]
constexpr int kSeg = 16;

__global__ void __launch_bounds__(kBuckets * kSeg)
poly_bucket_kernel(const double* __restrict__ poly_xy, const int32_t* __restrict__ poly_off, const double* __restrict__ bbox,
                   int32_t* __restrict__ boff, int32_t* __restrict__ entries) {
    __shared__ int cnt[kSeg][kBuckets];
    __shared__ int base[kBuckets + 1];
    const int p = blockIdx.x, tid = threadIdx.x, b = tid & (kBuckets - 1), sg = tid / kBuckets;
    const int o0 = poly_off[p], n_e = poly_off[p + 1] - 1 - o0;       // edges i -> i + 1, i in [o0, o0 + n_e)
    const double miny = bbox[p * 4 + 1], inv_h = inv_bucket_height(bbox[p * 4 + 1], bbox[p * 4 + 3]);
    int32_t* region = entries + (size_t)(kBuckets + 2) * o0;
    int32_t* span = region + (size_t)kBuckets * (n_e > 0 ? n_e : 0);   // [n_e][2]
    for (int i = tid; i < n_e; i += kBuckets * kSeg) {
        const double y0 = __ldg(poly_xy + 2 * (size_t)(o0 + i) + 1), y1 = __ldg(poly_xy + 2 * (size_t)(o0 + i) + 3);
        span[2 * i] = y_bucket(fmin(y0, y1), miny, inv_h);
        span[2 * i + 1] = y_bucket(fmax(y0, y1), miny, inv_h);
    }
    __syncthreads();
    const int seg = (n_e + kSeg - 1) / kSeg;
    const int i0 = sg * seg, i1 = min(i0 + seg, n_e);
    int n = 0;
    for (int i = i0; i < i1; ++i) n += (span[2 * i] <= b && b <= span[2 * i + 1]) ? 1 : 0;
    cnt[sg][b] = n;
    __syncthreads();
    if (tid < kBuckets) {                                            // runs of slab `tid`, in edge order
        int run = 0;
        for (int k = 0; k < kSeg; ++k) { const int c = cnt[k][tid]; cnt[k][tid] = run; run += c; }
        base[tid] = run;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < kBuckets; ++k) { const int c = base[k]; base[k] = acc; acc += c; }
        base[kBuckets] = acc;
    }
    __syncthreads();
    int32_t* bo = boff + (size_t)p * (kBuckets + 1);
    if (tid <= kBuckets) bo[tid] = base[tid];
    int32_t* dst = region + base[b] + cnt[sg][b];
    for (int i = i0; i < i1; ++i)
        if (span[2 * i] <= b && b <= span[2 * i + 1]) *dst++ = o0 + i;
}

// crossing-number test of q against edge (u, v), boundary cases left to the half-open rule
__device__ __forceinline__ int pip_edge(const D2& q, const D2& u, const D2& v) {
    if ((u.y > q.y) != (v.y > q.y)) {
        const double xi = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(v.x, u.x), __dsub_rn(q.y, u.y)), __dsub_rn(v.y, u.y)), u.x);
        return q.x < xi;
    }
    return 0;
}

// One THREAD per triangle (round 2; round 1 used a warp per triangle with lanes over the ring's edges).  With the y-slab
// lists a triangle meets a handful of edges per candidate polygon, so a warp per triangle left 28 lanes idle and spent
// its time on dependent global loads and shuffle reductions; a thread per triangle keeps 32 independent triangles in
// flight per warp and sums every integral sequentially in list order (deterministic).  The per-polygon records
// (bounding box, class, orientation) are staged in shared memory.
constexpr int kMaxSmemPolys = 1024;

__global__ void __launch_bounds__(128)
tri_label_kernel(const double* __restrict__ nodes, const int64_t* __restrict__ tri, long long T,
                 const double* __restrict__ poly_xy, const int32_t* __restrict__ poly_off,
                 const int32_t* __restrict__ poly_cls, int P, int outer_cls, const double* __restrict__ bbox,
                 const double* __restrict__ orient, const int32_t* __restrict__ boff, const int32_t* __restrict__ entries,
                 int32_t* __restrict__ cls_out) {
    extern __shared__ __align__(16) unsigned char k8sm[];
    double* sbox = reinterpret_cast<double*>(k8sm);                      // [Ps][4]
    double* sori = sbox + (size_t)min(P, kMaxSmemPolys) * 4;             // [Ps]
    int* scls = reinterpret_cast<int*>(sori + min(P, kMaxSmemPolys));    // [Ps]
    int* soff = scls + min(P, kMaxSmemPolys);                            // [Ps]
    const int Ps = min(P, kMaxSmemPolys);
    for (int i = threadIdx.x; i < Ps; i += blockDim.x) {
        sbox[i * 4] = bbox[i * 4]; sbox[i * 4 + 1] = bbox[i * 4 + 1]; sbox[i * 4 + 2] = bbox[i * 4 + 2]; sbox[i * 4 + 3] = bbox[i * 4 + 3];
        sori[i] = orient[i]; scls[i] = poly_cls[i]; soff[i] = poly_off[i];
    }
    __syncthreads();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
        D2 a = ld_pt(nodes, tri[t * 3]), b = ld_pt(nodes, tri[t * 3 + 1]), c = ld_pt(nodes, tri[t * 3 + 2]);
        double area2 = cross2(__dsub_rn(b.x, a.x), __dsub_rn(b.y, a.y), __dsub_rn(c.x, a.x), __dsub_rn(c.y, a.y));
        if (area2 < 0.0) { const D2 tmp = b; b = c; c = tmp; area2 = -area2; }       // make T ccw
        const double tri_area = area2 * 0.5;
        const D2 O = {__ddiv_rn(__dadd_rn(__dadd_rn(a.x, b.x), c.x), 3.0), __ddiv_rn(__dadd_rn(__dadd_rn(a.y, b.y), c.y), 3.0)};
        const double tminx = fmin(a.x, fmin(b.x, c.x)), tmaxx = fmax(a.x, fmax(b.x, c.x));
        const double tminy = fmin(a.y, fmin(b.y, c.y)), tmaxy = fmax(a.y, fmax(b.y, c.y));
        const D2 tv[3] = {a, b, c};

        int best = outer_cls;
        double max_inter = 0.0;
        for (int p = 0; p < P; ++p) {
            const bool sm = p < Ps;
            const int pc = sm ? scls[p] : poly_cls[p];
            if (pc == outer_cls) continue;
            const double bx0 = sm ? sbox[p * 4] : bbox[p * 4], by0 = sm ? sbox[p * 4 + 1] : bbox[p * 4 + 1];
            const double bx1 = sm ? sbox[p * 4 + 2] : bbox[p * 4 + 2], by1 = sm ? sbox[p * 4 + 3] : bbox[p * 4 + 3];
            const bool cin = O.x >= bx0 && O.x <= bx1 && O.y >= by0 && O.y <= by1;
            const bool cand = cin || (tri_area > 0.0 && tmaxx >= bx0 && tminx <= bx1 && tmaxy >= by0 && tminy <= by1);
            if (!cand) continue;
            const int o0 = sm ? soff[p] : poly_off[p];
            const double inv_h = inv_bucket_height(by0, by1);
            const int32_t* bo = boff + (size_t)p * (kBuckets + 1);
            const int32_t* ent = entries + (size_t)(kBuckets + 2) * o0;
            // ---- pass 1: centroid strictly inside?  (only edges whose y-span holds O.y can cross its ray)
            if (cin) {
                int par = 0;
                const int bq = y_bucket(O.y, by0, inv_h);
                for (int k = bo[bq]; k < bo[bq + 1]; ++k) {
                    const int i = ent[k];
                    par ^= pip_edge(O, ld_pt(poly_xy, i), ld_pt(poly_xy, i + 1));
                }
                if (par) { best = pc; break; }
            }
            if (!(tri_area > 0.0)) continue;
            // ---- pass 2: intersection area
            double sum_p = 0.0;            // pieces of P's edges inside T
            double len[3] = {0.0, 0.0, 0.0};
            int wn[3] = {0, 0, 0};         // signed winding number of P about b_e (ccw positive)
            const int bb0 = y_bucket(tminy, by0, inv_h), bb1 = y_bucket(tmaxy, by0, inv_h);
            for (int bk = bb0; bk <= bb1; ++bk)
            for (int k = bo[bk]; k < bo[bk + 1]; ++k) {
                const int i = ent[k];
                const D2 u = ld_pt(poly_xy, i), v = ld_pt(poly_xy, i + 1);
                // an edge listed in several of these slabs is taken in the first one
                if (bk > bb0 && y_bucket(fmin(u.y, v.y), by0, inv_h) < bk) continue;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const D2 q = tv[(e + 1) % 3];
                    if (pip_edge(q, u, v)) wn[e] += v.y > q.y ? 1 : -1;
                }
                if (fmax(u.x, v.x) < tminx || fmin(u.x, v.x) > tmaxx || fmax(u.y, v.y) < tminy || fmin(u.y, v.y) > tmaxy)
                    continue;
                const double wx = __dsub_rn(v.x, u.x), wy = __dsub_rn(v.y, u.y);
                double t0 = 0.0, t1 = 1.0;
                bool rej = false;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const D2 ea = tv[e], eb = tv[(e + 1) % 3];
                    const double dx = __dsub_rn(eb.x, ea.x), dy = __dsub_rn(eb.y, ea.y);
                    const double su = cross2(dx, dy, __dsub_rn(u.x, ea.x), __dsub_rn(u.y, ea.y));   // side of u
                    const double sv = cross2(dx, dy, __dsub_rn(v.x, ea.x), __dsub_rn(v.y, ea.y));   // side of v
                    const double dn = cross2(dx, dy, wx, wy);
                    // Cyrus-Beck against the half-plane left of ea->eb
                    if (dn == 0.0) {
                        if (su < 0.0) rej = true;
                    } else {
                        const double ts = __ddiv_rn(-su, dn);
                        if (dn > 0.0) t0 = fmax(t0, ts); else t1 = fmin(t1, ts);
                    }
                    // crossing of the T edge with this P edge (half-open on P's parameter)
                    if ((su > 0.0) != (sv > 0.0)) {
                        const double sc = __ddiv_rn(cross2(__dsub_rn(u.x, ea.x), __dsub_rn(u.y, ea.y), wx, wy), dn);
                        if (sc >= 0.0 && sc <= 1.0) {
                            // moving along ea->eb we enter P (ccw) when cross(w, d) > 0, i.e. dn < 0
                            len[e] = dn < 0.0 ? __dsub_rn(len[e], sc) : __dadd_rn(len[e], sc);
                        }
                    }
                }
                if (!rej && t0 < t1)
                    sum_p = __dadd_rn(sum_p, __dmul_rn(__dsub_rn(t1, t0),
                                                      cross2(__dsub_rn(u.x, O.x), __dsub_rn(u.y, O.y),
                                                             __dsub_rn(v.x, O.x), __dsub_rn(v.y, O.y))));
            }
            double total = sum_p;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const double le = __dadd_rn(len[e], (double)wn[e]);
                const D2 ea = tv[e], eb = tv[(e + 1) % 3];
                total = __dadd_rn(total, __dmul_rn(le, cross2(__dsub_rn(ea.x, O.x), __dsub_rn(ea.y, O.y),
                                                               __dsub_rn(eb.x, O.x), __dsub_rn(eb.y, O.y))));
            }
            double inter = __dmul_rn(__dmul_rn(sm ? sori[p] : orient[p], total), 0.5);
            // lower-dimensional overlaps (zero-width whiskers, touching edges) have area exactly 0
            // in GEOS; fp64 sums leave ~1e-12 px^2 of noise, which must not win "inter > 0"
            if (!(inter > __dmul_rn(kAreaNoiseFloor, tri_area))) inter = 0.0;
            if (__ddiv_rn(inter, tri_area) > 0.5) { best = pc; break; }
            if (inter > max_inter) { max_inter = inter; best = pc; }
        }
        cls_out[t] = best;
    }
}

__global__ void __launch_bounds__(256)
tri_label_raster_kernel(const double* __restrict__ nodes, const int64_t* __restrict__ tri, long long T,
                        const uint8_t* __restrict__ code, int H, int W, int outer_cls, int32_t* __restrict__ cls_out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
        const D2 a = ld_pt(nodes, tri[t * 3]), b = ld_pt(nodes, tri[t * 3 + 1]), c = ld_pt(nodes, tri[t * 3 + 2]);
        const double cx = (a.x + b.x + c.x) / 3.0, cy = (a.y + b.y + c.y) / 3.0;
        const long long px = (long long)floor(cx + 0.5), py = (long long)floor(cy + 0.5);   // pixel centres sit on integers
        int cls = outer_cls;
        if (px >= 0 && px < W && py >= 0 && py < H) {
            const int cc = code[py * W + px];
            cls = cc == EITB_CODE_BONE ? 0 : cc == EITB_CODE_MUSCLE ? 1 : cc == EITB_CODE_LUNG ? 2 : cc == EITB_CODE_ADIPOSE ? 3 : outer_cls;
        }
        cls_out[t] = cls;
    }
}

}  // namespace

// per-polygon bounding boxes and orientation (5 doubles), slab offsets (kBuckets + 1 ints) and the slab lists: an edge is
// listed at most kBuckets times, so (kBuckets + 2) ints per vertex bound every polygon's region
extern "C" size_t eitb_tri_label_workspace_bytes(int P, int V) {
    const size_t p = (size_t)(P > 0 ? P : 0), v = (size_t)(V > 0 ? V : 0);
    return p * 5 * sizeof(double) + (p * (kBuckets + 1) + v * (kBuckets + 2)) * sizeof(int32_t) + 64;
}

extern "C" int eitb_tri_label(const double* nodes_xy, int64_t n_nodes, const int64_t* tri, int64_t T,
                              const double* poly_xy, const int32_t* poly_off, const int32_t* poly_cls, int P,
                              int V, int outer_cls, int32_t* cls_out, void* ws, size_t ws_bytes, eitb_stream_t stream) {
    if (T < 0 || P < 0 || n_nodes < 0 || V < 0) return EITB_ERR_BAD_ARG;
    if (T == 0) return EITB_OK;
    if (!nodes_xy || !tri || !cls_out || (P > 0 && (!poly_xy || !poly_off || !poly_cls))) return EITB_ERR_BAD_ARG;
    if (ws_bytes < eitb_tri_label_workspace_bytes(P, V) || (P > 0 && !ws)) return EITB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(nodes_xy) & 15) || (reinterpret_cast<uintptr_t>(poly_xy) & 15)) return EITB_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    double* bbox = reinterpret_cast<double*>(ws);
    double* orient = bbox + (size_t)P * 4;
    int32_t* boff = reinterpret_cast<int32_t*>(orient + P);
    int32_t* entries = boff + (size_t)P * (kBuckets + 1);
    if (P > 0) {
        eitb_prof_begin("poly_prep_kernel", s);
        poly_prep_kernel<<<eitb_div_up(P, 4), 128, 0, s>>>(poly_xy, poly_off, P, bbox, orient);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("poly_bucket_kernel", s);
        poly_bucket_kernel<<<P, kBuckets * kSeg, 0, s>>>(poly_xy, poly_off, bbox, boff, entries);
        EITB_CHECK_LAUNCH();
    }
    const int Ps = P < kMaxSmemPolys ? P : kMaxSmemPolys;
    const size_t smem = (size_t)Ps * (5 * sizeof(double) + 2 * sizeof(int)) + 16;
    const int grid = eitb_grid(T, 128, 12);
    eitb_prof_begin("tri_label_kernel", s);
    tri_label_kernel<<<grid, 128, smem, s>>>(nodes_xy, tri, T, poly_xy, poly_off, poly_cls, P, outer_cls, bbox, orient, boff, entries, cls_out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_tri_label_raster(const double* nodes_xy, int64_t n_nodes, const int64_t* tri, int64_t T,
                                     const uint8_t* code, int H, int W, int outer_cls, int32_t* cls_out,
                                     eitb_stream_t stream) {
    if (T < 0 || n_nodes < 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if (T == 0) return EITB_OK;
    if (!nodes_xy || !tri || !code || !cls_out) return EITB_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(nodes_xy) & 15) return EITB_ERR_BAD_ARG;
    eitb_prof_begin("tri_label_raster_kernel", (cudaStream_t)stream);
    tri_label_raster_kernel<<<eitb_grid(T, 256, 8), 256, 0, (cudaStream_t)stream>>>(nodes_xy, tri, T, code, H, W, outer_cls, cls_out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
