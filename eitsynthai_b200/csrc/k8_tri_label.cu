// K8: per-triangle tissue labelling with the reference's polygon semantics, in fp64.
//
// Reference: process_triangle (kt_service/ai_tools/mesh_tools/femm_generator.py:118-184) over
// polygons sorted by ascending area (femm_generator.py:59-60), skipping the outer class:
//   centroid strictly inside the polygon      -> that class, stop            (:172-174)
//   area(tri ∩ poly) / area(tri) > 0.5        -> that class, stop            (:175-178)
//   area(tri ∩ poly) > best so far (> 0)      -> remember the class          (:179-181)
//   nothing                                   -> outer class                 (:162)
//
// One warp per triangle; lanes stride over polygon edges.  Candidate polygons are found 32 at
// a time by bounding-box ballot.  The intersection area needs no clipped-polygon storage: by
// Green's theorem about the centroid O the boundary of (T ∩ P) splits into
//   * the pieces of P's edges inside T  (Cyrus-Beck parameter interval [t0,t1] per edge):
//         (t1 - t0) * cross(u - O, v - O), signed by P's orientation, and
//   * the pieces of T's edges inside P:  cross(a - O, b - O) * (winding-weighted fraction of a->b),
//         fraction = wn(b) + sum(s at leaving crossings) - sum(s at entering crossings),
//     with wn the signed winding number of P about the edge's end point -- so rings that touch
//     or retrace themselves (1-px whiskers from findContours) integrate exactly like the
//     signed area of a clipped ring,
// all of which are sums over P's edges and reduce with warp shuffles in a fixed order
// (deterministic results).
#include "common.cuh"

namespace {

constexpr double kAreaNoiseFloor = 1e-9;   // relative to the triangle's area

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

// crossing-number test of q against edge (u, v), boundary cases left to the half-open rule
__device__ __forceinline__ int pip_edge(const D2& q, const D2& u, const D2& v) {
    if ((u.y > q.y) != (v.y > q.y)) {
        const double xi = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(v.x, u.x), __dsub_rn(q.y, u.y)), __dsub_rn(v.y, u.y)), u.x);
        return q.x < xi;
    }
    return 0;
}

__global__ void __launch_bounds__(256)
tri_label_kernel(const double* __restrict__ nodes, const int64_t* __restrict__ tri, long long T,
                 const double* __restrict__ poly_xy, const int32_t* __restrict__ poly_off,
                 const int32_t* __restrict__ poly_cls, int P, int outer_cls, const double* __restrict__ bbox,
                 const double* __restrict__ orient, int32_t* __restrict__ cls_out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long t = warp0; t < T; t += nwarps) {
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
        bool done = false;
        for (int p0 = 0; p0 < P && !done; p0 += 32) {
            const int pl = p0 + lane;
            bool cand = false, cin = false;
            if (pl < P && poly_cls[pl] != outer_cls) {
                const double bx0 = bbox[pl * 4], by0 = bbox[pl * 4 + 1], bx1 = bbox[pl * 4 + 2], by1 = bbox[pl * 4 + 3];
                cin = O.x >= bx0 && O.x <= bx1 && O.y >= by0 && O.y <= by1;
                cand = cin || (tri_area > 0.0 && tmaxx >= bx0 && tminx <= bx1 && tmaxy >= by0 && tminy <= by1);
            }
            unsigned m = __ballot_sync(0xffffffffu, cand);
            const unsigned mc = __ballot_sync(0xffffffffu, cin);
            while (m && !done) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const int p = p0 + l;
                const int o0 = poly_off[p], o1 = poly_off[p + 1] - 1;       // edges i -> i+1, i in [o0, o1)
                const int pc = poly_cls[p];
                // ---- pass 1: centroid strictly inside?
                if ((mc >> l) & 1u) {
                    int par = 0;
                    for (int i = o0 + lane; i < o1; i += 32) par ^= pip_edge(O, ld_pt(poly_xy, i), ld_pt(poly_xy, i + 1));
                    if (__popc(__ballot_sync(0xffffffffu, par)) & 1) { best = pc; done = true; break; }
                }
                if (!(tri_area > 0.0)) continue;
                // ---- pass 2: intersection area
                double sum_p = 0.0;            // pieces of P's edges inside T
                double len[3] = {0.0, 0.0, 0.0};
                int wn[3] = {0, 0, 0};         // signed winding number of P about b_e (ccw positive)
                for (int i = o0 + lane; i < o1; i += 32) {
                    const D2 u = ld_pt(poly_xy, i), v = ld_pt(poly_xy, i + 1);
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
                            const double s = __ddiv_rn(cross2(__dsub_rn(u.x, ea.x), __dsub_rn(u.y, ea.y), wx, wy), dn);
                            if (s >= 0.0 && s <= 1.0) {
                                // moving along ea->eb we enter P (ccw) when cross(w, d) > 0, i.e. dn < 0
                                len[e] = dn < 0.0 ? __dsub_rn(len[e], s) : __dadd_rn(len[e], s);
                            }
                        }
                    }
                    if (!rej && t0 < t1)
                        sum_p = __dadd_rn(sum_p, __dmul_rn(__dsub_rn(t1, t0),
                                                          cross2(__dsub_rn(u.x, O.x), __dsub_rn(u.y, O.y),
                                                                 __dsub_rn(v.x, O.x), __dsub_rn(v.y, O.y))));
                }
                double total = warp_sum_d(sum_p);
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const double le = __dadd_rn(warp_sum_d(len[e]), (double)warp_sum(wn[e]));
                    const D2 ea = tv[e], eb = tv[(e + 1) % 3];
                    total = __dadd_rn(total, __dmul_rn(le, cross2(__dsub_rn(ea.x, O.x), __dsub_rn(ea.y, O.y),
                                                                   __dsub_rn(eb.x, O.x), __dsub_rn(eb.y, O.y))));
                }
                double inter = __dmul_rn(__dmul_rn(orient[p], total), 0.5);
                // lower-dimensional overlaps (zero-width whiskers, touching edges) have area exactly 0
                // in GEOS; fp64 sums leave ~1e-12 px^2 of noise, which must not win "inter > 0"
                if (!(inter > __dmul_rn(kAreaNoiseFloor, tri_area))) inter = 0.0;
                if (__ddiv_rn(inter, tri_area) > 0.5) { best = pc; done = true; break; }
                if (inter > max_inter) { max_inter = inter; best = pc; }
            }
        }
        if (lane == 0) cls_out[t] = best;
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

extern "C" size_t eitb_tri_label_workspace_bytes(int P) { return (size_t)(P > 0 ? P : 0) * 5 * sizeof(double); }

extern "C" int eitb_tri_label(const double* nodes_xy, int64_t n_nodes, const int64_t* tri, int64_t T,
                              const double* poly_xy, const int32_t* poly_off, const int32_t* poly_cls, int P,
                              int outer_cls, int32_t* cls_out, void* ws, size_t ws_bytes, eitb_stream_t stream) {
    if (T < 0 || P < 0 || n_nodes < 0) return EITB_ERR_BAD_ARG;
    if (T == 0) return EITB_OK;
    if (!nodes_xy || !tri || !cls_out || (P > 0 && (!poly_xy || !poly_off || !poly_cls))) return EITB_ERR_BAD_ARG;
    if (ws_bytes < eitb_tri_label_workspace_bytes(P) || (P > 0 && !ws)) return EITB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(nodes_xy) & 15) || (reinterpret_cast<uintptr_t>(poly_xy) & 15)) return EITB_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    double* bbox = reinterpret_cast<double*>(ws);
    double* orient = bbox + (size_t)P * 4;
    if (P > 0) {
        eitb_prof_begin("poly_prep_kernel", s);
        poly_prep_kernel<<<eitb_div_up(P, 4), 128, 0, s>>>(poly_xy, poly_off, P, bbox, orient);
        EITB_CHECK_LAUNCH();
    }
    const int grid = eitb_grid(T * 32, 256, 8);
    eitb_prof_begin("tri_label_kernel", s);
    tri_label_kernel<<<grid, 256, 0, s>>>(nodes_xy, tri, T, poly_xy, poly_off, poly_cls, P, outer_cls, bbox, orient, cls_out);
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
