/*
 * CPU oracle for the per-triangle tissue labeller (TEST INFRASTRUCTURE ONLY; never linked
 * into the product library).
 *
 * Restates process_triangle of the reference
 * (kt_service/ai_tools/mesh_tools/femm_generator.py:118-184) over polygons already sorted by
 * ascending area (femm_generator.py:59-60), with the Shapely 2.1.2 / GEOS predicates it calls
 * restated from their published semantics (Shapely is not installed here):
 *   Polygon.contains(Point)        -> point strictly in the interior (crossing number)
 *   tri.intersection(poly).area    -> area of the Sutherland-Hodgman clip of the polygon ring
 *                                     against the (convex) triangle, shoelace formula
 *   triangle centroid              -> mean of the three vertices
 * PARITY UNPINNED against GEOS itself: the reference holds no expected labels
 * (SURVEY.md §4, §8c).  The formulation is deliberately different from the CUDA kernel's
 * (explicit clipped vertex lists here, boundary integrals there).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/liboracle_tri.so oracle/tri_label.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { double x, y; } pt;

static int contains(const double* ring, int n, pt q) { /* ring: n closed vertices (last == first) */
    int in = 0;
    for (int i = 0; i + 1 < n; ++i) {
        const double ux = ring[2 * i], uy = ring[2 * i + 1], vx = ring[2 * i + 2], vy = ring[2 * i + 3];
        if ((uy > q.y) != (vy > q.y)) {
            const double xi = (vx - ux) * (q.y - uy) / (vy - uy) + ux;
            if (q.x < xi) in = !in;
        }
    }
    return in;
}

static double shoelace(const pt* p, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const pt a = p[i], b = p[(i + 1) % n];
        s += a.x * b.y - a.y * b.x;
    }
    return 0.5 * s;
}

/* clip subject polygon (n open vertices) against the half-plane left of a->b */
static int clip_halfplane(const pt* in, int n, pt a, pt b, pt* out) {
    int m = 0;
    const double dx = b.x - a.x, dy = b.y - a.y;
    for (int i = 0; i < n; ++i) {
        const pt p = in[i], q = in[(i + 1) % n];
        const double sp = dx * (p.y - a.y) - dy * (p.x - a.x);
        const double sq = dx * (q.y - a.y) - dy * (q.x - a.x);
        if (sp >= 0.0) out[m++] = p;
        if ((sp >= 0.0) != (sq >= 0.0)) {
            const double t = sp / (sp - sq);
            pt r = {p.x + t * (q.x - p.x), p.y + t * (q.y - p.y)};
            out[m++] = r;
        }
    }
    return m;
}

static double intersection_area(const double* ring, int n, pt a, pt b, pt c, pt* buf0, pt* buf1) {
    int m = n - 1; /* drop the closing vertex */
    for (int i = 0; i < m; ++i) { buf0[i].x = ring[2 * i]; buf0[i].y = ring[2 * i + 1]; }
    const double ring_area = shoelace(buf0, m);
    m = clip_halfplane(buf0, m, a, b, buf1); if (m == 0) return 0.0;
    m = clip_halfplane(buf1, m, b, c, buf0); if (m == 0) return 0.0;
    m = clip_halfplane(buf0, m, c, a, buf1); if (m == 0) return 0.0;
    const double s = shoelace(buf1, m);
    return ring_area >= 0.0 ? s : -s;
}

/* returns 0, or -1 on allocation failure */
int oracle_tri_label(const double* nodes_xy, const int64_t* tri, int64_t T, const double* poly_xy,
                     const int32_t* poly_off, const int32_t* poly_cls, int P, int outer_cls, int32_t* cls_out) {
    int vmax = 0;
    for (int p = 0; p < P; ++p) if (poly_off[p + 1] - poly_off[p] > vmax) vmax = poly_off[p + 1] - poly_off[p];
    pt* buf0 = (pt*)malloc(sizeof(pt) * (size_t)(2 * vmax + 16));
    pt* buf1 = (pt*)malloc(sizeof(pt) * (size_t)(2 * vmax + 16));
    if (!buf0 || !buf1) { free(buf0); free(buf1); return -1; }
    for (int64_t t = 0; t < T; ++t) {
        pt a = {nodes_xy[2 * tri[3 * t]], nodes_xy[2 * tri[3 * t] + 1]};
        pt b = {nodes_xy[2 * tri[3 * t + 1]], nodes_xy[2 * tri[3 * t + 1] + 1]};
        pt c = {nodes_xy[2 * tri[3 * t + 2]], nodes_xy[2 * tri[3 * t + 2] + 1]};
        double a2 = (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
        if (a2 < 0.0) { pt tmp = b; b = c; c = tmp; a2 = -a2; }
        const double tri_area = 0.5 * a2;
        const pt ctr = {(a.x + b.x + c.x) / 3.0, (a.y + b.y + c.y) / 3.0};
        int best = outer_cls;
        double max_inter = 0.0;
        for (int p = 0; p < P; ++p) {
            if (poly_cls[p] == outer_cls) continue;
            const double* ring = poly_xy + 2 * (size_t)poly_off[p];
            const int n = poly_off[p + 1] - poly_off[p];
            if (contains(ring, n, ctr)) { best = poly_cls[p]; break; }
            if (!(tri_area > 0.0)) continue;            /* inter / 0 raises in the reference -> skipped */
            double inter = intersection_area(ring, n, a, b, c, buf0, buf1);
            /* GEOS returns an empty (area 0.0) intersection for lower-dimensional overlaps such as
             * the zero-width whiskers findContours rings contain; clipping in fp64 leaves ~1e-12
             * px^2 of noise there, which must not win the "inter > max_intersection" test */
            if (!(inter > 1e-9 * tri_area)) inter = 0.0;
            if (inter / tri_area > 0.5) { best = poly_cls[p]; break; }
            if (inter > max_inter) { max_inter = inter; best = poly_cls[p]; }
        }
        cls_out[t] = best;
    }
    free(buf0); free(buf1);
    return 0;
}


/* Decision margins (sensitivity report): for every triangle the smallest distance of any quantity
 * process_triangle compared on its way to the label from the value at which the comparison flips, all made
 * dimensionless with the triangle's area / size:
 *   - centroid to the nearest edge of every polygon tested with contains(), divided by sqrt(triangle area);
 *   - |inter / area - 0.5| of every intersection ratio tested against 0.5;
 *   - inter / area when it is positive (how far the "inter > max_intersection" winner is from an empty intersection),
 *     and |inter - max_intersection| / area against the running maximum.
 * A triangle whose margin is >= 1e-6 is decided identically by any fp64 implementation of the same predicates
 * (GEOS included): the restatement's own rounding is ~1e-12.  margin_out[t] = that minimum (1e30: nothing was compared). */
static double seg_dist(pt q, double ux, double uy, double vx, double vy) {
    const double dx = vx - ux, dy = vy - uy, L = dx * dx + dy * dy;
    double t = L > 0.0 ? ((q.x - ux) * dx + (q.y - uy) * dy) / L : 0.0;
    if (t < 0.0) t = 0.0;
    if (t > 1.0) t = 1.0;
    const double ex = ux + t * dx - q.x, ey = uy + t * dy - q.y;
    return sqrt(ex * ex + ey * ey);
}

int oracle_tri_margins(const double* nodes_xy, const int64_t* tri, int64_t T, const double* poly_xy,
                       const int32_t* poly_off, const int32_t* poly_cls, int P, int outer_cls, double* margin_out) {
    int vmax = 0;
    for (int p = 0; p < P; ++p) if (poly_off[p + 1] - poly_off[p] > vmax) vmax = poly_off[p + 1] - poly_off[p];
    pt* buf0 = (pt*)malloc(sizeof(pt) * (size_t)(2 * vmax + 16));
    pt* buf1 = (pt*)malloc(sizeof(pt) * (size_t)(2 * vmax + 16));
    if (!buf0 || !buf1) { free(buf0); free(buf1); return -1; }
    for (int64_t t = 0; t < T; ++t) {
        pt a = {nodes_xy[2 * tri[3 * t]], nodes_xy[2 * tri[3 * t] + 1]};
        pt b = {nodes_xy[2 * tri[3 * t + 1]], nodes_xy[2 * tri[3 * t + 1] + 1]};
        pt c = {nodes_xy[2 * tri[3 * t + 2]], nodes_xy[2 * tri[3 * t + 2] + 1]};
        double a2 = (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
        if (a2 < 0.0) { pt tmp = b; b = c; c = tmp; a2 = -a2; }
        const double tri_area = 0.5 * a2, size = sqrt(tri_area > 0.0 ? tri_area : 1.0);
        const pt ctr = {(a.x + b.x + c.x) / 3.0, (a.y + b.y + c.y) / 3.0};
        double margin = 1e30, max_inter = 0.0;
        for (int p = 0; p < P; ++p) {
            if (poly_cls[p] == outer_cls) continue;
            const double* ring = poly_xy + 2 * (size_t)poly_off[p];
            const int n = poly_off[p + 1] - poly_off[p];
            double d = 1e30;
            for (int i = 0; i + 1 < n; ++i) {
                const double e = seg_dist(ctr, ring[2 * i], ring[2 * i + 1], ring[2 * i + 2], ring[2 * i + 3]);
                if (e < d) d = e;
            }
            if (d / size < margin) margin = d / size;
            if (contains(ring, n, ctr)) break;
            if (!(tri_area > 0.0)) continue;
            double inter = intersection_area(ring, n, a, b, c, buf0, buf1);
            if (!(inter > 1e-9 * tri_area)) inter = 0.0;
            const double r = inter / tri_area;
            if (fabs(r - 0.5) < margin) margin = fabs(r - 0.5);
            if (r > 0.5) break;
            if (inter > 0.0) {
                if (r < margin) margin = r;
                if (max_inter > 0.0 && fabs(inter - max_inter) / tri_area < margin) margin = fabs(inter - max_inter) / tri_area;
            }
            if (inter > max_inter) max_inter = inter;
        }
        margin_out[t] = margin;
    }
    free(buf0); free(buf1);
    return 0;
}
