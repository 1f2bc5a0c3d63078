"""Mirror of the labelling half of the reference's kt_service/ai_tools/mesh_tools/femm_generator.py.

``create_mesh`` keeps the reference signature (femm_generator.py:369-371) and return value
``(img, {'NODES', 'TRIANGLES', 'CLASS'})``.  Mesh *generation* is Gmsh's job in the reference
(:445-478) and out of the hot path: pass ``mesh=(nodes_xy, triangles)`` (e.g. from Gmsh), or let
the jittered-grid Delaunay stand-in build one when Gmsh is not importable.  Everything from
``divide_triangles_into_groups`` on (:12-184, CLASS of :187-265) runs in libeitb200 (K8).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .... import host, ops

logger = logging.getLogger(__name__)


def _device():
    from ... import config
    return torch.device(config.device())


def find_outer_contour(polygons, distance_threshold=0.1):
    """femm_generator.py:553-625, class-'4' fast path (:588-590); the Shapely union fallback is not mirrored."""
    return host.find_outer_index(polygons)


def divide_triangles_into_groups(contours, outer_contour_class, outer_contour=None, skin_width=1, *, nodes_xy=None,
                                 triangles=None, device_polygons=None):
    """femm_generator.py:12-85 with the mesh passed explicitly instead of read from Gmsh's global
    state.  Returns ``{class_id: [element index, ...]}``; pops short contours from the caller's list
    like the reference (:49-56)."""
    k = -1
    for _ in range(len(contours)):
        k += 1
        if k <= len(contours) - 1 and len(contours[k]) < 9:
            contours.pop(k)
            k -= 1
    cls = label_triangles(nodes_xy, triangles, contours, outer_contour_class, device_polygons)
    groups = {}
    for i, c in enumerate(cls.tolist()):
        groups.setdefault(c, []).append(i)
    return groups


def label_triangles(nodes_xy, triangles, contours, outer_contour_class=4, device_polygons=None) -> np.ndarray:
    """process_triangle for every element (femm_generator.py:118-184): int32 class per triangle.
    ``device_polygons``: K13's result for this label image (``ops.LabelPolygons``): the polygon list then goes
    K13 -> eitb_polygons_for_mesh -> K8 without leaving the device, instead of being parsed back from the strings."""
    dev = _device()
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    if device_polygons is not None:
        xy, off, pc, n = ops.polygons_for_mesh(device_polygons)
        P = int(n[0])
        V = int(off[0, P])
        xy, off, pc = xy[0, :V].contiguous(), off[0, :P + 1].contiguous(), pc[0, :P].contiguous()
    else:
        xy, off, pc = host.prepare_polygons(contours)
        xy, off, pc = t(xy, np.float64), t(off, np.int32), t(pc, np.int32)
    out = ops.tri_label(t(nodes_xy, np.float64), t(triangles, np.int64), xy, off, pc, int(outer_contour_class))
    return out.cpu().numpy()


#: class ids as the segmentation writes them (utils.py:498-505: the polygon strings' first field) ...
LABEL_CLASS_NAMES = {0: "bone", 1: "muscles", 2: "lung", 3: "adipose", 4: "skin"}
#: ... and as the FEMM model generator reads them (femm_tools/model_generator.py:13).  Ids 2 and 3 are swapped between
#: the two tables in the reference; ``export_mesh_for_femm`` keeps the reference's behaviour (ids pass through
#: unchanged) unless ``femm_class_order=True`` asks for the ids the consumer's table means.
FEMM_CLASS_NAMES = {0: "bone", 1: "muscles", 2: "fat", 3: "lung", 4: "skin"}


def export_mesh_for_femm(filename, nodes_xy, triangles, classes, isSaveToFile=False, femm_class_order=False):
    """femm_generator.py:187-265: used-node compaction (sorted tags -> 0-based) and the CLASS vector,
    without the O(T^2) tag search.  ``femm_class_order``: translate lung / adipose (2 / 3 of utils.py:498-505) to
    the ids femm_tools/model_generator.py:13 gives them (3 / 2); off by default, like the reference."""
    tri = np.asarray(triangles, np.int64)
    used = np.unique(tri)
    remap = np.full(int(used.max()) + 1 if used.size else 0, -1, np.int64)
    remap[used] = np.arange(used.size)
    if femm_class_order:
        swap = {2: 3, 3: 2}
        classes = [swap.get(int(c), int(c)) for c in classes]
    data = {"NODES": np.asarray(nodes_xy, np.float64)[used].tolist(), "TRIANGLES": remap[tri].tolist(),
            "CLASS": [int(c) for c in classes]}
    if isSaveToFile is True and filename:
        with open(filename, "w") as f:
            f.write("# NODES\n")
            for i, (x, y) in enumerate(data["NODES"], 1):
                f.write(f"{i} {x:.12f} {y:.12f}\n")
            f.write("\n# TRIANGLES\n")
            for (a, b, c), k in zip(data["TRIANGLES"], data["CLASS"]):
                f.write(f"{a + 1} {b + 1} {c + 1} {k}\n")
    return data


def create_mesh(pixel_spacing, polygons, lc=7, distance_threshold=1.3, skin_width=1, is_show_inner_contours=False,
                show_meshing_result_method="opencv", number_of_showed_class=-1, is_saving_to_file=False,
                export_filename=None, mesh=None, device_polygons=None):
    """femm_generator.py:369-491.  ``mesh=(nodes_xy [Nn,2], triangles [T,3])`` supplies the Gmsh
    output; the skin buffer polygon (add_skin, Shapely) and the OpenCV rendering are not mirrored
    (``img`` is None)."""
    img, mesh_data = [], []
    try:
        polygons = list(polygons)
        outer = find_outer_contour(polygons, distance_threshold)
        contours = host.parse_contours(polygons, outer)
        outer_class = int(float(polygons[outer].split(" ")[0])) if outer is not None else 4
        if mesh is None:
            from ....synth import delaunay_mesh
            if outer is None:
                raise ValueError("no outer (class 4) contour and no mesh given")
            o = np.asarray(list(map(float, polygons[outer].strip().split(" ")))[1:], np.float64).reshape(-1, 2)
            inside = _inside_tester(o)
            mesh = delaunay_mesh((o[:, 0].min(), o[:, 1].min(), o[:, 0].max(), o[:, 1].max()), float(lc), 0, inside)
        nodes_xy, triangles = mesh
        groups = divide_triangles_into_groups(contours, outer_class, None, skin_width, nodes_xy=nodes_xy,
                                              triangles=triangles, device_polygons=device_polygons if outer_class == 4 else None)
        cls = np.empty(len(triangles), np.int64)
        for c, idx in groups.items():
            cls[idx] = c
        mesh_data = export_mesh_for_femm(export_filename, nodes_xy, triangles, cls, is_saving_to_file)
        img = None
    except Exception as e:
        logger.error(f"create_mesh failed: {e}")
    return img, mesh_data


def _inside_tester(ring):
    def inside(x, y):
        res = np.zeros(len(x), bool)
        for (ux, uy), (vx, vy) in zip(ring, np.roll(ring, -1, axis=0)):
            cond = (uy > y) != (vy > y)
            with np.errstate(divide="ignore", invalid="ignore"):
                xi = (vx - ux) * (y - uy) / (vy - uy) + ux
            res ^= cond & (x < xi)
        return res
    return inside
