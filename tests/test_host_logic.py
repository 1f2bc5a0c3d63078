"""Host-side rules of the product (eitsynthai_b200/host.py) against the oracle's statements, CPU only."""
import json
import os

import numpy as np
import pytest

from eitsynthai_b200 import host
from oracle import imaging as O
from oracle import tri_label as TL
from oracle import yolo_post as Y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("h,w,imgsz", [(320, 512, 640), (512, 512, 512), (256, 256, 256), (40, 512, 640), (301, 512, 640),
                                       (700, 512, 640), (512, 300, 640)])
def test_letterbox_and_scale_boxes_geometry(h, w, imgsz):
    assert host.letterbox_geometry(h, w, imgsz) == Y.letterbox_geometry(h, w, imgsz)
    nh, nw, top, bottom, left, right = host.letterbox_geometry(h, w, imgsz)
    gain, px, py = host.scale_boxes_params((nh + top + bottom, nw + left + right), (h, w))
    import torch
    b = torch.tensor([[10.0, 20.0, 200.0, 300.0]])
    want = Y.scale_boxes((nh + top + bottom, nw + left + right), b, (h, w))
    got = ((b - torch.tensor([px, py, px, py])) / gain)
    got[:, [0, 2]] = got[:, [0, 2]].clamp(0, w)
    got[:, [1, 3]] = got[:, [1, 3]].clamp(0, h)
    assert torch.equal(got, want)


@pytest.mark.parametrize("pp,iop,po", [("HFS", [1, 0, 0, 0, 1, 0], None), ("FFS", [1, 0, 0, 0, 1, 0], None),
                                       ("FFS", [-1, 0, 0, 0, -1, 0], ["L", "P"]), ("HFP", [1, 0, 0, 0, -1, 0], ["L", "A"]),
                                       ("HFS", [-1, 0, 0, 0, 1, 0], ["L", "P"])])
def test_front_geometry_equals_oracle_rows(pp, iop, po):
    rng = np.random.default_rng(0)
    vol = rng.integers(-1000, 1000, (9, 16, 24)).astype(np.int16)
    row, fx, fz = host.front_geometry(16, pp, iop, po)
    rows = vol[:, row, :]
    rows = rows[:, ::-1] if fx else rows
    rows = rows[::-1] if fz else rows
    assert np.array_equal(rows, O.front_rows(vol, pp, iop, po))


def test_instance_order_is_stable():
    inst = np.array([3, 1, 2, 1, 3, 2])
    assert host.instance_order(inst).tolist() == [1, 3, 2, 5, 0, 4]


def test_polygon_preparation_matches_oracle():
    with open(os.path.join(ROOT, "tests", "golden", "reference_polygons.json")) as f:
        polys = json.load(f)
    for tag, lst in polys.items():
        strs = lst[2:]
        outer = host.find_outer_index(strs)
        assert outer == next((i for i, s in enumerate(strs) if isinstance(s, str) and s[:1] == "4"), None)
        a = host.prepare_polygons(host.parse_contours(strs, outer))
        b = TL.prepare_polygons(TL.parse_contours(strs, outer))
        for x, y in zip(a, b[:3]):
            assert np.array_equal(x, y), tag
    xy, off, cls = host.prepare_polygons([[1, 0, 0, 1, 1], [2.0, 0, 0, 4, 0, 4, 4, 0, 4]])     # first is too short
    assert off.tolist() == [0, 5] and cls.tolist() == [2] and np.array_equal(xy[0], xy[-1])
    assert host.prepare_polygons([])[1].tolist() == [0]


def test_tri_label_oracle_c_vs_python_on_a_real_polygon_set():
    """The C restatement of process_triangle against the independent pure-Python one, on real polygons."""
    from eitsynthai_b200 import synth
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_polygon_sets.npz"))
    xy, off, cls = z["set1_xy"], z["set1_off"], z["set1_cls"]
    contours = [[float(cls[p])] + xy[off[p]:off[p + 1]].reshape(-1).tolist() for p in range(len(cls))]
    pxy, poff, pcls, _ = TL.prepare_polygons(contours)
    nodes, tri = synth.delaunay_mesh((40, 60, 470, 440), 22.0, seed=2)
    assert np.array_equal(TL.label_triangles(nodes, tri, pxy, poff, pcls), TL.label_triangles_py(nodes, tri, pxy, poff, pcls))


def test_config_surface_keeps_the_reference_names():
    """kt_service_config.py:1-13 and ai_fsi_config.toml:1-9 of the reference: same attribute / key names."""
    from eitsynthai_b200.kt_service import config, kt_service_config as c
    for name in ("ribs_segm_model", "axial_slice_segm_model_256", "axial_slice_segm_model_512"):
        assert getattr(c, name).endswith("_best.pt")
    assert c.service_version == "1.0" and c.save_log_path == ["ai_logs"] and c.device == ""
    cfg = config.load()
    assert set(cfg["main_settings"]) >= {"service_version", "save_log_path"}
    assert set(cfg["ai_settings"]) >= {"device", "weights_ribs", "weights_segmentation"}
    assert config.device() == "cuda:0"


def test_tri_label_oracle_against_exact_rational_arithmetic_on_the_most_delicate_triangles():
    """oracle/tri_label.c (fp64, with its 1e-9 * area noise floor) against process_triangle evaluated in exact rational
    arithmetic (no rounding, no noise-floor rule), on the triangles of two real polygon sets whose decisions are closest
    to flipping; and the sensitivity report: almost no triangle comes within 1e-6 of a decision boundary."""
    from eitsynthai_b200 import synth
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_polygon_sets.npz"))
    for s, pitch in ((1, 6.0), (6, 6.0)):
        xy, off, cls = z[f"set{s}_xy"], z[f"set{s}_off"], z[f"set{s}_cls"]
        nodes, tri = synth.delaunay_mesh((xy[:, 0].min(), xy[:, 1].min(), xy[:, 0].max(), xy[:, 1].max()), pitch, seed=s)
        m = TL.decision_margins(nodes, tri, xy, off, cls)
        assert (m < 1e-6).mean() < 1e-3
        idx = np.argsort(m)[:60]
        assert np.array_equal(TL.label_triangles_exact(nodes, tri[idx], xy, off, cls), TL.label_triangles(nodes, tri[idx], xy, off, cls))
    # a degenerate overlap: the triangle touches the polygon along an edge only (area exactly 0 in rational arithmetic)
    sq = np.array([[0, 0], [10, 0], [10, 10], [0, 10], [0, 0]], np.float64)
    nodes = np.array([[10, 2], [14, 5], [10, 8], [3, 3], [6, 3], [4, 7]], np.float64)
    tri = np.array([[0, 1, 2], [3, 4, 5]], np.int64)
    args = (nodes, tri, sq, np.array([0, 5], np.int32), np.array([2], np.int32))
    assert TL.label_triangles_exact(*args).tolist() == [4, 2] == TL.label_triangles(*args).tolist()


def test_export_mesh_for_femm_matches_the_reference_statement(tmp_path):
    """export_mesh_for_femm of the mirror against a literal restatement of femm_generator.py:187-265 working on
    Gmsh-style arrays (1-based node tags with unused nodes, element tags, class groups and its O(T^2) tag search):
    same dictionary, same file bytes."""
    from eitsynthai_b200.kt_service.ai_tools.mesh_tools import femm_generator as FG
    rng = np.random.default_rng(3)
    n_nodes, T = 40, 55
    coords = rng.uniform(0, 100, (n_nodes, 2))
    tri = rng.integers(0, n_nodes - 5, (T, 3))                    # the last five nodes stay unused
    cls = rng.integers(0, 5, T)
    # --- the reference's procedure, on gmsh-like inputs
    node_tags = np.arange(1, n_nodes + 1)
    node_dict = {int(t): (coords[i, 0], coords[i, 1]) for i, t in enumerate(node_tags)}
    elem_tags = np.arange(101, 101 + T)
    class_groups = {}
    for e, c in zip(elem_tags, cls):
        class_groups.setdefault(int(c), []).append(int(e))
    triangle_data, used = [], set()
    for i in range(T):
        n1, n2, n3 = (int(v) + 1 for v in tri[i])
        used.update([n1, n2, n3])
        cid = next(c for c, tags in class_groups.items() if int(elem_tags[i]) in tags)
        triangle_data.append((n1, n2, n3, cid))
    tag_to_index = {t: i + 1 for i, t in enumerate(sorted(used))}
    want = {"NODES": [[float(node_dict[t][0]), float(node_dict[t][1])] for t in sorted(used)],
            "TRIANGLES": [[tag_to_index[a] - 1, tag_to_index[b] - 1, tag_to_index[c] - 1] for a, b, c, _ in triangle_data],
            "CLASS": [int(float(c)) for *_, c in triangle_data]}
    lines = ["# NODES\n"] + [f"{tag_to_index[t]} {node_dict[t][0]:.12f} {node_dict[t][1]:.12f}\n" for t in sorted(used)]
    lines += ["\n# TRIANGLES\n"] + [f"{tag_to_index[a]} {tag_to_index[b]} {tag_to_index[c]} {k}\n" for a, b, c, k in triangle_data]
    # --- the mirror
    path = tmp_path / "mesh.txt"
    got = FG.export_mesh_for_femm(str(path), coords, tri, cls, True)
    assert got == want
    assert path.read_text() == "".join(lines)
    # the lung / fat id swap between utils.py:498-505 and femm_tools/model_generator.py:13 is explicit, off by default
    swapped = FG.export_mesh_for_femm(None, coords, tri, cls, False, femm_class_order=True)["CLASS"]
    assert swapped == [{2: 3, 3: 2}.get(int(c), int(c)) for c in cls]
    assert FG.LABEL_CLASS_NAMES[2] == "lung" and FG.FEMM_CLASS_NAMES[2] == "fat"
