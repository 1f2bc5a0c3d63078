"""Freeze the reference's real-data polygon sets as test INPUTS (run in the authoring container).

    python oracle/gen_polygon_fixtures.py

``kt_service/ai_tools/mesh_tools/mesh_service_trials.py:10-322`` embeds six polygon lists exported from
real segmentations (58-111 polygons, 1.3k-16k vertices; set 6 is in millimetres and carries a class-4
body contour).  They are the only real-data fixture for the triangle labeller; no expected
labels exist.  The lists are read through the reference's own ``get_test_data()`` (its ``create_mesh``
import is stubbed) and stored as arrays in ``tests/golden/reference_polygon_sets.npz``:
``set{k}_xy`` [V,2] float64 (vertices as written, rings NOT closed/sorted), ``set{k}_off`` [P+1],
``set{k}_cls`` [P].
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/kt_service/ai_tools/mesh_tools/mesh_service_trials.py"


def main():
    for name in ("kt_service", "kt_service.ai_tools", "kt_service.ai_tools.mesh_tools"):
        sys.modules.setdefault(name, types.ModuleType(name))
    stub = types.ModuleType("kt_service.ai_tools.mesh_tools.femm_generator")
    stub.create_mesh = None
    sys.modules["kt_service.ai_tools.mesh_tools.femm_generator"] = stub
    spec = importlib.util.spec_from_file_location("ref_trials", SRC)
    mod = importlib.util.module_from_spec(spec)
    mod.__name__ = "ref_trials"                                # not "__main__": the timing loop at the bottom stays off
    spec.loader.exec_module(mod)
    sets = mod.get_test_data()
    out = {}
    for k, lines in enumerate(sets, 1):
        xy, off, cls = [], [0], []
        for line in lines:
            v = list(map(float, line.strip().split(" ")))
            cls.append(int(v[0]))
            pts = np.asarray(v[1:1 + 2 * ((len(v) - 1) // 2)], np.float64).reshape(-1, 2)
            xy.append(pts)
            off.append(off[-1] + len(pts))
        out[f"set{k}_xy"] = np.concatenate(xy)
        out[f"set{k}_off"] = np.asarray(off, np.int32)
        out[f"set{k}_cls"] = np.asarray(cls, np.int32)
        print(k, len(cls), "polygons", off[-1], "vertices", "classes", np.bincount(cls))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_polygon_sets.npz"), **out)


if __name__ == "__main__":
    main()
