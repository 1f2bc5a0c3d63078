"""K8 alone on the reference's largest polygon set (mesh_service_trials.py set 6: 110 polygons, 15 k vertices) over a
200 k-triangle Delaunay mesh: the workload of bench.py's mesh_labelling.reference_set6, for ncu captures."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eitsynthai_b200 import host, ops, synth                             # noqa: E402

z6 = np.load(os.path.join(ROOT, "tests", "golden", "reference_polygon_sets.npz"))
rx, ro, rc = z6["set6_xy"], z6["set6_off"], z6["set6_cls"]
cont = [[float(rc[p_])] + rx[ro[p_]:ro[p_ + 1]].reshape(-1).tolist() for p_ in range(len(rc))]
outer6 = next((i_ for i_, c_ in enumerate(cont) if int(c_[0]) == 4), None)
xy6, off6, cls6 = host.prepare_polygons([c_ for i_, c_ in enumerate(cont) if i_ != outer6])
lo, hi = xy6.min(0), xy6.max(0)
pitch = float(np.sqrt((hi[0] - lo[0]) * (hi[1] - lo[1]) * 2 / 200000.0))
nodes6, tris6 = synth.delaunay_mesh((lo[0], lo[1], hi[0], hi[1]), pitch, seed=0)
dm = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (nodes6, tris6, xy6, off6, cls6)]
for _ in range(3):
    out = ops.tri_label(*dm)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = ops.tri_label(*dm)
b.record()
torch.cuda.synchronize()
print(f"{len(tris6)} triangles, {len(cls6)} polygons, {len(xy6)} vertices: {a.elapsed_time(b) / 5:.3f} ms per call; classes {torch.bincount(out, minlength=5).tolist()}")
