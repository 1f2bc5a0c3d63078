"""K2 + K7 alone on bench-like label images (the random-init network's output on 64 phantom slices), for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eitsynthai_b200 import cabi, ops, synth                               # noqa: E402
from eitsynthai_b200.pipeline import ImagingPipeline                       # noqa: E402

pipe = ImagingPipeline("cuda:0")
px = torch.from_numpy(np.stack([synth.phantom_slice(s) for s in range(64)])).cuda()
body = ops.body_mask(px, 1, -1024, True)
x = pipe.window_input(px, body)
with torch.no_grad():
    head, protos = pipe._net(pipe.axial_model_512, x)
dets, _, n = ops.nms(head.contiguous(), 4, want_idx=False)
code0, _, _ = ops.mask_decode(dets, n, protos, 0)
for _ in range(2):
    ops.body_mask(px, 1, -1024, True)
    ops.label_cleanup(code0.clone(), body)
torch.cuda.synchronize()
cabi.profile_enable(True)
for _ in range(5):
    ops.body_mask(px, 1, -1024, True)
    ops.label_cleanup(code0.clone(), body)
torch.cuda.synchronize()
rep = cabi.profile_report()
cabi.profile_enable(False)
for name, (cnt, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms / 5 * 1e3:9.1f} us per 64 slices  x{cnt // 5}  {name}")
