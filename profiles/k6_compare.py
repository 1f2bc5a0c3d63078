"""Time K6 on the tensor cores vs the CUDA cores on the bench workload (CNN outputs of phantom slices)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eitsynthai_b200 import ops, synth
from eitsynthai_b200.pipeline import ImagingPipeline
pipe = ImagingPipeline("cuda:0")
px = torch.from_numpy(np.stack([synth.phantom_slice(s) for s in range(64)])).cuda()
body = ops.body_mask(px, 1, -1024, True)
_, x = ops.hu_window(px, body_mask=body, want_u8=False, channels_last=True)
with torch.no_grad():
    head, protos = pipe.axial_model_512(x)
dets, _, n = ops.nms(head.contiguous(), 4, want_idx=False)
print("mean dets", float(n.float().mean()), "mean box w", float((dets[..., 2] - dets[..., 0])[dets[..., 4] > 0].mean()))
rh, rp = synth.random_heads(64, 300, seed=1)
dets300, _, n300 = ops.nms(torch.from_numpy(rh).cuda().half(), 4, want_idx=False)
p300 = torch.from_numpy(rp).cuda().half().contiguous(memory_format=torch.channels_last)
for name, d, nn, pr in (("bench-like", dets, n, protos), ("300 dets", dets300, n300, p300)):
    for label, var in (("tcgen05", 0x20), ("scalar", 0x10), ("warp-mma", 0)):
        for _ in range(3):
            ops.mask_decode(d, nn, pr, var)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ops.mask_decode(d, nn, pr, var)
        b.record(); torch.cuda.synchronize()
        print(f"{name:12s} {label:10s} {a.elapsed_time(b) / 10 * 1e3:8.1f} us per 64 slices  (mean n = {float(nn.float().mean()):.0f})")
