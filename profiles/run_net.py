"""One YOLO11s-seg forward on 160 slices through libeitb200's own network executor (K11/K12/K10) inside a
profiler range (for ncu --profile-from-start off), followed by one K2 -> K1 -> K5 -> K6 -> K7 pass.

    python profiles/run_net.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from eitsynthai_b200 import ops, synth
from eitsynthai_b200.convnet import ConvNet
from eitsynthai_b200.yolo_seg import build_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
m = build_model(4, "cuda:0", torch.float16, seed=1)
net = ConvNet(m)
px = torch.from_numpy(np.stack([synth.phantom_slice(s) for s in range(8)])).cuda().repeat(B // 8, 1, 1).contiguous()
head_r, protos_r = synth.random_heads(8, 50, seed=3)
head_r = torch.from_numpy(head_r).cuda().half().repeat(B // 8, 1, 1).contiguous()
protos_r = torch.from_numpy(protos_r).cuda().half().repeat(B // 8, 1, 1, 1).contiguous(memory_format=torch.channels_last)


def step():
    body = ops.body_mask(px, 1, -1024, True)
    _, x = ops.hu_window(px, body_mask=body, want_u8=False, channels_last=True)
    head, protos = net(x, gray=True)
    dets, _, k = ops.nms(head_r, 4, want_idx=False)
    code, _, _ = ops.mask_decode(dets, k, protos_r)
    ops.label_cleanup(code, body)
    return k


with torch.no_grad():
    for _ in range(3):
        k = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    k = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", int(k.sum()))
