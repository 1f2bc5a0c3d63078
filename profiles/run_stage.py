"""Small driver for ncu captures: runs the K2 / K6 / K7 stages on 64 phantom slices a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eitsynthai_b200 import ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
px = torch.from_numpy(np.stack([synth.phantom_slice(s) for s in range(8)])).cuda().repeat(n // 8, 1, 1).contiguous()
head, protos = synth.random_heads(8, 50, seed=3)
head = torch.from_numpy(head).cuda().half().repeat(n // 8, 1, 1).contiguous()
protos = torch.from_numpy(protos).cuda().half().repeat(n // 8, 1, 1, 1).contiguous()
act = torch.randn(n, 64, 128, 128, device="cuda", dtype=torch.half).contiguous(memory_format=torch.channels_last)
bias = torch.randn(64, device="cuda")
for it in range(4):
    ops.bias_act_(act, bias, True)
    body = ops.body_mask(px, 1, -1024, True)
    _, x = ops.hu_window(px, body_mask=body, want_u8=False)
    dets, _, k = ops.nms(head, 4, want_idx=False)
    code, _, _ = ops.mask_decode(dets, k, protos)
    ops.label_cleanup(code, body)
torch.cuda.synchronize()
print("ok", int(k.sum()))
