"""Time K1 (HU window + mask + channels-last fp16) alone: GB/s against the measured copy peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eitsynthai_b200 import ops
px = torch.randint(-1200, 1500, (320, 512, 512), dtype=torch.int16, device="cuda")
body = (torch.rand((320, 512, 512), device="cuda") > 0.3).to(torch.uint8) * 255
for cl in (True, False):
    for _ in range(5):
        ops.hu_window(px, body_mask=body, want_u8=False, channels_last=cl)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.hu_window(px, body_mask=body, want_u8=False, channels_last=cl)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"channels_last={cl}: {ms*1e3:.1f} us per 320 slices, {320*512*512*9/ms/1e6:.0f} GB/s")
