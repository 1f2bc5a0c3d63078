"""One YOLO11s-seg forward on 160 slices inside a profiler range (for ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eitsynthai_b200.yolo_seg import build_model
torch.backends.cudnn.benchmark = True
m = build_model(4, "cuda:0", torch.float16, seed=1)
x = torch.rand(160, 3, 512, 512, device="cuda").half().contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
