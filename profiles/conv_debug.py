"""Experiments on single K11 layers: halo mode on/off (and its descriptor variant), correctness against an
fp32 reference and time.   python profiles/conv_debug.py [layer]"""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eitsynthai_b200 import cabi
from eitsynthai_b200.convnet import Act, PackedConv, conv

lib = cabi.load()
dev = torch.device("cuda:0")


def timed(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


SHAPES = {"l2.cv1": (64, 64, 1, 1, 128), "l2.b.cv1": (32, 16, 3, 1, 128), "mc0": (128, 32, 3, 1, 64), "box0": (128, 64, 3, 1, 64),
          "proto.cv2": (128, 128, 3, 1, 128), "l5": (256, 256, 3, 2, 64), "l8.b": (128, 128, 3, 1, 16)}
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, (cin, cout, k, s, H) in SHAPES.items():
    if only and only != name:
        continue
    x = torch.randn((160, H, H, cin), device=dev).half()
    w = (torch.randn((cout, cin, k, k), device=dev) / (cin * k * k) ** 0.5).half()
    bias = torch.randn((cout,), device=dev)
    L = PackedConv.from_weight(w, bias, s, 1, True)
    ref = F.silu(F.conv2d(x[:2].permute(0, 3, 1, 2).float(), w.float(), bias, s, k // 2)).permute(0, 2, 3, 1)
    for dbg in ((0, 4, 16, 20, 32, 48) if k == 3 else (0, 32)):
        lib.eitb_conv2d_debug(dbg)
        y = conv(Act(x), L)
        torch.cuda.synchronize()
        err = float((y.buf[:2, ..., :cout].float() - ref).abs().max())
        t = timed(lambda: conv(Act(x), L))
        flops = 2.0 * 160 * (H // s) ** 2 * cout * cin * k * k
        print(json.dumps({"layer": name, "dbg": dbg, "ms": round(t, 4), "max_err": round(err, 4), "ok": err < 0.02 or bool(dbg & 0x707),
                          "tflops": round(flops / t / 1e9), "gbs": round((x.numel() + y.buf.numel()) * 2 / t / 1e6)}), flush=True)
lib.eitb_conv2d_debug(0)
