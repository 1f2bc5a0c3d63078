"""Every convolution shape of YOLO11s-seg at 512x512: K11/K12 against PyTorch (fp32 reference for the
numbers, cuDNN fp16 + the K9 epilogue pass for the time).  Writes gpurun_out/conv_layers.json.

    python profiles/conv_layers.py [--batch 160] [--quick]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eitsynthai_b200 import cabi, ops                                   # noqa: E402
from eitsynthai_b200.convnet import Act, PackedConv, conv                # noqa: E402

# (name, Cin, Cout, k, s, H (input), act, residual, groups)
LAYERS = [
    ("l1", 32, 64, 3, 2, 256, 1, 0, 1), ("l2.cv1", 64, 64, 1, 1, 128, 1, 0, 1), ("l2.b.cv1", 32, 16, 3, 1, 128, 1, 0, 1),
    ("l2.b.cv2", 16, 32, 3, 1, 128, 1, 1, 1), ("l2.cv2", 96, 128, 1, 1, 128, 1, 0, 1), ("l3", 128, 128, 3, 2, 128, 1, 0, 1),
    ("l4.cv1", 128, 128, 1, 1, 64, 1, 0, 1), ("l4.b.cv1", 64, 32, 3, 1, 64, 1, 0, 1), ("l4.b.cv2", 32, 64, 3, 1, 64, 1, 1, 1),
    ("l4.cv2", 192, 256, 1, 1, 64, 1, 0, 1), ("l5", 256, 256, 3, 2, 64, 1, 0, 1), ("l6.cv1", 256, 256, 1, 1, 32, 1, 0, 1),
    ("l6.c3k.cv12", 128, 128, 1, 1, 32, 1, 0, 1), ("l6.c3k.b", 64, 64, 3, 1, 32, 1, 1, 1), ("l6.cv2", 384, 256, 1, 1, 32, 1, 0, 1),
    ("l7", 256, 512, 3, 2, 32, 1, 0, 1), ("l8.cv1", 512, 512, 1, 1, 16, 1, 0, 1), ("l8.c3k.b", 128, 128, 3, 1, 16, 1, 1, 1),
    ("l8.cv2", 768, 512, 1, 1, 16, 1, 0, 1), ("l9.cv2", 1024, 512, 1, 1, 16, 1, 0, 1), ("l13.cv1", 768, 256, 1, 1, 32, 1, 0, 1),
    ("l13.b.cv1", 128, 64, 3, 1, 32, 1, 0, 1), ("l16.cv1", 512, 128, 1, 1, 64, 1, 0, 1), ("head.box0.0", 128, 64, 3, 1, 64, 1, 0, 1),
    ("head.mc0.0", 128, 32, 3, 1, 64, 1, 0, 1), ("head.cls.last", 128, 4, 1, 1, 64, 0, 0, 1), ("proto.cv1", 128, 128, 3, 1, 64, 1, 0, 1),
    ("proto.cv2", 128, 128, 3, 1, 128, 1, 0, 1), ("proto.cv3", 128, 32, 1, 1, 128, 1, 0, 1), ("head.cls.dw", 128, 128, 3, 1, 64, 1, 0, 128),
]


def timed(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=160)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--debug", type=int, default=0, help="eitb_conv2d_debug flags (include/eitb200.h)")
    ap.add_argument("--noact", action="store_true", help="time every layer without SiLU (epilogue cost experiment)")
    ap.add_argument("--own-only", action="store_true", help="skip the cuDNN timings")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cabi.load().eitb_conv2d_debug(args.debug)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    out = []
    for name, cin, cout, k, s, H, act, has_res, groups in LAYERS:
        if args.only and args.only not in name:
            continue
        B = 4 if args.quick else args.batch
        if args.noact:
            act = 0
        x = torch.randn((B, H, H, cin), device=dev).half()
        w = (torch.randn((cout, cin // groups, k, k), device=dev) * (1.0 / (cin // groups * k * k) ** 0.5)).half()
        bias = torch.randn((cout,), device=dev)
        Ho = (H + 2 * (k // 2) - k) // s + 1
        res = torch.randn((B, Ho, Ho, cout), device=dev).half() if has_res else None
        L = PackedConv.from_weight(w, bias, s, groups, bool(act))
        y = conv(Act(x), L, res=Act(res) if has_res else None)
        torch.cuda.synchronize()
        # fp32 reference on a sub-batch (same fp16-valued operands)
        nb = min(B, 4)
        xr = x[:nb].permute(0, 3, 1, 2).float()
        ref = F.conv2d(xr, w.float(), bias, s, k // 2, 1, groups)
        if act:
            ref = F.silu(ref)
        if has_res:
            ref = ref + res[:nb].permute(0, 3, 1, 2).float()
        got = y.buf[:nb, :, :, :cout].permute(0, 3, 1, 2).float()
        err = float((got - ref).abs().max())
        scale = float(ref.abs().max())
        # last image too (tile scheduler tail)
        xr2 = x[-1:].permute(0, 3, 1, 2).float()
        ref2 = F.conv2d(xr2, w.float(), bias, s, k // 2, 1, groups)
        if act:
            ref2 = F.silu(ref2)
        if has_res:
            ref2 = ref2 + res[-1:].permute(0, 3, 1, 2).float()
        err = max(err, float((y.buf[-1:, :, :, :cout].permute(0, 3, 1, 2).float() - ref2).abs().max()))
        rec = {"layer": name, "cin": cin, "cout": cout, "k": k, "s": s, "H": H, "B": B, "max_abs_err": err, "ref_max": scale,
               "ok": err <= 2e-3 * max(scale, 1.0) + 2e-3}
        if not args.quick:
            t_own = timed(lambda: conv(Act(x), L, res=Act(res) if has_res else None))
            xc = x.permute(0, 3, 1, 2)                                  # channels-last view
            wc = w.contiguous(memory_format=torch.channels_last)
            rc = res.permute(0, 3, 1, 2) if has_res else None

            def cudnn_path():
                yy = F.conv2d(xc, wc, None, s, k // 2, 1, groups)
                if cout % 8:
                    return yy
                return ops.conv_epilogue(yy, bias, bool(act), rc, True, None, 0) if has_res else ops.bias_act_(yy, bias, bool(act))
            t_ref = t_own if args.own_only else timed(cudnn_path)
            t_conv = t_own if args.own_only else timed(lambda: F.conv2d(xc, wc, None, s, k // 2, 1, groups))
            flops = 2.0 * B * Ho * Ho * cout * (cin // groups) * k * k
            byts = (x.numel() + y.buf.numel() + (res.numel() if has_res else 0)) * 2
            rec.update({"ms_own": t_own, "ms_cudnn_plus_k9": t_ref, "ms_cudnn_conv_only": t_conv, "tflops_own": flops / t_own / 1e9,
                        "gbs_own": byts / t_own / 1e6, "speedup_vs_cudnn_k9": t_ref / t_own})
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_layers.json", "w") as f:
        json.dump(out, f, indent=1)
    bad = [r["layer"] for r in out if not r["ok"]]
    print("FAILED:" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
