"""Where the convolution time of one pass goes: every K11 launch of the axial network (and the rib network) tagged with
its layer shape and configuration by the library's launch profiler (eitb_conv2d_debug bit 20), timed with CUDA events
around each launch over an eager pass of ``--batch`` slices.

    python profiles/conv_shapes.py [--batch 320] [--reps 3]   ->  gpurun_out/conv_shapes.txt
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eitsynthai_b200 import cabi, convnet                                # noqa: E402
from eitsynthai_b200.pipeline import ImagingPipeline                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=320)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--debug", type=int, default=0)
    args = ap.parse_args()
    pipe = ImagingPipeline("cuda:0")
    x = torch.randint(0, 255, (args.batch, 512, 512), dtype=torch.uint8, device="cuda:0")
    net = pipe.axial_model_512
    for _ in range(2):
        net(x, gray=True)
    torch.cuda.synchronize()
    cabi.load().eitb_conv2d_debug(args.debug | (1 << 20))
    cabi.profile_enable(True)
    for _ in range(args.reps):
        net(x, gray=True)
    torch.cuda.synchronize()
    rep = cabi.profile_report()
    cabi.profile_enable(False)
    cabi.load().eitb_conv2d_debug(0)
    convnet.STATS = None
    rows = sorted(((ms / args.reps, cnt // args.reps, name) for name, (cnt, ms) in rep.items()), reverse=True)
    total = sum(r[0] for r in rows)
    lines = [f"# one forward of the axial network, batch {args.batch}: {total:.3f} ms over {sum(r[1] for r in rows)} own launches "
             f"(CUDA events around every launch, eager)"]
    for ms, cnt, name in rows:
        extra = ""
        if name.startswith("conv_tc_kernel:"):
            parts = name.split(":")
            k, s_ = int(parts[1][1]), int(parts[1][3])
            cin, rest = parts[2].split("->")
            cout, dims = rest.split("@")
            H, W, N = (int(v) for v in dims.split("x"))
            cin, cout = int(cin), int(cout)
            Ho, Wo = (H + s_ - 1) // s_, (W + s_ - 1) // s_
            flops = 2.0 * N * Ho * Wo * cout * cin * k * k
            byts = 2.0 * N * (H * W * cin + Ho * Wo * cout * (2 if "res" in parts[-1] else 1))
            per = ms / max(cnt, 1)
            extra = f"  {flops / per / 1e9:8.0f} TFLOP/s  {byts / per / 1e6:7.0f} GB/s  ({per * 1e3:7.1f} us each)"
        lines.append(f"{ms:8.3f} ms {100 * ms / total:5.1f}%  x{cnt:<3d} {name}{extra}")
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/conv_shapes.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
