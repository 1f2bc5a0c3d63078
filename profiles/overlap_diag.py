"""Diagnostic for the two-deep host path of SeriesBatchRunner: per pass and chunk, how many label pixels differ from the
single-stream runner's result for the same series, and whether they match another pass's result instead."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from eitsynthai_b200 import synth
from eitsynthai_b200.pipeline import ImagingPipeline, SeriesBatchRunner, SeriesMeta

pipe = ImagingPipeline("cuda:0")
N, CH = 48, 16
base = synth.phantom_series(N, seed=31)[0]
vols = [base, np.roll(base, 37, axis=2).copy(), np.roll(base, -53, axis=1).copy()]
inst = synth.phantom_series(N, seed=31)[1]
hosts = [torch.from_numpy(v[None]).pin_memory() for v in vols]
plain = SeriesBatchRunner(pipe, [SeriesMeta(inst)], N, 512, chunk=CH, overlap=False)
fast = SeriesBatchRunner(pipe, [SeriesMeta(inst)], N, 512, chunk=CH, overlap=True)
def input_sensitive(r):
    plain_stage = r.cnn_stage
    def stage(x, out=None):
        head, protos = plain_stage(x, out=out)
        flip = (x.reshape(x.shape[0], -1)[:, ::997].sum(1, dtype=torch.int64) & 1) == 1   # parity of a strided checksum
        protos.mul_(torch.where(flip, 1.0, -1.0).to(protos.dtype)[:, None, None, None])
        return head, protos
    r.cnn_stage = stage
for r in (plain, fast):
    input_sensitive(r)
    r.load(hosts[0]); r.capture(warm=1)
want = []
for h in hosts:
    out = torch.zeros((1, N, 512, 512), dtype=torch.uint8).pin_memory()
    plain.step_host(h, out); want.append(out.clone())
mode = sys.argv[1] if len(sys.argv) > 1 else "two"
outs = [torch.zeros((1, N, 512, 512), dtype=torch.uint8).pin_memory() for _ in range(6)]
hs = []
for k in range(6):
    hs.append(fast.submit_host(hosts[k % 3], outs[k]))
    if mode == "one":
        fast.wait_host(hs[k])
    elif k >= 1:
        fast.wait_host(hs[k - 1])
fast.wait_host(hs[5]); fast.join(); torch.cuda.synchronize()
for k in range(6):
    for ci, (a, b) in enumerate(fast.bounds):
        d = [(outs[k][0, a:b] != want[j][0, a:b]).sum().item() for j in range(3)]
        bad_slices = [int(i) for i in torch.nonzero((outs[k][0, a:b] != want[k % 3][0, a:b]).flatten(1).any(1)).flatten()]
        print(f"pass {k} (series {k % 3}) chunk {ci}: differing px vs series 0/1/2 = {d}  bad slices {bad_slices}")
