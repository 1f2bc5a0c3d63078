#!/bin/bash
# Round-2, second capture: default bench line, then ncu --set full (with source) of the kernels rewritten this round.
set -x
mkdir -p gpurun_out; R=/tmp/r2rep; mkdir -p $R
python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err || { tail -30 gpurun_out/bench_e.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_e_ref.json 2> gpurun_out/bench_e_ref.err
python profiles/run_net.py 160 > gpurun_out/r2_run_net.log 2>&1 || { tail -20 gpurun_out/r2_run_net.log; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"stem_u8|dwconv3x3|mask_decode_kernel|contour_cand3|contour_repaint|small_repaint|cc_local|area_kernel|frame_flood" \
    -f -o $R/r2b_full_own python profiles/run_net.py 160 > gpurun_out/r2b_ncu.log 2>&1
ncu -i $R/r2b_full_own.ncu-rep --page raw --csv > gpurun_out/r2b_full_own_raw.csv
for k in stem_u8 dwconv3x3 mask_decode_kernel cc_local area_kernel; do
  ncu -i $R/r2b_full_own.ncu-rep --page source --csv -k regex:"$k" > gpurun_out/r2b_src_$k.csv 2>/dev/null
done
ls -la $R gpurun_out | tail -30
du -sh gpurun_out
