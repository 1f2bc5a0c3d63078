#!/bin/bash
# Round-2 ncu captures of one 160-slice step (K2 -> K1 -> own network -> K5 -> K6 -> K7), run under gpurun.
# Every ncu pass follows a plain run of the same command that exited 0.  Reports are exported to CSV on the box
# (gpurun brings back at most 64 MiB) and the .ncu-rep files kept only when small.
set -x
mkdir -p gpurun_out
R=/tmp/r2rep; mkdir -p $R
python profiles/run_net.py 160 > gpurun_out/r2_run_net.log 2>&1 || { tail -20 gpurun_out/r2_run_net.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches_net160.csv python profiles/run_net.py 160 > gpurun_out/r2_ncu1.log 2>&1
ncu --clock-control none --profile-from-start off \
    --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis \
    -f -o $R/r2_sections_net160 python profiles/run_net.py 160 > gpurun_out/r2_ncu2.log 2>&1
ncu -i $R/r2_sections_net160.ncu-rep --page raw --csv > gpurun_out/r2_sections_net160_raw.csv
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"contour_cand|contour_repaint|mask_decode|cc_local|area_kernel|frame_flood|small_first|small_repaint|stem_conv|dwconv|upsample2x" \
    -f -o $R/r2_full_own python profiles/run_net.py 160 > gpurun_out/r2_ncu3.log 2>&1
ncu -i $R/r2_full_own.ncu-rep --page raw --csv > gpurun_out/r2_full_own_raw.csv
ncu -i $R/r2_full_own.ncu-rep --page source --csv -k regex:"contour_cand" > gpurun_out/r2_src_contour_cand.csv 2>/dev/null
ncu -i $R/r2_full_own.ncu-rep --page source --csv -k regex:"stem_conv" > gpurun_out/r2_src_stem_conv.csv 2>/dev/null
ncu -i $R/r2_full_own.ncu-rep --page source --csv -k regex:"dwconv" > gpurun_out/r2_src_dwconv.csv 2>/dev/null
ls -la $R gpurun_out
for f in $R/*.ncu-rep; do s=$(stat -c %s $f); [ $s -lt 20000000 ] && cp $f gpurun_out/; done
du -sh gpurun_out
