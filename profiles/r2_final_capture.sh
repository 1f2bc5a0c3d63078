# Final capture of round 2 (run under gpurun from the repo root): tests, default bench + reference arm, then -- only after
# those commands exited 0 without a profiler -- the ncu launch list of the same bench command and `--set full` captures
# of the dominant kernel (K11 on the 3x3 128->128 @128x128 layer) and of the rewritten K6.
set -x
R=/tmp/r2final; mkdir -p $R gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.txt; cat gpurun_out/final_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_bench_reference_arm.err
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-mesh --no-cpu-baseline --no-e2e"
timeout 300 $BENCH > gpurun_out/final_bench_short.json 2>/dev/null || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_bench.csv $BENCH > gpurun_out/final_ncu_list.log 2>&1
python profiles/summarize_launches.py gpurun_out/final_launches_bench.csv 60 > gpurun_out/final_launches_bench.txt; head -30 gpurun_out/final_launches_bench.txt
timeout 200 python profiles/conv_layers.py --own-only --only proto.cv2 --batch 320 > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" --launch-skip 2 --launch-count 1 -f -o $R/protocv2 python profiles/conv_layers.py --own-only --only proto.cv2 --batch 320 > gpurun_out/final_ncu_protocv2.log 2>&1
ncu -i $R/protocv2.ncu-rep --page raw --csv > gpurun_out/final_ncu_protocv2_raw.csv
timeout 200 python profiles/k6_compare.py > gpurun_out/final_k6_compare.txt 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mask_decode_kernel" --launch-skip 16 --launch-count 1 -f -o $R/k6 python profiles/k6_compare.py > gpurun_out/final_ncu_k6.log 2>&1
ncu -i $R/k6.ncu-rep --page raw --csv > gpurun_out/final_ncu_k6_raw.csv
ls -la gpurun_out | tail -12
