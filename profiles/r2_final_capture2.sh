# second part of the final capture: launch list of the bench step (eager launches: ncu lists the nodes of a replayed CUDA
# graph only as one entry, so the same step runs with --no-graphs), per-shape convolution table, helper tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kt_service_gpu.py -x -q -m gpu 2>&1 | tail -3
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-mesh --no-cpu-baseline --no-e2e --no-graphs"
timeout 300 $BENCH > gpurun_out/final_bench_short_eager.json 2>/dev/null || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_bench.csv $BENCH > gpurun_out/final_ncu_list.log 2>&1
python profiles/summarize_launches.py gpurun_out/final_launches_bench.csv 60 > gpurun_out/final_launches_bench.txt; head -40 gpurun_out/final_launches_bench.txt
timeout 300 python profiles/conv_shapes.py --batch 320 > gpurun_out/final_conv_shapes.log 2>&1; head -12 gpurun_out/conv_shapes.txt
