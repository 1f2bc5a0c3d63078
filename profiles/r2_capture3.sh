set -x
R=/tmp/r2rep; mkdir -p $R gpurun_out
timeout 300 python -m pytest tests/test_kt_service_gpu.py -q -x 2>&1 | tail -3
timeout 200 python profiles/conv_layers.py --own-only --only l4.cv2 > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" --launch-skip 2 --launch-count 1 -f -o $R/l4cv2 python profiles/conv_layers.py --own-only --only l4.cv2 > gpurun_out/ncu_l4.log 2>&1
ncu -i $R/l4cv2.ncu-rep --page raw --csv > gpurun_out/r2c_l4cv2_raw.csv
ncu -i $R/l4cv2.ncu-rep --page source --csv > gpurun_out/r2c_l4cv2_src.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" --launch-skip 2 --launch-count 1 -f -o $R/l2cv2 python profiles/conv_layers.py --own-only --only l2.cv2 > gpurun_out/ncu_l2.log 2>&1
ncu -i $R/l2cv2.ncu-rep --page raw --csv > gpurun_out/r2c_l2cv2_raw.csv
ncu -i $R/l2cv2.ncu-rep --page source --csv > gpurun_out/r2c_l2cv2_src.csv
ls -la gpurun_out | tail -5
