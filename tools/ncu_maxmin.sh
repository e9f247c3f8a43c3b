ABY3_BASIC_ONLY=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_maxmin_rowmajor --launch-skip 4 --launch-count 1 -o /tmp/r2h_maxmin python tools/basic_profile.py 8388605 > gpurun_out/r2h_ncu_maxmin.log 2>&1
ncu -i /tmp/r2h_maxmin.ncu-rep --page raw --csv > gpurun_out/r2h_maxmin_raw.csv 2>/dev/null
ncu -i /tmp/r2h_maxmin.ncu-rep --page source --csv > gpurun_out/r2h_maxmin_source.csv 2>/dev/null
ncu -i /tmp/r2h_maxmin.ncu-rep --page details > gpurun_out/r2h_maxmin_details.txt 2>/dev/null
ls -la gpurun_out/r2h_*
