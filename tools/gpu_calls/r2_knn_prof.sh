#!/bin/bash
mkdir -p gpurun_out
timeout 80 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_knn|RadixSort" -s 9 -c 9 --csv --log-file gpurun_out/r2_knn_launches.csv python tools/knn_profile.py > gpurun_out/r2_knn_prof.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/r2_knn_prof.log; wc -l gpurun_out/r2_knn_launches.csv
