python tools/one_step.py 3 > /dev/null 2>&1 && ncu --set full --clock-control none --nvtx --nvtx-include "profiled_step/" -k regex:"adam_kernel|splitk_finish|bn_bwd_small|bn_running_update" -c 12 -o /tmp/r2_prof_b python tools/one_step.py 3 > gpurun_out/r2_ncu16.log 2>&1; echo "ncu rc $?" > gpurun_out/r2_rc16.txt
ncu -i /tmp/r2_prof_b.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw16.csv 2> /dev/null
cat gpurun_out/r2_rc16.txt; wc -c gpurun_out/r2_ncu_full_raw16.csv
