timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc $?" > gpurun_out/r2_rc12.txt
timeout 200 python bench.py --workload infer --steps 20 --warmup 3 > gpurun_out/r2_infer12.json 2> gpurun_out/r2_infer12.err; echo "infer rc $?" >> gpurun_out/r2_rc12.txt
timeout 400 python bench.py --workload train512 --steps 10 --warmup 3 > gpurun_out/r2_t512_12.json 2> gpurun_out/r2_t512_12.err; echo "t512 rc $?" >> gpurun_out/r2_rc12.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke12.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2_rc12.txt
timeout 300 python tools/step_table.py > gpurun_out/r2_step_table12.txt 2> gpurun_out/r2_step_table12.err; echo "steptable rc $?" >> gpurun_out/r2_rc12.txt
python tools/one_step.py 3 > gpurun_out/r2_onestep12.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file gpurun_out/r2_traffic12.csv python tools/one_step.py 3 > gpurun_out/r2_ncu12a.log 2>&1; echo "ncu-traffic rc $?" >> gpurun_out/r2_rc12.txt
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file gpurun_out/r2_launches12.csv python tools/one_step.py 3 > /dev/null 2>&1; echo "ncu-launches rc $?" >> gpurun_out/r2_rc12.txt
ncu --set full --clock-control none --nvtx --nvtx-include "profiled_step/" -c 40 -o /tmp/r2_prof_full python tools/one_step.py 3 > gpurun_out/r2_ncu12b.log 2>&1; echo "ncu-full rc $?" >> gpurun_out/r2_rc12.txt
ncu -i /tmp/r2_prof_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw12.csv 2> /dev/null; ls -la /tmp/r2_prof_full.ncu-rep gpurun_out/ >> gpurun_out/r2_rc12.txt
cat gpurun_out/r2_rc12.txt
