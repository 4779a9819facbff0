timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t11.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc11.txt
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc $?" >> gpurun_out/r2_rc11.txt
timeout 200 python bench.py --workload infer --steps 20 --warmup 3 > gpurun_out/r2_infer11.json 2> gpurun_out/r2_infer11.err; echo "infer rc $?" >> gpurun_out/r2_rc11.txt
timeout 400 python bench.py --workload train512 --steps 10 --warmup 3 > gpurun_out/r2_t512_11.json 2> gpurun_out/r2_t512_11.err; echo "t512 rc $?" >> gpurun_out/r2_rc11.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke11.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2_rc11.txt
timeout 300 python tools/step_table.py > gpurun_out/r2_step_table11.txt 2> gpurun_out/r2_step_table11.err; echo "steptable rc $?" >> gpurun_out/r2_rc11.txt
python tools/one_step.py 3 > gpurun_out/r2_onestep11.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file gpurun_out/r2_traffic11.csv python tools/one_step.py 3 > gpurun_out/r2_ncu11a.log 2>&1; echo "ncu-traffic rc $?" >> gpurun_out/r2_rc11.txt
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -c 44 -o gpurun_out/r2_prof_full python tools/one_step.py 3 > gpurun_out/r2_ncu11b.log 2>&1; echo "ncu-full rc $?" >> gpurun_out/r2_rc11.txt
ls -la gpurun_out/r2_prof_full.ncu-rep; tail -4 gpurun_out/r2_t11.log; cat gpurun_out/r2_rc11.txt
