timeout 1200 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2_t2.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc2.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc $?" >> gpurun_out/r2_rc2.txt
timeout 200 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2_infer2.json 2> gpurun_out/r2_infer2.err; echo "infer rc $?" >> gpurun_out/r2_rc2.txt
STCGAN_INFER_FOLD=0 timeout 200 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2_infer2_nofold.json 2> gpurun_out/r2_infer2_nofold.err; echo "infer-nofold rc $?" >> gpurun_out/r2_rc2.txt
timeout 400 python bench.py --workload train512 --steps 10 --warmup 3 > gpurun_out/r2_t512_2.json 2> gpurun_out/r2_t512_2.err; echo "t512 rc $?" >> gpurun_out/r2_rc2.txt
STCGAN_OVERLAP_REAL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cudnn-baseline > gpurun_out/r2_bench2_noreal.json 2> gpurun_out/r2_bench2_noreal.err; echo "bench-noreal rc $?" >> gpurun_out/r2_rc2.txt
python tools/one_step.py 3 > gpurun_out/r2_onestep.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file gpurun_out/r2_traffic.csv python tools/one_step.py 3 > gpurun_out/r2_ncu.log 2>&1; echo "ncu rc $?" >> gpurun_out/r2_rc2.txt
tail -3 gpurun_out/r2_t2.log; cat gpurun_out/r2_rc2.txt
