timeout 1000 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2_t1.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc1.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc $?" >> gpurun_out/r2_rc1.txt
timeout 200 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2_infer1.json 2> gpurun_out/r2_infer1.err; echo "infer rc $?" >> gpurun_out/r2_rc1.txt
timeout 300 python bench.py --workload train512 --steps 10 --warmup 3 > gpurun_out/r2_t512_1.json 2> gpurun_out/r2_t512_1.err; echo "t512 rc $?" >> gpurun_out/r2_rc1.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo "ref rc $?" >> gpurun_out/r2_rc1.txt
nproc >> gpurun_out/r2_rc1.txt; tail -3 gpurun_out/r2_t1.log; cat gpurun_out/r2_rc1.txt
