timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q --tb=short -s > gpurun_out/r2_ddp_t15.log 2>&1; echo "ddp pytest rc $?" > gpurun_out/r2_rc15.txt
grep -n "ddp\|passed\|failed\|^E  " gpurun_out/r2_ddp_t15.log | head -20; cat gpurun_out/r2_rc15.txt
