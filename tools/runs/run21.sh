timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "batchnorm" > gpurun_out/r2_t21.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc21.txt
run() { label=$1; shift
  env "$@" timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); h = d['roofline_hbm']['families']
print('$label', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms   bn_bwd_small us:', round(h['bn_bwd_small']['s'] * 1e6))" >> gpurun_out/r2_ab21.txt
}
run "small_bn=1 (512 threads)" X=1
run "small_bn=0              " STCGAN_SMALL_BN=0
run "small_bn=1 (512 threads)" X=1
cat gpurun_out/r2_ab21.txt; tail -2 gpurun_out/r2_t21.log; cat gpurun_out/r2_rc21.txt
