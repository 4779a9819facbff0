timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t7.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc7.txt
run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); h = d['roofline_hbm']['families']
print('$label', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), ' bn us:', {k: round(v['s'] * 1e6) for k, v in h.items() if k.startswith('bn')})" >> gpurun_out/r2_ab7.txt
}
run "default        " X=1
run "small_bn=0     " STCGAN_SMALL_BN=0
run "bn_var=1       " STCGAN_BN_VAR=1
run "bn_var=2       " STCGAN_BN_VAR=2
run "bn_var=3       " STCGAN_BN_VAR=3
run "adam_early=0   " STCGAN_ADAM_EARLY=0
run "default again  " X=1
timeout 200 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2_infer7.json 2> gpurun_out/r2_infer7.err; echo "infer rc $?" >> gpurun_out/r2_rc7.txt
cat gpurun_out/r2_ab7.txt; tail -5 gpurun_out/r2_t7.log; cat gpurun_out/r2_rc7.txt
