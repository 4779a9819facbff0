timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_t6.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc6.txt
for i in 1 2; do
  (cd _r1snap && timeout 200 python bench.py --steps 60 --warmup 5 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r1 tree ', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))") >> gpurun_out/r2_ab6.txt
  timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2 tree ', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))" >> gpurun_out/r2_ab6.txt
done
STCGAN_ADAM_EARLY=0 timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2 tree adam_early=0', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))" >> gpurun_out/r2_ab6.txt
timeout 200 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2_infer6.json 2> gpurun_out/r2_infer6.err; echo "infer rc $?" >> gpurun_out/r2_rc6.txt
cat gpurun_out/r2_ab6.txt; tail -5 gpurun_out/r2_t6.log; cat gpurun_out/r2_rc6.txt
