timeout 300 python tools/step_table.py > gpurun_out/r2_st_default.txt 2> /dev/null
STCGAN_TC_MT=2 timeout 300 python tools/step_table.py > gpurun_out/r2_st_mt2.txt 2> /dev/null
run() { label=$1; shift
  env "$@" timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  roof', round(d['roofline']['frac'], 3))" >> gpurun_out/r2_ab14.txt
}
run "default " X=1
run "TC_MT=2 " STCGAN_TC_MT=2
run "default " X=1
cat gpurun_out/r2_ab14.txt
