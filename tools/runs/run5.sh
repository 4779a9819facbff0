for i in 1 2; do
  (cd _r1snap && timeout 200 python bench.py --steps 60 --warmup 5 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r1 tree ', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))") >> gpurun_out/r2_ab5.txt
  timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r2 tree ', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))" >> gpurun_out/r2_ab5.txt
done
timeout 200 python tools/noise_check.py fp32 > gpurun_out/r2_noise5.txt 2>&1
timeout 200 python tools/noise_check.py bf16 >> gpurun_out/r2_noise5.txt 2>&1
python tools/one_step.py 3 > gpurun_out/r2_onestep5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file gpurun_out/r2_launches_r2tree.csv python tools/one_step.py 3 > /dev/null 2>&1
(cd _r1snap && python tools/one_step.py 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file ../gpurun_out/r2_launches_r1tree.csv python tools/one_step.py 3 > /dev/null 2>&1)
cat gpurun_out/r2_ab5.txt; cat gpurun_out/r2_noise5.txt | cut -c1-300
