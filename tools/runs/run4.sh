timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "graph_replay or autograd_modules or inference or infer or epilogue or thin_first or deferred" > gpurun_out/r2_t4.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc4.txt
timeout 200 python tools/noise_check.py fp32 > gpurun_out/r2_noise.txt 2>&1; echo "noise rc $?" >> gpurun_out/r2_rc4.txt
for ed in 1 0; do for ae in 1 0; do for orl in 1 0; do
  STCGAN_EARLY_D1=$ed STCGAN_ADAM_EARLY=$ae STCGAN_OVERLAP_REAL=$orl timeout 200 python bench.py --steps 60 --warmup 5 --no-cudnn-baseline 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('early_d1=$ed adam_early=$ae overlap_real=$orl', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))" >> gpurun_out/r2_ab.txt
done; done; done
python tools/infer_profile.py > gpurun_out/r2_inferprof.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "profiled/" --csv --log-file gpurun_out/r2_infer_launches.csv python tools/infer_profile.py > gpurun_out/r2_ncu_infer.log 2>&1; echo "ncu rc $?" >> gpurun_out/r2_rc4.txt
timeout 200 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2_infer4.json 2> gpurun_out/r2_infer4.err; echo "infer rc $?" >> gpurun_out/r2_rc4.txt
cat gpurun_out/r2_ab.txt; tail -3 gpurun_out/r2_t4.log; cat gpurun_out/r2_rc4.txt
