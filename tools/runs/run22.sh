timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t22.log 2>&1; echo "pytest rc $?" > gpurun_out/r2_rc22.txt
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke22.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2_rc22.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err; echo "bench rc $?" >> gpurun_out/r2_rc22.txt
tail -3 gpurun_out/r2_t22.log; cat gpurun_out/r2_smoke22.log | tail -2; cat gpurun_out/r2_rc22.txt
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench22.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['roofline']['traffic'])"
