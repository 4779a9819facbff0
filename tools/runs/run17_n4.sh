timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err; echo "n4 rc $?" > gpurun_out/r2_rc17.txt
timeout 120 python bench.py --steps 30 --warmup 5 --no-cudnn-baseline > gpurun_out/r2_bench_n1_box17.json 2> /dev/null; echo "n1 rc $?" >> gpurun_out/r2_rc17.txt
python - <<'PY'
import json
for f in ('r2_bench_n4_final', 'r2_bench_n1_box17'):
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, round(d['value'], 1), round(d['ms_per_step'], 3))
PY
cat gpurun_out/r2_rc17.txt
