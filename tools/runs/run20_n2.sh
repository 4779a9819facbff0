timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload train512 --steps 10 --warmup 3 > gpurun_out/r2_t512_n2.json 2> gpurun_out/r2_t512_n2.err; echo "t512 n2 rc $?" > gpurun_out/r2_rc20.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_t512_n2.json').read().strip().splitlines()[-1]); print('train512 N=2', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms', d['config']['global_batch'])
PY
cat gpurun_out/r2_rc20.txt
