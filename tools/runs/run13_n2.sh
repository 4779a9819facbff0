timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q --tb=short -s > gpurun_out/r2_ddp_t13.log 2>&1; echo "ddp pytest rc $?" > gpurun_out/r2_rc13.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; echo "n2 rc $?" >> gpurun_out/r2_rc13.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_n2.json 2> gpurun_out/r2_ref_n2.err; echo "ref n2 rc $?" >> gpurun_out/r2_rc13.txt
timeout 120 python bench.py --steps 30 --warmup 5 --no-cudnn-baseline > gpurun_out/r2_bench_n1_box13.json 2> /dev/null; echo "n1 rc $?" >> gpurun_out/r2_rc13.txt
grep -n "ddp\|passed\|failed" gpurun_out/r2_ddp_t13.log | head; cat gpurun_out/r2_rc13.txt
python - <<'PY'
import json
for f in ('r2_bench_n2_final', 'r2_bench_n1_box13', 'r2_ref_n2'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, round(d['value'], 1), round(d['ms_per_step'], 3), d.get('impl'), d.get('steps'))
    except Exception as e:
        print(f, 'ERR', e)
PY
