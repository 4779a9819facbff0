b() { # label nproc extra-env...
  label=$1; n=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r2_bench_${label}.json 2> gpurun_out/r2_bench_${label}.err
  echo "$label rc $?" >> gpurun_out/r2_rc10.txt
}
b n8_final_a 8 X=1
b n8_nodeep 8 STCGAN_DEEP_BUCKET=0
b n8_final_b 8 X=1
timeout 120 python bench.py --steps 30 --warmup 5 --no-cudnn-baseline > gpurun_out/r2_bench_n1_box10.json 2> /dev/null; echo "n1 rc $?" >> gpurun_out/r2_rc10.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n8_final*.json') + ['gpurun_out/r2_bench_n8_nodeep.json', 'gpurun_out/r2_bench_n1_box10.json']):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/r2_rc10.txt
