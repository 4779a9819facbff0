# 8 GPUs: the scaling bench at 8 (default NCCL settings, then fewer channels), then 4; the reference arm is the driver's business
b() { # label nproc extra-env...
  label=$1; n=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r2_bench_${label}.json 2> gpurun_out/r2_bench_${label}.err
  echo "$label rc $?" >> gpurun_out/r2_rc8.txt
}
b n8_a 8 X=1
b n8_b 8 X=1
b n8_ch8 8 NCCL_MAX_NCHANNELS=8
b n8_ch16 8 NCCL_MAX_NCHANNELS=16
b n8_early0 8 STCGAN_ADAM_EARLY=0 STCGAN_EARLY_D1=0
b n4_a 4 X=1
timeout 120 python bench.py --steps 30 --warmup 5 --no-cudnn-baseline > gpurun_out/r2_bench_n1_box8.json 2> /dev/null; echo "n1 rc $?" >> gpurun_out/r2_rc8.txt
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29990 bench.py --gpus 8 --steps 3 --warmup 3 > /dev/null 2> gpurun_out/r2_nccl_dbg8.err; grep -i -m8 "nvls\|nchannels\|Connected all" gpurun_out/r2_nccl_dbg8.err > gpurun_out/r2_nccl_dbg8.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n8_*.json') + glob.glob('gpurun_out/r2_bench_n4_*.json') + ['gpurun_out/r2_bench_n1_box8.json']):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/r2_rc8.txt; cat gpurun_out/r2_nccl_dbg8.txt | cut -c1-200
