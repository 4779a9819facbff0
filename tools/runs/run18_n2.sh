b() { label=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 40 --warmup 5 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms')" >> gpurun_out/r2_ab19.txt
}
b "default(ae0)" X=1
b "adam_tail=1  " STCGAN_ADAM_TAIL=1
b "default(ae0)" X=1
b "adam_tail=1  " STCGAN_ADAM_TAIL=1
cat gpurun_out/r2_ab19.txt
