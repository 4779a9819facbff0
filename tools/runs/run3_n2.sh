# 2 GPUs: the NCCL parity test, then the bench several times back to back (teardown must not hang), then the eager fallback
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q --tb=short -s > gpurun_out/r2_ddp_t.log 2>&1; echo "ddp pytest rc $?" > gpurun_out/r2_rc3.txt
for i in 1 2 3 4 5; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_$i.json 2> gpurun_out/r2_bench_n2_$i.err; echo "n2 run $i rc $?" >> gpurun_out/r2_rc3.txt
done
STCGAN_NCCL_IN_GRAPH=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_eager.json 2> gpurun_out/r2_bench_n2_eager.err; echo "n2 eager rc $?" >> gpurun_out/r2_rc3.txt
NCCL_DEBUG=INFO timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_dbg.json 2> gpurun_out/r2_bench_n2_dbg.err; grep -i -m5 "nvls\|algo\|proto" gpurun_out/r2_bench_n2_dbg.err >> gpurun_out/r2_rc3.txt
timeout 120 python bench.py --steps 20 --warmup 5 --no-cudnn-baseline > gpurun_out/r2_bench_n1_same_box.json 2> /dev/null; echo "n1 rc $?" >> gpurun_out/r2_rc3.txt
tail -4 gpurun_out/r2_ddp_t.log; cat gpurun_out/r2_rc3.txt
