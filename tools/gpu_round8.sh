mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/nccl_test_8gpu.log; cat gpurun_out/nccl_test_8gpu.log
for n in 8; do MCG_BENCH_VERBOSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo rc=$?; grep "bench " gpurun_out/bench_n$n.err | tail -12; done
for p in 1 4; do MCG_SHARD_PARTS=$p python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu --no-check > gpurun_out/bench_n8_parts$p.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_n8_parts$p.json'))
print('parts $p', 'dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['phases_ms_per_step_rank0'])
"; done
