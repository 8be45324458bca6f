# usage: gpu_final_multi.sh N...   (bench at every N given, NCCL parity tests first)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py -m gpu -q $NCCL_TEST_ARGS 2>&1 | tail -3 | tee gpurun_out/r02_nccl_tests_$(nvidia-smi -L | wc -l)gpu.log
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "N=$n rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$n.json'))
print('N=$n dev ms', round(d['ms_per_step'],3), 'value M', round(d['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e M', round(d['e2e']['value']/1e6,1), d['phases_ms_per_step_rank0'], d['parity']['verdict'], 'frac', round(d['roofline']['frac'],3))
PY
done
