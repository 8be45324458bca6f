# e2e step time against the number of parts a sharded call is cut into (MCG_SHARD_PARTS)
mkdir -p gpurun_out
for P in 3 4 5 6 8; do
  MCG_SHARD_PARTS=$P python bench.py --steps 5 --warmup 3 --e2e-steps 10 --no-extras --no-cpu --no-check > gpurun_out/parts_$P.json 2> gpurun_out/parts_$P.err
  python - <<PY
import json
d=json.load(open('gpurun_out/parts_$P.json'))
print('parts $P dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3))
PY
done
