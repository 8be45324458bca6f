mkdir -p gpurun_out
for sub in 262144 65536; do MCG_BENCH_VERBOSE=1 MCG_BENCH_E2E_PROFILE=1 MCG_QUERY_SUB=$sub python bench.py --no-extras --no-cpu --no-check --steps 4 --warmup 3 2>&1 >/dev/null | grep -E "e2e" | tail -3; done
