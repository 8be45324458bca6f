mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"
$CMD > gpurun_out/plain.log 2>&1; echo plain=$?
ncu --set full --clock-control none --import-source on -k 'regex:^(runs_kernel|table_build|rs_hist|tgt_coarse)' -c 8 -o gpurun_out/r02_build_kernels -f $CMD > gpurun_out/ncu3.log 2>&1; echo ncu3=$?
