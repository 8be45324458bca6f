mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"
$CMD > gpurun_out/plain.log 2>&1; echo plain=$?
ncu --set full --clock-control none --import-source on -k regex:q_fused_warp -s 2 -c 1 -o gpurun_out/r02_query_fused_kernel -f $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2=$?

