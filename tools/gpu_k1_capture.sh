mkdir -p gpurun_out
CMD="python tools/microbench.py 1000 sketch"
$CMD > gpurun_out/plain_k1.log 2>&1; echo plain=$?; tail -2 gpurun_out/plain_k1.log
ncu --set full --clock-control none --import-source on -k regex:sketch_tpw -s 1 -c 1 -o gpurun_out/r02_sketch_tpw_kernel -f $CMD > gpurun_out/ncu_k1.log 2>&1; echo ncu=$?
