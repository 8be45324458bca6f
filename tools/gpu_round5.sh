mkdir -p gpurun_out
python -m pytest tests/test_gpu_build.py tests/test_gpu_cli_dropin.py tests/test_gpu_cli_config0.py tests/test_gpu_nccl.py -m gpu -x -q 2>&1 | tail -3
MCG_TRACE=1 python tools/cli_c1.py 2>&1 | tail -40 > gpurun_out/cli_c1_2gpu.log; cat gpurun_out/cli_c1_2gpu.log
export CUDA_VISIBLE_DEVICES=0
MCG_TRACE=1 python bench.py --no-extras --no-cpu --no-check --steps 2 --warmup 1 2>&1 >/dev/null | grep -E "mcg finalize" | tail -6
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(q_|rs_|runs_|table_|tile_|sketch_|cta_seq|window_count|count_sizes|set_|exp_|imp_|ovp_|remove_keys|tgt_)' -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
ncu --set full --clock-control none --import-source on -k regex:q_fused_warp -s 2 -c 1 -o gpurun_out/r02_query_fused_kernel -f $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2=$?
ncu --set full --clock-control none --import-source on -k 'regex:^(rs_scatter|runs_kernel|table_build)' -s 6 -c 6 -o gpurun_out/r02_build_kernels -f $CMD > gpurun_out/ncu3.log 2>&1; echo ncu3=$?
