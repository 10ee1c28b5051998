set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_bwd_duo_kernel|hashgrid_bwd_grouped_kernel|hashgrid_fwd_split_kernel|mlp_tc_fwd_kernel|losses_fwd_kernel|losses_bwd_kernel|mlp_tc_bwd_duo96_kernel" --launch-skip 60 --launch-count 14 -o gpurun_out/r02f_step_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_full.log 2>&1
ls -la gpurun_out/r02f_step_kernels.ncu-rep
