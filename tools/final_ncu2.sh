set -x
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_bwd_duo_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/r02f_mlp_bwd_duo -f python tools/prof_mlp.py tc geo 4194304 > gpurun_out/r02f_ncu_duo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hashgrid_bwd_grouped_kernel|hashgrid_fwd_split_kernel" --launch-skip 4 --launch-count 4 -o gpurun_out/r02f_hashgrid_taps -f python tools/prof_hg_taps.py > gpurun_out/r02f_ncu_hg.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
