# Round 2 (later build): ncu --set full (with source) of the resident C5 dilation and of the dual-form C5 erosion.
set -x
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh|k_scan_compact|k_order" -s 12 -c 6 -o gpurun_out/r2s_c5 python scripts/run_vol.py torus_z 2048 0 32 dilation 3 > gpurun_out/r2s_p_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows_dual|k_empty_dist|k_thresh" -s 8 -c 4 -o gpurun_out/r2s_ero python scripts/run_vol.py torus_z 2048 34 32 erosion 3 > gpurun_out/r2s_p_ero.log 2>&1
tail -2 gpurun_out/r2s_p_c5.log gpurun_out/r2s_p_ero.log
