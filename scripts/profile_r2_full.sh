# Round 2: ncu --set full (with source) of the pass-1 tile kernel and pass 2 on the C5 erosion (two-hull variant) and on the
# C3 lattice (lean list launch). Run under gpurun on one GPU; summaries go to profiles/ through scripts/profile_summary.py
# and scripts/ncu_lines.py.
set -x
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows" -s 10 -c 5 -o gpurun_out/r2_ero python scripts/run_vol.py torus_z 2048 34 32 erosion 3 > gpurun_out/r2_p_ero.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile" -s 12 -c 4 -o gpurun_out/r2_c3 python scripts/run_vol.py lattice 512 10 5 dilation 4 > gpurun_out/r2_p_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh|k_scan_compact" -s 14 -c 7 -o gpurun_out/r2_c5 python scripts/run_vol.py torus_z 2048 0 32 dilation 3 > gpurun_out/r2_p_c5.log 2>&1
tail -2 gpurun_out/r2_p_ero.log gpurun_out/r2_p_c3.log gpurun_out/r2_p_c5.log
