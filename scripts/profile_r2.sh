# Round-2 launch lists (ncu --metrics gpu__time_duration.sum: per-launch times, cold-cache and serialised) of the resident
# operations, and --set full captures of the kernels VERDICT r1 names. Run under gpurun on one GPU.
set -x
for spec in "torus_z 2048 0 32 dilation c5_dilation" "torus_z 2048 34 32 erosion c5_erosion" "lattice 512 10 5 dilation c3_dilation" "torus_z 1024 18 16 dilation c4_dilation"; do
  set -- $spec
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$6.csv python scripts/run_vol.py $1 $2 $3 $4 $5 3 > gpurun_out/r2_run_$6.log 2>&1
done
