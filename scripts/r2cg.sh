# (the carveout option this script sets existed for this run only: every forced carve-out was slower, see profiles/r2cg_carveout.txt)
# preferred shared-memory carve-out of the small kernels around the tile kernel (does an SM reconfiguration sit in the gaps?)
mkdir -p gpurun_out
for rep in 1 2; do
for c in -1 100 75 50 0; do
  for spec in "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation"; do
    if [ "$c" = "-1" ]; then python scripts/time_vol.py $spec 30; else python scripts/time_vol.py $spec 30 carveout=$c; fi
  done
done
done 2>&1 | grep -v "^$" | tee gpurun_out/r2cg_carveout.txt
