# register-only pass 2: 10 CTAs per SM (48 registers, no spills) against 12 (40 registers, 136 bytes spilled)
mkdir -p gpurun_out
for rep in 1 2 3; do
for spec in "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation"; do
  echo -n "minb12 "; python scripts/time_vol.py $spec 20 | tail -1
  echo -n "minb10 "; VO_LIB=$PWD/build/ab/libvo_p2r10.so python scripts/time_vol.py $spec 20 | tail -1
done
done 2>&1 | tee gpurun_out/r2cn_pass2_regonly_minb.txt
