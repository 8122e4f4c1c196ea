# (the VO_P2REG environment switch existed for this run only; the variant it selected is now the pass2_union option, default on shallow input)
# pass 2 with a register-only running union (capacity 2, third interval -> redo launch): 12 and 16 CTAs per SM against the current kernel
mkdir -p gpurun_out
for rep in 1 2; do
for spec in "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation"; do
  echo "== $spec"
  echo -n "cur   "; python scripts/time_vol.py $spec 20 | tail -1
  echo -n "reg12 "; VO_P2REG=12 python scripts/time_vol.py $spec 20 | tail -1
  echo -n "reg16 "; VO_P2REG=16 python scripts/time_vol.py $spec 20 | tail -1
done
done 2>&1 | tee gpurun_out/r2ck_pass2_regonly.txt
