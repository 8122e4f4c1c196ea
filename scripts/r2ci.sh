# (the libvo_defer*.so variants were built with -DVO_DEFER_GENERAL [-DP1_MAXWARPS_V=N], a switch that existed for this run only: the result is the GEN template parameter of k_pass1_tile)
# first launch of the tile kernel without the inline sorted-list union (complex classes -> redo list): 110 / 88 / 80 registers
# instead of 128 + spills, so 16 / 20 / 22 / 23 warps per CTA. A/B against the current build + parity of the 20-warp variant
mkdir -p gpurun_out
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
for rep in 1 2; do
for spec in "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation" "torus_z 2048 34 32 erosion" "blobs 1024 8 16 dilation"; do
  echo "== $spec"
  for v in base defer defer20 defer22 defer24; do echo -n "$v "; run build/ab/libvo_$v.so $spec 20; done
done
done 2>&1 | tee gpurun_out/r2ci_ab.txt
VO_LIB=$PWD/build/ab/libvo_defer20.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2ci_pytest_defer20.log
