# dual erosion: 32-bit class masks in k_pass2_rows_dual, tile minima of the empty-column distances (skip the per-column scan inside
# a solid), one-compare window test in the row kernels: parity + A/B against the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2co_pytest.log 2>&1; tail -3 gpurun_out/r2co_pytest.log
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so
for rep in 1 2; do
for spec in "torus_z 2048 34 32 erosion" "torus_z 2048 34 32 closing" "torus_z 1024 18 16 erosion" "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation"; do
  echo "== $spec"
  echo -n "old "; run $OLD $spec 20
  echo -n "new "; run $NEW $spec 20
done
done 2>&1 | tee gpurun_out/r2co_ab.txt
