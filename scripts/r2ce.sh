# layer-major order out of line + 14 warps for deep lists: A/B against the previous build; C1 with the tile kernel forced
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ce_pytest.log 2>&1; tail -3 gpurun_out/r2ce_pytest.log
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so
for rep in 1 2; do
for spec in "lattice 512 10 5 dilation" "lattice 256 14 12 dilation" "lattice 512 14 12 dilation" "lattice 512 10 8 dilation" "torus_z 2048 34 32 erosion" "torus_z 2048 0 32 dilation" "lattice 512 10 5 erosion"; do
  extra=""; case "$spec" in torus*erosion) extra="erosion=general";; esac
  echo "== $spec"
  echo -n "old      "; run $OLD $spec 20 $extra
  echo -n "new auto "; run $NEW $spec 20 $extra
done
done 2>&1 | tee gpurun_out/r2ce_ab.txt
for w in 12 14 16; do for spec in "lattice 512 10 8 dilation" "lattice 512 14 12 dilation"; do echo -n "new multi_warps=$w "; run $NEW $spec 20 multi_warps=$w; done; done 2>&1 | tee -a gpurun_out/r2ce_ab.txt
echo "== C1" | tee -a gpurun_out/r2ce_ab.txt
for o in auto tile simple; do echo -n "pass1=$o "; run $NEW torus_x 256 0 8 dilation 30 pass1=$o; done 2>&1 | tee -a gpurun_out/r2ce_ab.txt
