# loads one entry ahead of the list insertions (pool lists in pass 2, survivors in class_general): A/B against the previous build
mkdir -p gpurun_out
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so
for rep in 1 2; do
for spec in "lattice 512 10 5 dilation" "lattice 512 10 8 dilation" "lattice 512 14 12 dilation" "lattice 512 10 5 erosion" "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation" "torus_z 2048 34 32 erosion" "torus_z 2048 34 32 closing"; do
  echo "== $spec"
  echo -n "old "; run $OLD $spec 20
  echo -n "new "; run $NEW $spec 20
done
done 2>&1 | tee gpurun_out/r2ch_ab.txt
