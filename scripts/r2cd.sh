# as r2bc, with the lean multi-interval launch sized by the deepest column (shared memory not asked for stays L1)
mkdir -p gpurun_out
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so
for rep in 1 2; do
for spec in "lattice 512 10 5 dilation" "lattice 256 14 12 dilation" "lattice 512 14 12 dilation" "lattice 512 10 8 dilation" "torus_z 2048 34 32 erosion" ; do
  extra=""; case "$spec" in *erosion) extra="erosion=general";; esac
  echo "== $spec"
  echo -n "old      "; run $OLD $spec 20 $extra
  echo -n "new auto "; run $NEW $spec 20 $extra
  echo -n "new col  "; run $NEW $spec 20 $extra cand_order=column
done
done 2>&1 | tee gpurun_out/r2cd_ab.txt
for w in 10 12 14 16; do echo -n "new auto multi_warps=$w "; run $NEW lattice 512 10 5 dilation 20 multi_warps=$w; done 2>&1 | tee -a gpurun_out/r2cd_ab.txt
