# layer-major candidate order (tiles with three layers and more): min / median over 20 iterations per build and workload
mkdir -p gpurun_out
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so; NT=build/ab/libvo_notail2.so
for rep in 1 2; do
for spec in "lattice 512 10 5 dilation" "lattice 256 14 12 dilation" "lattice 512 14 12 dilation" "lattice 512 10 8 dilation" "torus_z 2048 34 32 erosion" ; do
  extra=""; case "$spec" in *erosion) extra="erosion=general";; esac
  echo "== $spec"
  echo -n "old      "; run $OLD $spec 20 $extra
  echo -n "new auto "; run $NEW $spec 20 $extra
  echo -n "new col  "; run $NEW $spec 20 $extra cand_order=column
  echo -n "notail2  "; run $NT $spec 20 $extra
done
done 2>&1 | tee gpurun_out/r2cc_ab.txt
