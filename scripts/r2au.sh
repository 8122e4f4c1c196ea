for m in 8 12 14 16 8 12; do
  lib=build/lib_p2m$m.so; [ $m = 8 ] && lib=voroffset_b200/libvoroffset_b200.so
  echo "== pass 2 min CTAs $m"
  VO_LIB=$lib python scripts/quick_c5.py
  VO_LIB=$lib python scripts/run_vol.py torus_z 2048 34 32 erosion 5 2>&1 | tail -1
  VO_LIB=$lib python scripts/run_vol.py lattice 512 10 5 dilation 5 2>&1 | tail -1
done
