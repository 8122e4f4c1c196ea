for w in 8 12 16 20 24 32; do
  lib=build/lib_mw$w.so; [ $w = 16 ] && lib=voroffset_b200/libvoroffset_b200.so
  echo "== multi warps $w"
  VO_LIB=$lib python scripts/run_vol.py lattice 512 10 5 dilation 5 2>&1 | tail -1
  VO_LIB=$lib python scripts/run_vol.py torus_z 2048 34 32 erosion 4 erosion=general 2>&1 | tail -1
done
