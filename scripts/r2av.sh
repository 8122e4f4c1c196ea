for m in 6 5 8 6 8; do
  lib=build/lib_th$m.so; [ $m = 6 ] && lib=voroffset_b200/libvoroffset_b200.so
  echo "== k_thresh min blocks $m"
  VO_LIB=$lib python scripts/quick_c5.py
done
python scripts/run_vol.py torus_z 2048 34 32 erosion 5 2>&1 | tail -1
python scripts/run_vol.py lattice 512 10 5 dilation 5 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
