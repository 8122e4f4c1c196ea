# A/B of the layer-major candidate order (lean multi-interval launches) against the previous build (build/ab/libvo_base.so)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ca_pytest.log 2>&1; tail -3 gpurun_out/r2ca_pytest.log
for rep in 1 2; do
  for lib in build/ab/libvo_base.so voroffset_b200/libvoroffset_b200.so; do
    echo "== $lib"
    for w in 16 12; do
      VO_LIB=$PWD/$lib python scripts/run_vol.py lattice 512 10 5 dilation 6 multi_warps=$w 2>&1 | tail -1
    done
    VO_LIB=$PWD/$lib python scripts/run_vol.py lattice 256 14 12 dilation 6 2>&1 | tail -1
    VO_LIB=$PWD/$lib python scripts/run_vol.py torus_z 2048 34 32 erosion 6 erosion=general 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/r2ca_ab.txt
