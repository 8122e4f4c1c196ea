# layer-major candidate order decided per tile (cand_order = auto) against forced orders and the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2cb_pytest.log 2>&1; tail -3 gpurun_out/r2cb_pytest.log
run() { VO_LIB=$PWD/$1 python scripts/run_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so
for rep in 1 2; do
  echo "== old"; run $OLD lattice 512 10 5 dilation 6; run $OLD lattice 256 14 12 dilation 6; run $OLD lattice 512 10 8 dilation 6; run $OLD torus_z 2048 34 32 erosion 6 erosion=general
  for o in auto column layer; do
    echo "== new cand_order=$o"; run $NEW lattice 512 10 5 dilation 6 cand_order=$o; run $NEW lattice 256 14 12 dilation 6 cand_order=$o; run $NEW lattice 512 10 8 dilation 6 cand_order=$o; run $NEW torus_z 2048 34 32 erosion 6 erosion=general cand_order=$o
  done
done 2>&1 | tee gpurun_out/r2cb_ab.txt
for w in 8 10 12 14 16; do echo "== new auto, multi_warps $w"; run $NEW lattice 512 10 5 dilation 6 multi_warps=$w; done 2>&1 | tee -a gpurun_out/r2cb_ab.txt
