# lean first tile launch + register-only pass 2 against the build before both: parity, A/B, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2cl_pytest.log 2>&1; tail -3 gpurun_out/r2cl_pytest.log
run() { VO_LIB=$PWD/$1 python scripts/time_vol.py "${@:2}" 2>&1 | tail -1; }
NEW=voroffset_b200/libvoroffset_b200.so; OLD=build/ab/libvo_base.so
for rep in 1 2; do
for spec in "torus_z 2048 0 32 dilation" "torus_z 1024 18 16 dilation" "torus_z 2048 34 32 erosion" "torus_z 2048 34 32 closing" "lattice 512 10 5 dilation" "blobs 1024 8 16 dilation" "torus_x 256 0 8 dilation"; do
  echo "== $spec"
  echo -n "old "; run $OLD $spec 20
  echo -n "new "; run $NEW $spec 20
done
done 2>&1 | tee gpurun_out/r2cl_ab.txt
for i in 1 2; do
VO_LIB=$PWD/$OLD timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/r2cl_bench.txt
