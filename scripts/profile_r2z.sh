# Round 2, final build: tests, smoke, both bench arms, launch list of the bench, every config, launch lists of C1 / C3 /
# C4 / C5 (dilation) and C5 erosion, --set full captures of C5 dilation and C3 dilation. One gpurun call, one GPU.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_pytest.log 2>&1
tail -3 gpurun_out/r2z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/r2z_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/b.log 2>&1
timeout 900 python scripts/bench_configs.py > gpurun_out/r2z_configs.log 2>&1; tail -3 gpurun_out/r2z_configs.log
for spec in "torus_x 256 0 8 dilation c1" "lattice 512 10 5 dilation c3" "torus_z 1024 18 16 dilation c4" "torus_z 2048 0 32 dilation c5" "torus_z 2048 34 32 erosion c5ero"; do
  set -- $spec
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_$6.csv python scripts/run_vol.py $1 $2 $3 $4 $5 3 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh|k_scan_compact|k_order" -s 12 -c 6 -o gpurun_out/r2z_c5 python scripts/run_vol.py torus_z 2048 0 32 dilation 3 > gpurun_out/r2z_p_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh|k_scan_compact" -s 14 -c 7 -o gpurun_out/r2z_c3 python scripts/run_vol.py lattice 512 10 5 dilation 3 > gpurun_out/r2z_p_c3.log 2>&1
cut -c1-600 gpurun_out/r2z_bench.json
cut -c1-400 gpurun_out/r2z_bench_ref.json
nproc; free -g | head -2
