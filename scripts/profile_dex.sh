set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dexelize.py -m gpu -x -q ) > gpurun_out/pytest_dex.log 2>&1
tail -5 gpurun_out/pytest_dex.log
timeout 600 python scripts/dexelize_time.py gpurun_out/dexelize_times.jsonl 2> gpurun_out/dex.err
cat gpurun_out/dexelize_times.jsonl
ncu --set full --clock-control none --import-source on -k regex:"k_dex" -s 5 -c 5 -o gpurun_out/r1e_dex python -c "
import sys; sys.path.insert(0,'.')
from voroffset_b200 import _lib, synth
from voroffset_b200.dexelize import dexelize_dev, grid_for
ctx=_lib.Context(0); V,F=synth.torus_mesh(1024,256); g=grid_for(V,None,0,2048)
for _ in range(3): dexelize_dev(ctx,V,F,g)[0].free()
" > gpurun_out/pdex.log 2>&1
tail -2 gpurun_out/pdex.log
timeout 900 python scripts/bench_configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log
