VO_TRACE=1 python scripts/run_vol.py lattice 512 10 5 dilation 5 2>&1 | tail -7
VO_TRACE=1 python scripts/run_vol.py lattice 256 14 12 dilation 5 2>&1 | tail -7
VO_TRACE=1 python scripts/run_vol.py torus_z 2048 34 32 erosion 3 erosion=general 2>&1 | tail -4
VO_TRACE=1 python scripts/run_vol.py torus_x 512 20 10 dilation 3 2>&1 | tail -4
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
