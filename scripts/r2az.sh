python scripts/gpu_pass1.py 2>&1 | tail -1
for o in on off on off; do echo "== need_masks $o"; python scripts/quick_c5.py need_masks=$o; done
python scripts/run_vol.py lattice 512 10 5 dilation 5 need_masks=on 2>&1 | tail -1
python scripts/run_vol.py lattice 512 10 5 dilation 5 need_masks=off 2>&1 | tail -1
python scripts/run_vol.py torus_z 2048 34 32 closing 4 need_masks=on 2>&1 | tail -1
python scripts/run_vol.py torus_z 2048 34 32 closing 4 need_masks=off 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
