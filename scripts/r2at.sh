python scripts/gpu_pass1.py 2>&1 | grep "lattice\|random\|BAD"
for o in on off on off; do python scripts/run_vol.py lattice 512 10 5 dilation 6 thresh_iv=$o 2>&1 | tail -1; done
python scripts/run_vol.py lattice 256 14 12 dilation 6 thresh_iv=on 2>&1 | tail -1
python scripts/run_vol.py lattice 256 14 12 dilation 6 thresh_iv=off 2>&1 | tail -1
echo "== pass 2 min CTAs 8 / 10 / 12"
python scripts/quick_c5.py
VO_LIB=build/lib_p2m10.so python scripts/quick_c5.py
VO_LIB=build/lib_p2m12.so python scripts/quick_c5.py
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
