for w in 16 14 12 10; do
  echo "== multi warps $w"
  python scripts/run_vol.py lattice 512 10 5 dilation 6 multi_warps=$w 2>&1 | tail -1
done
python scripts/run_vol.py lattice 512 10 5 dilation 6 2>&1 | tail -1
python scripts/run_vol.py lattice 256 14 12 dilation 6 2>&1 | tail -1
python scripts/run_vol.py lattice 256 14 12 dilation 6 multi_warps=16 2>&1 | tail -1
python scripts/gpu_pass1.py 2>&1 | tail -1
