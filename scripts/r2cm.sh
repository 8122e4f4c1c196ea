# host-buffer call: the two new defaults (20-warp first tile launch, register-only pass 2) against the previous behaviour, interleaved
mkdir -p gpurun_out
for i in 1 2; do python scripts/e2e_ab.py - pass2_union=lists tile_general=inline pass2_union=lists,tile_general=inline; done 2>&1 | tee gpurun_out/r2cm_e2e_ab.txt
