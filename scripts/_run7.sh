( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
python scripts/run_vol.py torus_z 1024 18 16 dilation 8 | tail -2
python scripts/run_vol.py torus_z 2048 0 32 dilation 8 | tail -2
python - <<'P'
import sys
sys.path.insert(0,'.')
from voroffset_b200 import synth, morpho, _lib
ctx=_lib.Context(0); op=morpho.make_operator("ours",ctx)
d=morpho.DeviceVolume.upload(ctx, synth.torus_z(2048))
for i in range(5):
    r,t1,t2=op.morph_dev("dilation", d, 32.0); k1,k2=ctx.last_profile(); r.free()
print("k_pass1_tile", round(k1,4), "k_pass2_rows", round(k2,4), "pass1", round(t1,4), "pass2", round(t2,4))
P
