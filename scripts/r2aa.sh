export VO_LIB=build/lib_ktrace.so
python scripts/e2e_dry.py
python scripts/e2e_dry.py band_split=1
python scripts/e2e_dry.py bands=4
python scripts/e2e_dry.py bands=16
