python scripts/pcie_probe.py
python scripts/pcie_bands.py 8 2
python scripts/pcie_bands.py 8 0
python scripts/pcie_bands.py 4 1
python scripts/pcie_bands.py 16 3
