mkdir -p gpurun_out
python tools/pcie_ceiling.py > gpurun_out/pcie_ceiling.txt 2>&1
cat gpurun_out/pcie_ceiling.txt
