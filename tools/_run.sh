mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streaming" > gpurun_out/pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stream.log
tail -25 gpurun_out/pytest_stream.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
