set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "ccm or eax or siv or dropin" > gpurun_out/pytest_row4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_row4.log
tail -40 gpurun_out/pytest_row4.log
