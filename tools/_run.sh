set -x
mkdir -p gpurun_out; rm -f gpurun_out/bench_ccm.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "ccm or dropin" > gpurun_out/pytest_ccm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ccm.log
tail -30 gpurun_out/pytest_ccm.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload ccm128batch --gib-per-gpu 4 >> gpurun_out/bench_ccm.jsonl 2>> gpurun_out/bench_err.log
tail -3 gpurun_out/bench_err.log
