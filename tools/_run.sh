set -x
mkdir -p gpurun_out; rm -f gpurun_out/bench_row4b.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_batch.log
tail -30 gpurun_out/pytest_batch.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload gcm128batch --gib-per-gpu 4 >> gpurun_out/bench_row4b.jsonl 2>> gpurun_out/bench_err.log
tail -3 gpurun_out/bench_err.log
