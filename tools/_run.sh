mkdir -p gpurun_out; rm -f gpurun_out/bench_gcmmix.jsonl
for r in 0 1 2 3 4 5 8; do
  echo "clmul_rows=$r" >> gpurun_out/bench_gcmmix.jsonl
  UAES_GCM_CLMUL_ROWS=$r python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload gcm128 --gib-per-gpu 4 >> gpurun_out/bench_gcmmix.jsonl 2>> gpurun_out/bench_err.log
done
UAES_GCM_CLMUL_ROWS=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gcm" > gpurun_out/pytest_gcmmix.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gcmmix.log
tail -3 gpurun_out/pytest_gcmmix.log
