mkdir -p gpurun_out; rm -f gpurun_out/bench_xtsh2.jsonl
for r in 120 130 150 160; do
  echo "xts256 share=$r" >> gpurun_out/bench_xtsh2.jsonl
  UAES_XTS_BS_PERMILLE=$r python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload xts256 >> gpurun_out/bench_xtsh2.jsonl 2>> gpurun_out/bench_err.log
done
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
