mkdir -p gpurun_out; rm -f gpurun_out/bench_xtsd2.jsonl
for r in 1 30 60; do
  echo "xts256dec share=$r" >> gpurun_out/bench_xtsd2.jsonl
  UAES_XTS_BS_PERMILLE=$r python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload xts256dec >> gpurun_out/bench_xtsd2.jsonl 2>> gpurun_out/bench_err.log
done
echo "xts256 share=1" >> gpurun_out/bench_xtsd2.jsonl
UAES_XTS_BS_PERMILLE=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload xts256 >> gpurun_out/bench_xtsd2.jsonl 2>> gpurun_out/bench_err.log
