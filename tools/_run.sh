set -x
mkdir -p gpurun_out; rm -f gpurun_out/bench_bs7.jsonl
for c in "256 0" "256 170" "256 200" "256 230" "256 260" "256 300"; do set -- $c
  echo "ctr128 threads=$1 share=$2" >> gpurun_out/bench_bs7.jsonl
  UAES_CTR_THREADS=$1 UAES_CTR_BS_PERMILLE=$2 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload ctr128 >> gpurun_out/bench_bs7.jsonl 2>> gpurun_out/bench_err.log
done
