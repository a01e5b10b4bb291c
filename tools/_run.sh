set -x
mkdir -p gpurun_out; rm -f gpurun_out/bench_row4.jsonl
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for w in ccm128batch eax128batch siv128batch; do
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload $w --gib-per-gpu 4 >> gpurun_out/bench_row4.jsonl 2>> gpurun_out/bench_err.log
done
tail -3 gpurun_out/bench_err.log
