set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.jsonl 2> gpurun_out/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n4.jsonl 2> gpurun_out/bench_n4.err
tail -2 gpurun_out/bench_n8.err
