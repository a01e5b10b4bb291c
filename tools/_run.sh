set -x
mkdir -p gpurun_out; rm -f gpurun_out/bench_all.jsonl
for w in ctr128 ctr256 ecb128 ecb128dec xts256 xts256dec ocb128 cbc128dec cfb128dec; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_err.log
done
for w in gcm128 gcmsiv128 ccm128batch eax128batch siv128batch gcm128batch; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w --gib-per-gpu 4 >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_err.log
done
prof() { # name regex workload gib blocks
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o /tmp/prof_$1 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --workload $3 --gib-per-gpu $4 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --workload $3 --gib-per-gpu $4   (round 1, IDP.4A lookup addressing)" $5 > gpurun_out/r1_$1_ncu_full.txt 2>> gpurun_out/ncu_$1.log
  rm -f /tmp/prof_$1.ncu-rep
}
prof gcm_bulk_kernel gcm_bulk_kernel gcm128 4 268435456
prof xts_sectors_kernel xts_sectors_kernel xts256 16 1073741824
prof ccm_batch_kernel ccm_batch_kernel ccm128batch 4 268435456
ls -la gpurun_out
