# usage: bash tools/sweep_dp8.sh "4 4" "4 8" ...   (bucket MB, NCCL CTA cap)
for cfg in "$@"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$((20 + RANDOM % 70)) bench.py --gpus 8 --steps 30 --warmup 3 --no-rooflines --no-parity --no-sampling --no-cpu-baseline --bucket-mb $1 --comm-ctas $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bucket_mb $1 ctas $2:', round(d['ms_per_step'],3), round(d['value'],1))"
done
