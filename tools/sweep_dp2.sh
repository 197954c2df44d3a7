for cfg in "4 4" "4 1" "4 2" "1 2" "64 4" "4 8"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$((20 + RANDOM % 70)) bench.py --gpus 2 --steps 30 --warmup 3 --no-rooflines --no-parity --no-sampling --no-cpu-baseline --bucket-mb $1 --comm-ctas $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bucket_mb $1 ctas $2:', round(d['ms_per_step'],3), round(d['value'],1))"
done
