# usage: bash scripts/run_weak.sh N   -- the driver's multi-GPU bench command, then a longer device-only run
N=$1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $R --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/weak_n$N.log 2>&1
[ -n "$LONG" ] && timeout 300 $R --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --no-e2e > gpurun_out/weak_n${N}_long.log 2>&1
grep -c metric gpurun_out/weak_n$N.log gpurun_out/weak_n${N}_long.log
