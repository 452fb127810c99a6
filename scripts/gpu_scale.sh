# weak scaling on the GPUs of one box: N = 1 first, then the full node (run under `gpurun --gpus N`)
set -x
mkdir -p gpurun_out
N=${N:-8}
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-profile-pass > gpurun_out/scale_1.log 2>&1; tail -1 gpurun_out/scale_1.log | cut -c1-160
for n in ${NS:-2 4 8}; do
  if [ "$n" -le "$N" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-profile-pass > gpurun_out/scale_$n.log 2>&1
    tail -1 gpurun_out/scale_$n.log | cut -c1-160
  fi
done
