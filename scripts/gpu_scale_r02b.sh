# 8 GPUs of one box (run under `gpurun --gpus 8`): config 4 (full step) and config 5 (distance sweep), each bounded by `timeout`
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --workload full_step --steps 5 --warmup 3 > gpurun_out/r2e_full_8gpu.log 2>&1
tail -1 gpurun_out/r2e_full_8gpu.log | cut -c1-200
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --workload distance --clips 16384 --steps 2 --warmup 3 --no-profile-pass > gpurun_out/r2e_dist_8gpu.log 2>&1
tail -1 gpurun_out/r2e_dist_8gpu.log | cut -c1-200
timeout 200 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-profile-pass > gpurun_out/r2e_stage_8gpu.log 2>&1
tail -1 gpurun_out/r2e_stage_8gpu.log | cut -c1-200
