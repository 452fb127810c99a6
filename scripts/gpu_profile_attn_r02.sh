# Round 2: ncu --set full of the decoder cross-attention kernels (csrc/attention.cu) inside one full training step
# (64 clips per micro-batch, full 20-vertex complex: 250 queries x 6,175 memory rows x 4 heads per clip),
# after the plain run of the same command exited 0.
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload full_step --steps 1 --warmup 3 --no-profile-pass"
$CMD > gpurun_out/r02_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_" -s 3 -c 3 -f -o gpurun_out/r02_prof_attn $CMD > gpurun_out/r02_attn_ncu.log 2>&1
echo "exit attn: $?"
