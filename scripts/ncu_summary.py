"""Summarise an .ncu-rep: per-launch headline metrics and the hottest SASS lines (stall samples)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__pcsamp_sample_count', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_warps_issue_stalled_sleeping',
        'smsp__pcsamp_warps_issue_stalled_tex_throttle', 'smsp__pcsamp_warps_issue_stalled_drain',
        'smsp__pcsamp_warps_issue_stalled_imc_miss', 'smsp__pcsamp_warps_issue_stalled_misc']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        d = {'name': r[hdr.index('Kernel Name')][:70]}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)]
        res.append(d)
    return res


def sass(rep, which, top):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    s, e = starts[which], starts[which + 1]
    hdr = rows[s + 1]
    si, src = hdr.index('# Samples'), hdr.index('Source')
    ex = hdr.index('Instructions Executed')
    body = rows[s + 2:e]
    tot = sum(int(r[si]) for r in body)
    print(f'-- SASS of launch {which}: {len(body)} instructions, {tot} samples, {sum(int(r[ex]) for r in body)} warp-instr executed')
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:top]
    for i in sorted(idx):
        print(f'{i:5d} {int(body[i][si]):6d} {100.0 * int(body[i][si]) / max(tot, 1):5.1f}%  exec={body[i][ex]:>8s}  {body[i][src][:100]}')


if __name__ == '__main__':
    rep = sys.argv[1]
    launches = raw(rep)
    for i, d in enumerate(launches):
        print(f'== launch {i}: {d["name"]}')
        for k in KEYS:
            if k in d:
                print(f'   {k} = {d[k]}')
    if len(sys.argv) > 2:
        sass(rep, int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
