#!/usr/bin/env python3
"""Turn an .ncu-rep (one kernel launch, `ncu --set full`) into the short text summary kept under
profiles/.  Usage: tools/ncu_summary.py report.ncu-rep "<command that produced it>" [blocks] > out.txt"""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg',
        'device__attribute_clock_rate',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def page(rep, name):
    return list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True,
                                          text=True).stdout.splitlines()))


def main():
    rep, cmd = sys.argv[1], sys.argv[2]
    blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = page(rep, 'raw')
    d = dict(zip(rows[0], zip(rows[1], rows[2])))
    print(f"# {cmd}\n# extracted from {rep.split('/')[-1]} with `ncu -i ... --page raw --csv` / `--page source --csv`\n")
    for k in KEEP:
        if k in d:
            print(f"{k:90s} {d[k][0]:16s} {d[k][1]}")
    src = page(rep, 'source')
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in src[2:]:
        if len(r) < len(hdr):
            continue
        s = r[ix['Source']].strip().split()
        if not s:
            continue
        op = s[1] if s[0].startswith('@') else s[0]
        if op.startswith(('LDS', 'LDG', 'STG', 'STS')):
            a = agg.setdefault(op, [0, 0, 0])
            a[0] += int(r[ix['Instructions Executed']] or 0)
            a[1] += int(r[ix['L1 Wavefronts Shared']] or 0)
            a[2] += int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
    print("\n# per-opcode totals from the source page: warp-instructions, shared wavefronts, ideal shared wavefronts")
    for op, (n, w, i) in sorted(agg.items()):
        print(f"{op:24s} {n:14d} {w:14d} {i:14d}")
    # warp-state sampling split by role: the bitsliced warps' code lies between the two USETMAXREG
    # instructions (setmaxnreg.inc ... setmaxnreg.dec), the table-driven warps' code after the second
    marks = [i for i, r in enumerate(src[2:]) if len(r) >= len(hdr) and 'USETMAXREG' in r[ix['Source']]]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in src[2:] if len(r) >= len(hdr)]
    regions = None
    if len(marks) >= 2:
        regions = (('bitsliced co-runner warps', body[marks[0]:marks[-1]]), ('table-driven warps', body[marks[-1]:]))
    else:
        # kernels without setmaxnreg (ctr_queue8_kernel): the co-runner role lies between the table fill's barrier and
        # the first unpredicated EXIT after it, the table-driven role behind that
        bar = next((i for i, r in enumerate(body) if r[ix['Source']].strip().startswith('BAR.SYNC')), None)
        if bar is not None:
            ex = next((i for i in range(bar, len(body)) if body[i][ix['Source']].strip().startswith('EXIT')), None)
            if ex is not None:
                regions = (('bitsliced co-runner warps', body[bar:ex + 1]), ('table-driven warps', body[ex + 1:]))
    if regions and stalls:
        for name, rows_ in regions:
            tot = {h: sum(int(r[ix[h]] or 0) for r in rows_) for h in stalls}
            n = sum(tot.values()) or 1
            inst = sum(int(r[ix['Instructions Executed']] or 0) for r in rows_)
            top = sorted(tot.items(), key=lambda kv: -kv[1])[:9]
            print(f"\n# {name}: static {len(rows_)} instructions, executed {inst} warp-instructions, {n} samples; "
                  + " ".join(f"{k[6:]}={v / n:.2f}" for k, v in top))
    if blocks:
        rows32 = blocks / 32
        cyc = float(d['sm__cycles_active.avg'][1])
        sh = float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'][1])
        print(f"\n# derived, per warp-row of 32 blocks ({blocks} blocks): SM cycles = {cyc * 148 / rows32:.1f}, "
              f"shared-pipe wavefronts = {sh / rows32:.1f}")


if __name__ == '__main__':
    main()
