"""Summarise an .ncu-rep (read offline with `ncu -i`) into a small text table + traffic JSON for profiles/."""
import csv, io, json, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']


def main(rep, out_txt, out_json=None):
    if rep.endswith('.csv'):          # already exported on the GPU box (tools/ncu_capture.sh)
        raw = open(rep).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    launches = []
    with open(out_txt, 'w') as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (cold-cache, serialised replays)\n")
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            f.write(f"\n== {name}\n")
            vals = {}
            for k in KEYS:
                if k in hdr:
                    vals[k] = r[hdr.index(k)]
                    f.write(f"  {k:86s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}\n")
            try:
                def tobytes(k):
                    v, u = float(vals[k].replace(',', '')), units[hdr.index(k)]
                    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
                short = name.split('<')[0].split('(')[0].replace('void ', '').replace('scc::', '')
                traffic.setdefault(short, []).append(tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'))
                launches.append([name, tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'),
                                 vals.get('gpu__time_duration.sum')])
            except Exception:
                pass
    if out_json:
        with open(out_json, 'w') as f:
            out = {k: sum(v) / len(v) for k, v in traffic.items()}
            out["per_launch"] = launches          # in launch order: [kernel, dram bytes, duration us]
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main(*sys.argv[1:])
