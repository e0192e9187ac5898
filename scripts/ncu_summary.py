"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) and optionally the SASS source page: per kernel the
duration, issue/pipe utilisation, DRAM bytes, occupancy limits and the top stall reasons / top stalled instructions."""
import csv, sys

WANT = ['gpu__time_duration.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("KERNEL", d['Kernel Name'][:70])
        for k in WANT:
            if k in d:
                print("   %-75s %s %s" % (k, d[k], rows[1][hdr.index(k)]))
        st = [(float(d[k]), k) for k in hdr if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and d[k] not in ('', 'n/a')]
        for v, k in sorted(st, reverse=True)[:7]:
            print("     stall %-22s %.2f" % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


def source(path, top=30):
    rows = list(csv.reader(open(path)))
    kern, hdr, data = None, None, []
    def flush():
        if not data:
            return
        idx = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[2]) for r in data)
        print("KERNEL", kern[:70], "samples", tot, "instructions", len(data))
        best = sorted(range(len(data)), key=lambda i: -int(data[i][2]))[:top]
        for i in sorted(best):
            r = data[i]
            st = {h.replace('stall_', ''): int(r[idx[h]]) for h in hdr if h.startswith('stall_') and '(' not in h and r[idx[h]] not in ('', '0')}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print("  %5d %-58s %6s %s" % (i, r[1].strip()[:58], r[2], st))
    for r in rows:
        if r and r[0] == 'Kernel Name':
            flush(); kern, hdr, data = r[1], None, []
        elif r and r[0] == 'Address':
            hdr = r
        elif hdr and len(r) >= len(hdr) - 2:
            data.append(r)
    flush()


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
