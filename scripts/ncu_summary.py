"""Summarise .ncu-rep captures into one JSON (read here, no GPU):  python scripts/ncu_summary.py out.json rep1 rep2 ..."""
import csv, io, json, subprocess, sys

KEYS = {'gpu__time_duration.sum': 'duration_us', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'launch__registers_per_thread': 'regs', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
        'smsp__inst_executed.sum': 'warp_inst', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_fma_pct',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_alu_pct',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'pipe_xu_pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_tensor_pct',
        'sm__inst_executed_pipe_tensor.sum': 'tensor_inst',
        'launch__occupancy_limit_registers': 'occ_limit_regs', 'launch__occupancy_limit_shared_mem': 'occ_limit_smem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts'}


def summarise(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    res = []
    for v in rows[2:]:
        d = dict(zip(h, v))
        e = {'kernel': d.get('Kernel Name', '')[:120]}
        for k, name in KEYS.items():
            if k in d and d[k] not in ('', 'n/a'):
                try:
                    e[name] = float(d[k].replace(',', ''))
                except ValueError:
                    e[name] = d[k]
        stalls = []
        for k, x in d.items():
            if 'average_warp' in k and 'issue_stalled' in k and 'not_issued' not in k and k.endswith('.ratio'):
                try:
                    stalls.append((round(float(x), 2), k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        e['top_stalls_per_issue'] = sorted(stalls, reverse=True)[:5]
        units = {k: rows[1][i] for i, k in enumerate(h)}
        e['units'] = {KEYS[k]: units.get(k, '') for k in KEYS if k in d}
        res.append(e)
    return res


if __name__ == '__main__':
    out = {rep.split('/')[-1]: summarise(rep) for rep in sys.argv[2:]}
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
    for k, v in out.items():
        for e in v:
            print(k, {x: e[x] for x in e if x not in ('units',)})
