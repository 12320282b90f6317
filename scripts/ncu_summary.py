"""Summarise an .ncu-rep (ncu --set full) per kernel launch: duration, DRAM traffic, pipe utilisation, stall reasons.
usage: python scripts/ncu_summary.py file.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_active.avg', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in data:
    print('=' * 100)
    print(r[col['Kernel Name']][:110], ' id', r[col['ID']])
    for w in want:
        if w in col: print(f'  {w:75s} {r[col[w]]:>16s} {units[col[w]]}')
    st = sorted(((float(r[col[h]] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in stalls), reverse=True)[:6]
    print('  top stall reasons (warps per issue-active cycle):', ', '.join(f'{n} {v:.2f}' for v, n in st))
