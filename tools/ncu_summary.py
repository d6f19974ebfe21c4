"""Markdown summary of an ncu --set full report (one kernel launch): the counters the roofline discussion uses.
    ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv [tile_bytes]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def get(name, default='n/a'):
    return m.get(name, (default, ''))[0]


def f(name):
    try:
        return float(get(name, 'nan').replace(',', ''))
    except ValueError:
        return float('nan')


dur_us = f('gpu__time_duration.sum')
unit = m.get('gpu__time_duration.sum', ('', 'us'))[1]
if unit.startswith('ms'):
    dur_us *= 1e3
elif unit.startswith('ns'):
    dur_us /= 1e3
rd, wr = f('dram__bytes_read.sum'), f('dram__bytes_write.sum')
scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
rd *= scale.get(m['dram__bytes_read.sum'][1], 1)
wr *= scale.get(m['dram__bytes_write.sum'][1], 1)
print('| counter | value |\n|---|---|')
print('| kernel | `%s` |' % get('Kernel Name'))
print('| grid x block | %s x %s, %s registers/thread, %s KB dynamic shared memory |' % (get('Grid Size'), get('Block Size'), get('launch__registers_per_thread'), get('launch__shared_mem_per_block_dynamic')))
print('| duration (under ncu, clocks not locked) | %.1f us at SM %.2f GHz |' % (dur_us, f('sm__cycles_elapsed.avg.per_second')))
print('| DRAM read / written | %.1f MB / %.1f MB -> %.0f GB/s, %s %% of peak |' % (rd / 1e6, wr / 1e6, (rd + wr) / dur_us / 1e3, get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')))
print('| tensor pipe active | %s %% |' % get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'))
print('| issue slots busy | %s %% (%.0f warp instructions) |' % (get('smsp__issue_active.avg.pct_of_peak_sustained_active'), f('smsp__inst_executed.sum')))
lsu, tc = f('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'), f('l1tex__data_pipe_tc_wavefronts_mem_shared.sum')
cyc = f('sm__cycles_elapsed.avg')
print('| shared-memory wavefronts: threads (ld / st) | %.0f (%.0f / %.0f), bank conflicts %.0f |' % (lsu, f('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum'), f('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum'), f('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')))
print('| shared-memory wavefronts: tensor-core operand reads | %.0f |' % tc)
print('| shared-memory pipe, threads + tensor core | %.1f %% of one wavefront per cycle per SM |' % (100.0 * (lsu + tc) / (cyc * 148)))
print('| warps resident per SM | %s |' % get('sm__warps_active.avg.per_cycle_active'))
if len(sys.argv) > 2:
    tile = float(sys.argv[2])
    tiles = rd / tile
    print('| per %.0f-byte tile | %.0f warp instructions, %.0f + %.0f shared-memory wavefronts, %.0f SM cycles |' % (
        tile, f('smsp__inst_executed.sum') / tiles, lsu / tiles, tc / tiles, cyc * 148 / tiles))
