#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into one line per kernel launch: time, DRAM traffic, occupancy, issue utilisation,
warp execution efficiency, top stall reasons.  usage: tools/ncu_summary.py X.ncu-rep [> profiles/X_summary.txt]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=float('nan')):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return d
def unit(k): return units[ix[k]] if k in ix else ''
def scale(k, v):
    u = unit(k)
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1}.get(u, 1)
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    name = r[ix['Kernel Name']].split('(')[0]
    t = scale('gpu__time_duration.sum', g(r, 'gpu__time_duration.sum'))
    rd = scale('dram__bytes_read.sum', g(r, 'dram__bytes_read.sum')); wr = scale('dram__bytes_write.sum', g(r, 'dram__bytes_write.sum'))
    st = sorted(((g(r, h, 0.0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for h in stalls), reverse=True)[:4]
    print("%-28s grid %-14s blk %-12s %8.1f us | dram rd %7.1f MB wr %7.1f MB = %6.0f GB/s (%4.1f%% of peak) | L2 %6.1f MB | regs %3d occ %4.1f%% | issue %4.1f%% | thr/inst %4.1f | l1 hit %4.1f%% | stalls %s"
          % (name[-28:], r[ix['Grid Size']], r[ix['Block Size']], t * 1e6, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e9,
             g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), scale('lts__t_bytes.sum', g(r, 'lts__t_bytes.sum')) / 1e6,
             g(r, 'launch__registers_per_thread'), g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
             g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), g(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'),
             g(r, 'l1tex__t_sector_hit_rate.pct'), ", ".join("%s %.1f" % (n, v) for v, n in st)))
