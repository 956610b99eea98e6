"""Turn the ncu reports brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py r1      # reads gpurun_out/prof_r1_*.ncu-rep, gpurun_out/launches_r1.csv

Writes profiles/<round>_ncu_summary.md (per-kernel counters + top source lines by stall samples),
profiles/<round>_launch_shares.md (share of device time per kernel from the launch list) and profiles/ncu_traffic.json
(dram bytes per launch, read by bench.py for `roofline.traffic`).  Runs here (no GPU needed).
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [('gpu__time_duration.sum', 'duration'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs/thread'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe active %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('smsp__inst_executed.sum', 'warp instructions'),
        ('sm__icc_request_hit_rate.pct', 'instruction cache hit %'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %')]
UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source_top(rep, kernel, top=12):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k',
                          'regex:' + kernel, '-c', '1'], capture_output=True, text=True).stdout
    cur = hdr = line = None
    agg, src = collections.defaultdict(lambda: [0, 0]), {}
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) >= 2 and r[0] == 'Line No':
            hdr = r
            iI, iS = hdr.index('Instructions Executed'), hdr.index('# Samples')
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != '':
            line = r[0]
            src[(cur, line)] = r[1]
        try:
            agg[(cur, line)][0] += int(r[iI])
            agg[(cur, line)][1] += int(r[iS])
        except ValueError:
            pass
    ti, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
    return [(f, l, 100.0 * v[0] / ti, 100.0 * v[1] / ts, src.get((f, l), '').strip()[:90]) for (f, l), v in rows]


def main(rnd):
    counters = {}
    traffic, md = {}, ['# ncu summary, round %s' % rnd, '',
                       'Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200, '
                       'after the same command had exited 0 without ncu.  Durations under ncu are serialised and cold-cache: '
                       'use them for shares and counters, not as bench numbers.', '']
    for rep in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', 'prof_%s_*.ncu-rep' % rnd))):
        hdr, units, rows = raw(rep)
        seen = set()
        for r in rows:
            name = r[hdr.index('Kernel Name')].split('(')[0]
            if name in seen:
                continue
            seen.add(name)
            if '_c4' in os.path.basename(rep):          # the same kernel captured on the config-4 workload
                name += '@config4'
            md += ['## %s  (`%s`)' % (name, os.path.basename(rep)), '', '| counter | value |', '|---|---|']
            byt = 0.0
            for k, label in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    md.append('| %s | %s %s |' % (label, r[i], units[i]))
                    if k.startswith('dram__bytes'):
                        byt += float(r[i]) * UNIT.get(units[i], 1.0)
            traffic[name] = byt
            counters[name] = {label: r[hdr.index(k)] + ' ' + units[hdr.index(k)] for k, label in KEYS if k in hdr}
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
                    try:
                        stalls.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
                    except ValueError:
                        pass
            tot = sum(s for s, _ in stalls) or 1.0
            md += ['', 'Warp stall samples: ' + ', '.join('%s %.0f %%' % (n, 100 * s / tot) for s, n in sorted(stalls, reverse=True)[:6]), '',
                   'Top source lines by stall samples (file:line, % instructions, % samples):', '']
            for f, l, pi, ps, text in source_top(rep, name):
                md.append('* `%s:%s` %.1f %% inst, %.1f %% samples -- `%s`' % (f, l, pi, ps, text))
            md.append('')
    with open(os.path.join(ROOT, 'profiles', '%s_ncu_summary.md' % rnd), 'w') as fh:
        fh.write('\n'.join(md) + '\n')
    with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'), 'w') as fh:
        json.dump(traffic, fh, indent=1, sort_keys=True)
    with open(os.path.join(ROOT, 'profiles', 'ncu_counters.json'), 'w') as fh:
        json.dump(counters, fh, indent=1, sort_keys=True)
    # launch shares
    lst = os.path.join(ROOT, 'gpurun_out', 'launches_%s.csv' % rnd)
    if os.path.exists(lst):
        per = collections.defaultdict(lambda: [0, 0.0])
        rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
        hdr = rows[0]
        iN, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
        scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6, 'second': 1e3}
        for r in rows[1:]:
            n = r[iN].split('(')[0]
            per[n][0] += 1
            per[n][1] += float(r[iV].replace(',', '')) * scale.get(r[iU], 1e-6)
        tot = sum(v[1] for v in per.values()) or 1.0
        out = ['# launch list shares, round %s (`ncu --metrics gpu__time_duration.sum`, first %d launches of the bench command)' % (rnd, len(rows) - 1),
               '', '| kernel | launches | total ms | share |', '|---|---|---|---|']
        for n, (c, ms) in sorted(per.items(), key=lambda kv: -kv[1][1])[:25]:
            out.append('| %s | %d | %.3f | %.1f %% |' % (n[:80], c, ms, 100 * ms / tot))
        with open(os.path.join(ROOT, 'profiles', '%s_launch_shares.md' % rnd), 'w') as fh:
            fh.write('\n'.join(out) + '\n')
    print('wrote profiles/%s_ncu_summary.md, ncu_traffic.json' % rnd)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r1')
