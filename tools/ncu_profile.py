#!/usr/bin/env python
"""ncu capture + summary for the numbers bench.py quotes (no literals in the roofline).

  capture (on the GPU box, ONE GPU, never a multi-rank command; ncu replays each kernel ~40x):
      python tools/ncu_profile.py capture --out gpurun_out/ncu_r02 [--kernels REGEX] [--skip N] [--count N] -- <cmd ...>
    default <cmd> = python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline
  summarize (anywhere ncu is installed; reads the .ncu-rep):
      python tools/ncu_profile.py summarize gpurun_out/ncu_r02/prof.ncu-rep --out profiles/r02

``summarize`` writes ``ncu_metrics.json`` (per kernel family: duration, DRAM bytes read+written per launch, FMA / XU
(MUFU) / ALU pipe utilisation, issue-slot utilisation, registers, resident warps, stall-reason shares) -- the file
``bench.py`` reads for ``roofline.traffic`` and the pipe fractions -- and a human-readable ``ncu_summary.txt``.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

FAMILIES = [  # (key, regex on the demangled kernel name)
    ('render_sg_fwd', r'render_sg_group_kernel<0,'),
    ('render_sg_bwd', r'render_sg_group_kernel<1,'),
    ('sg_env_fwd', r'sg_env_group_kernel<0,|sg_env_tile_fwd_kernel'),
    ('sg_env_bwd', r'sg_env_group_kernel<1,'),
    ('sg_reconst', r'sg_env_group_kernel<[2-5],'),
    ('render_env_fwd', r'render_env_fwd|render_env_tma_kernel<0'),
    ('render_env_bwd', r'render_env_bwd|render_env_tma_kernel<1'),
    ('bg_q', r'bg_q_kernel'), ('bg_p', r'bg_p_kernel'), ('bg_xr', r'bg_xr_kernel'), ('bg_pcg', r'bg_pcg'),
    ('render_loss', r'render_loss_kernel'), ('lsreg_stats1', r'lsreg_stats1'), ('lsreg_stats2', r'lsreg_stats2'),
]
METRICS = {
    'duration_us': ('gpu__time_duration.sum', None),
    'dram_read_bytes': ('dram__bytes_read.sum', None),
    'dram_write_bytes': ('dram__bytes_write.sum', None),
    'registers': ('launch__registers_per_thread', None),
    'fma_pipe_pct': ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', None),
    'xu_pipe_pct': ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', None),
    'alu_pipe_pct': ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', None),
    'fma_inst_pct': ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', None),
    'issue_active_pct': ('smsp__issue_active.avg.pct_of_peak_sustained_active', None),
    'warps_active_pct': ('sm__warps_active.avg.pct_of_peak_sustained_active', None),
    'inst_executed': ('smsp__inst_executed.sum', None),
    'dram_throughput_pct': ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', None),
    'l2_hit_pct': ('lts__t_sector_hit_rate.pct', None),
    'smem_wavefront_pct': ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', None),
}
UNIT_SCALE = {'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6, 'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0}


def raw_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    return rows[start], rows[start + 1], rows[start + 2:]


def summarize(rep, outdir):
    hdr, units, rows = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for h, i in col.items() if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
    res, text = {}, []
    for r in rows:
        name = r[col['Kernel Name']]
        fam = next((k for k, rx in FAMILIES if re.search(rx, name)), None)
        if fam is None:
            fam = re.sub(r'[^A-Za-z0-9_]', '_', name.split('(')[0].replace('void ', ''))[:48]
        m = {'kernel': name[:120], 'grid': r[col['Grid Size']], 'block': r[col['Block Size']]}
        for key, (metric, _) in METRICS.items():
            if metric in col and r[col[metric]] != '':
                v = float(r[col[metric]].replace(',', ''))
                u = units[col[metric]]
                if key.endswith('_us') or key.endswith('_bytes'):
                    v *= UNIT_SCALE.get(u, 1.0)
                m[key] = v
        if 'dram_read_bytes' in m and 'dram_write_bytes' in m:
            m['dram_bytes'] = m['dram_read_bytes'] + m['dram_write_bytes']
            if m.get('duration_us'):
                m['dram_GBs'] = m['dram_bytes'] / m['duration_us'] / 1e3
        if 'warps_active_pct' in m:
            m['warps_per_sm'] = m['warps_active_pct'] / 100.0 * 64
        st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(r[i] or 0) for h, i in stall_cols}
        tot = sum(st.values())
        if tot > 0:
            m['stalls_pct'] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
        cnt = res.get(fam, {}).get('launches_captured', 0)
        if fam not in res or m.get('duration_us', 0) > res[fam].get('duration_us', 0):   # keep the longest instance
            res[fam] = m
        res[fam]['launches_captured'] = cnt + 1
        text.append('%s\n  %s' % (name[:110], '  '.join('%s=%s' % (k, ('%.4g' % v) if isinstance(v, float) else v)
                                                          for k, v in m.items() if k not in ('kernel', 'stalls_pct'))))
        if 'stalls_pct' in m:
            text.append('  stalls: ' + '  '.join('%s %.1f%%' % kv for kv in m['stalls_pct'].items()))
    os.makedirs(outdir, exist_ok=True)
    res['_source'] = {'report': os.path.basename(rep), 'how': 'ncu --set full --clock-control none --import-source on; '
                      'python tools/ncu_profile.py summarize'}
    with open(os.path.join(outdir, 'ncu_metrics.json'), 'w') as f:
        json.dump(res, f, indent=1, sort_keys=True)
    with open(os.path.join(outdir, 'ncu_summary.txt'), 'w') as f:
        f.write('\n'.join(text) + '\n')
    return res


def capture(out, kernels, skip, count, cmd):
    os.makedirs(out, exist_ok=True)
    rep = os.path.join(out, 'prof')
    full = ['ncu', '--set', 'full', '--clock-control', 'none', '--import-source', 'on', '-k', 'regex:' + kernels,
            '-s', str(skip), '-c', str(count), '-f', '-o', rep] + cmd
    print(' '.join(full), flush=True)
    r = subprocess.run(full, stdout=open(os.path.join(out, 'ncu_stdout.log'), 'w'), stderr=subprocess.STDOUT)
    return r.returncode


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='mode', required=True)
    c = sub.add_parser('capture')
    c.add_argument('--out', default='gpurun_out/ncu_r02')
    c.add_argument('--kernels', default='render_sg_group_kernel')
    c.add_argument('--skip', type=int, default=6)
    c.add_argument('--count', type=int, default=2)
    c.add_argument('cmd', nargs=argparse.REMAINDER)
    s = sub.add_parser('summarize')
    s.add_argument('rep')
    s.add_argument('--out', default='profiles/r02')
    a = ap.parse_args()
    if a.mode == 'capture':
        cmd = [x for x in a.cmd if x != '--'] or [sys.executable, 'bench.py', '--steps', '2', '--warmup', '3', '--no-extras',
                                                  '--no-cpu-baseline']
        sys.exit(capture(a.out, a.kernels, a.skip, a.count, cmd))
    r = summarize(a.rep, a.out)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != 'stalls_pct'} for k, v in r.items()}, indent=1)[:3000])


if __name__ == '__main__':
    main()
