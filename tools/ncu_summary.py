"""Summarise `ncu --set full` reports (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into a markdown table, one row per
captured launch, and (with --traffic) refresh profiles/traffic.json with dram read+write bytes per launch of each C-ABI entry.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] [--title T] [--out file.md] [--traffic]"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'), ('launch__registers_per_thread', 'regs'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long-sb / issue'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts')]
ENTRY = {'adarms_fwd_kernel': 'vbx_adarms_fwd', 'adarms_bwd_kernel': 'vbx_adarms_bwd', 'geglu_fwd_kernel': 'vbx_geglu_fwd',
         'geglu_bwd_kernel': 'vbx_geglu_bwd', 'convpos_fwd_kernel': 'vbx_convpos_fwd', 'convpos_bwd_kernel': 'vbx_convpos_bwd',
         'qkrope_fwd_kernel': 'vbx_qkrope_fwd', 'qkrope_bwd_kernel': 'vbx_qkrope_bwd', 'attn_fwd_kernel': 'vbx_attn_fwd',
         'attn_fwd2_kernel': 'vbx_attn_fwd', 'attn_bwd_kernel': 'vbx_attn_bwd', 'gemm_geglu_bwd_kernel': 'vbx_ff2_dgrad_geglu_bwd',
         'gemm_bf16_kernel<1': 'vbx_ff1_geglu', 'gemm_bf16_kernel<2': 'vbx_ff1_geglu', 'gemm_bf16_kernel<0': 'vbx_gemm_bf16'}
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def load(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'^vbx::', '', name)
    m = re.match(r'([A-Za-z0-9_:]+(<[^>(]*>)?)', name)
    return m.group(1) if m else name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('reports', nargs='+')
    ap.add_argument('--title', default=None)
    ap.add_argument('--out', default=None)
    ap.add_argument('--traffic', action='store_true')
    args = ap.parse_args()
    lines, traffic = [], {}
    for rep in args.reports:
        hdr, units, data = load(rep)
        lines += [f'## {args.title or os.path.basename(rep)}', '', '| kernel | ' + ' | '.join(c[1] for c in COLS) + ' |', '|---|' + '---|' * len(COLS)]
        iname = hdr.index('Kernel Name')
        for d in data:
            cells = []
            for met, _ in COLS:
                if met in hdr:
                    i = hdr.index(met)
                    v = d[i]
                    try:
                        v = f'{float(v.replace(",", "")):.4g}'
                    except ValueError:
                        pass
                    cells.append(f'{v} {units[i]}'.strip())
                else:
                    cells.append('-')
            k = short(d[iname])
            lines.append(f'| `{k}` | ' + ' | '.join(cells) + ' |')
            if 'dram__bytes_read.sum' in hdr:
                ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
                tot = float(d[ir].replace(',', '')) * SCALE.get(units[ir], 1.0) + float(d[iw].replace(',', '')) * SCALE.get(units[iw], 1.0)
                for key, entry in ENTRY.items():
                    if k.startswith(key):
                        traffic.setdefault(entry, []).append(tot)
                        break
        lines.append('')
    text = '\n'.join(lines)
    print(text)
    if args.out:
        with open(args.out, 'a') as f:
            f.write(text + '\n')
    if args.traffic:
        p = os.path.join(ROOT, 'profiles', 'traffic.json')
        cur = json.load(open(p)) if os.path.exists(p) else {}
        for e, v in traffic.items():
            cur[e] = sorted(v)[len(v) // 2]      # median launch
        json.dump(cur, open(p, 'w'), indent=1)
        print('updated', p, {e: round(cur[e] / 1e6, 1) for e in traffic})


if __name__ == '__main__':
    main()
