"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iname, ival, iunit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    imet = hdr.index('Metric Name')
    for r in rd:
        if len(r) <= ival or r[imet] != 'gpu__time_duration.sum':
            continue
        v = float(r[ival].replace(',', ''))
        u = r[iunit]
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)  # -> us
        rows.append((r[iname], v))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        n = re.sub(r'<.*', '', n)          # drop template args
        n = re.sub(r'\(.*', '', n).strip()
        agg[n][0] += 1
        agg[n][1] += v

    def klass(n):
        if n.startswith('vbx::') or 'vbx' in n:
            return 'vbx (this repo)'
        if re.search(r'gemm|cutlass|nvjet|cublas|sm\d+_xmma|s\d+gemm', n, re.I):
            return 'library GEMM (cuBLASLt)'
        if re.search(r'nccl', n, re.I):
            return 'NCCL'
        return 'torch glue (casts, adds, optimizer, rng)'
    by_class = defaultdict(float)
    for n, (c, v) in agg.items():
        by_class[klass(n)] += v
    lines_out = [f'launches: {len(rows)}   serialized device time: {tot / 1e3:.2f} ms (cold-cache, per-launch; compare SHARES)', '',
                 '| class | ms | share |', '|---|---|---|']
    for k, v in sorted(by_class.items(), key=lambda kv: -kv[1]):
        lines_out.append(f'| {k} | {v / 1e3:.2f} | {100 * v / tot:.1f} % |')
    lines_out += ['', '| kernel | launches | total ms | avg us | share |', '|---|---|---|---|---|']
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines_out.append(f'| `{n[:90]}` | {c} | {v / 1e3:.3f} | {v / c:.1f} | {100 * v / tot:.1f} % |')
    text = '\n'.join(lines_out)
    print(text)
    if out:
        open(out, 'w').write(text + '\n')


if __name__ == '__main__':
    main(*sys.argv[1:3])
