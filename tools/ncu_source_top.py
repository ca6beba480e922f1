"""Reads `ncu -i <rep> --page source --csv --kernel-name <k>` output and prints where the warps of a kernel spend their
time: aggregate stall reasons, the SASS instructions with the most stall samples, and (with --seq lo hi) the instruction
stream with cumulative samples so that phases (waits, passes, epilogue) can be delimited.  CPU only.

    ncu -i prof.ncu-rep --page source --csv --kernel-name attn_fwd_kernel > fwd_src.csv
    python tools/ncu_source_top.py fwd_src.csv [--launch 0] [--top 40] [--seq 250 1100]
"""
import argparse
import csv

KEY_OPS = ('LDTM', 'STTM', 'BAR', 'SYNCS', 'UTC', 'STG', 'LDG', 'EXIT', 'LDS', 'STS', 'BRA', 'UBLK', 'UTMA')


def load(path):
    blocks, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'data': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None and len(r) > 10:
            cur['data'].append(r)
    return blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--launch', type=int, default=0)
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--seq', type=int, nargs=2, default=None)
    a = ap.parse_args()
    blocks = load(a.csv)
    b = blocks[a.launch]
    hdr, data = b['hdr'], b['data']
    col = {h: i for i, h in enumerate(hdr)}
    smp = lambda r: int(r[col['# Samples']])  # noqa: E731
    tot = sum(smp(r) for r in data)
    inst = sum(int(r[col['Instructions Executed']]) for r in data)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print(f'{len(blocks)} captured launches; launch {a.launch}: {b["name"][:60]}  samples {tot}  warp-instructions {inst}  SASS lines {len(data)}')
    agg = sorted(((sum(int(r[col[h]]) for r in data), h[6:]) for h in stalls), reverse=True)
    print('stall reasons:', ', '.join(f'{n} {100 * v / tot:.1f}%' for v, n in agg[:9]))
    if a.seq is None:
        for idx, r in sorted(enumerate(data), key=lambda x: -smp(x[1]))[:a.top]:
            st = sorted(((int(r[col[h]]), h[6:]) for h in stalls), reverse=True)[:2]
            print(f'{idx:5d} {smp(r):6d} {100 * smp(r) / tot:5.2f}%  ex={int(r[col["Instructions Executed"]]):9d}  '
                  f'{r[col["Source"]].strip()[:60]:60s} {st}')
    else:
        lo, hi = a.seq
        acc = sum(smp(r) for r in data[:lo])
        for idx, r in enumerate(data[lo:hi], lo):
            acc += smp(r)
            src = r[col['Source']].strip()
            if smp(r) * 200 >= tot or any(k in src for k in KEY_OPS):
                print(f'{idx:5d} {smp(r):6d}  cum {100 * acc / tot:5.1f}%  ex={int(r[col["Instructions Executed"]]):9d}  {src[:80]}')


if __name__ == '__main__':
    main()
