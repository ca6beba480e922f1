"""Opcode histogram of the shipped library per kernel (offline: cuobjdump -sass), restricted to the mnemonics that prove the
Blackwell-native paths: tcgen05 (UTCHMMA / LDTM / STTM / UTC*), TMA (UTMA*, UBLKCP = 1-D bulk copy, UBLKPF = bulk L2 prefetch),
mbarrier (SYNCS.*), cluster barriers (UCGABAR_*), packed fp32 (FFMA2 / FMUL2 / FADD2), 3-input min/max (FMNMX3), 256-bit global
accesses, MUFU.  usage: python tools/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'voicebox-pytorch_b200', 'lib', 'libvbx_sm100a.so')
KEEP = re.compile(r'^(UTC|LDTM|STTM|UTMA|UBLK|SYNCS|UCGABAR|ELECT|FFMA2|FMUL2|FADD2|FMNMX3|MUFU|HMMA|BAR\.|RED|ATOM|REDG|LDG\.E\.(ENL2\.)?256|STG\.E\.(ENL2\.)?256)')
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', txt)), capture_output=True, text=True).stdout.split('\n')
kern = OrderedDict()
cur = None
it = iter(names)
for ln in txt.split('\n'):
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = next(it)
        kern[cur] = Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)', ln)
    if m and cur is not None:
        op = m.group(1)
        if KEEP.match(op):
            op = re.sub(r'\.(SYS|STRONG|GPU|CONSTANT|WEAK|EF|EL|NODEC)\b', '', op)
            kern[cur][op] += 1
rev = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True, cwd=ROOT).stdout.strip()
print(f'# SASS opcode histogram of the shipped library ({os.path.getsize(lib)} bytes, built from csrc/ at {rev}): cuobjdump -sass, mnemonics that prove the Blackwell-native paths\n')
for k, c in kern.items():
    if not c:
        continue
    print('## ' + k[:150])
    for op, n in c.most_common():
        print(f'{n:8d}  {op}')
