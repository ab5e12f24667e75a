#!/usr/bin/env python
"""Instruction mix of one kernel's hottest loop, from `cuobjdump -sass` (no GPU needed).

    python tools/sass_mix.py <mangled-name-substring> [--all]

Finds the function, takes the body of its longest backward branch (the main loop; --all = whole function) and
counts opcodes by class. Used to budget instructions per cell update before spending GPU time.
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   'lattice_boltzmann_parallel_solver_b200', 'liblbm_b200.so')


def main():
    want = sys.argv[1]
    whole = '--all' in sys.argv
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    funcs = re.split(r'\n\s*Function : ', txt)[1:]
    for fn in funcs:
        name = fn.split('\n', 1)[0].strip()
        if want not in name:
            continue
        ins = re.findall(r'/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\w+\s+)?)([A-Z0-9_.]+)([^;]*);', fn)
        addr = [int(a, 16) for a, _, _, _ in ins]
        lo, hi = 0, addr[-1]
        if not whole:
            best = 0
            for a, _, op, rest in ins:
                if op.startswith('BRA'):
                    m = re.search(r'0x([0-9a-f]+)', rest)
                    if m and int(m.group(1), 16) < int(a, 16) and int(a, 16) - int(m.group(1), 16) > best:
                        best = int(a, 16) - int(m.group(1), 16)
                        lo, hi = int(m.group(1), 16), int(a, 16)
        ops = [op for (a, _, op, _) in ins if lo <= int(a, 16) <= hi]
        c = collections.Counter(ops)
        cls = collections.Counter()
        for op, k in c.items():
            b = op.split('.')[0]
            if b in ('DADD', 'DMUL', 'DFMA', 'MUFU'):
                cls['fp64 pipe'] += k
            elif b in ('LDG', 'STG', 'LDS', 'STS', 'LDC', 'LDCU', 'LD', 'ST', 'LDL', 'STL'):
                cls['memory' + (' (LOCAL!)' if b in ('LDL', 'STL') else '')] += k
            elif b in ('MOV', 'UMOV', 'CS2R') or op.startswith('IMAD.MOV'):
                cls['moves'] += k
            elif b in ('FSEL', 'SEL'):
                cls['selects'] += k
            elif b in ('DSETP', 'FSETP', 'ISETP', 'PLOP3'):
                cls['compares'] += k
            elif b in ('BRA', 'BSSY', 'BSYNC', 'CALL', 'RET', 'EXIT', 'BAR', 'NOP', 'WARPSYNC'):
                cls['control'] += k
            else:
                cls['integer/other'] += k
        print(f'{name}: loop 0x{lo:x}..0x{hi:x}, {len(ops)} instructions')
        for k, v in cls.most_common():
            print(f'  {k:16s} {v:5d}')
        print('  ' + ', '.join(f'{op} {k}' for op, k in c.most_common(14)))


if __name__ == '__main__':
    main()
