#!/usr/bin/env python
"""Instruction mix of the loops of one kernel in libirb_b200.so (cuobjdump -sass): for every backward branch, the loop body
size and a histogram of opcodes -- used to reason about FMA-pipe / issue-slot budgets without a GPU.
  python tools/sass_loops.py 'render_sg_group_kernelILi1ELi2ELb1ELb1ELi2E' [min_body]"""
import collections
import re
import subprocess
import sys

LIB = 'inverserenderingofindoorscene_b200/libirb_b200.so'


def functions(lib=LIB):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res


def opcode(ins):
    t = ins.split()
    if t[0].startswith('@'):
        t = t[1:]
    return t[0].split('.')[0] if t else ''


def fullop(ins):
    t = ins.split()
    if t[0].startswith('@'):
        t = t[1:]
    return t[0]


def main():
    pat = sys.argv[1]
    minb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    for name, ins in functions().items():
        if pat not in name:
            continue
        print(name, len(ins), 'instructions')
        addr = [a for a, _ in ins]
        for i, (a, s) in enumerate(ins):
            m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', s)
            if m and opcode(s) == 'BRA':
                tgt = int(m.group(1), 16)
                if tgt < a:
                    j = addr.index(tgt) if tgt in addr else None
                    if j is None or i - j < minb:
                        continue
                    body = ins[j:i + 1]
                    h = collections.Counter(opcode(x) for _, x in body)
                    print('  loop %#x..%#x: %d instr' % (tgt, a, len(body)))
                    print('    ' + '  '.join('%s %d' % kv for kv in h.most_common(30)))
                    h2 = collections.Counter(fullop(x) for _, x in body if opcode(x) in ('MUFU', 'SHFL', 'LDS', 'LDC', 'MOV', 'FSEL', 'SEL'))
                    print('    detail: ' + '  '.join('%s %d' % kv for kv in h2.most_common(20)))


if __name__ == '__main__':
    main()
