"""Static instruction count of a kernel's loops from its SASS (no GPU needed).

    python scripts/sass_loop.py build/obj/risp_pipeline.o <mangled-name-substring> [px_per_iteration] [loop index]

Disassembles with cuobjdump, finds every backward branch (= a loop), and prints an opcode histogram of the
largest loop body (all instructions between the branch target and the branch), plus per-pipe totals:
fma (FFMA/FMUL/FADD and their packed forms; packed forms count twice in `fma lane-ops`), alu, xu (MUFU), lsu.
Straight-line loop bodies only: instructions on both sides of a forward branch inside the body are all counted.
"""
import collections
import re
import subprocess
import sys

FMA = {'FFMA', 'FMUL', 'FADD', 'FFMA2', 'FMUL2', 'FADD2', 'IMAD', 'HFMA2', 'HADD2', 'HMUL2'}
XU = {'MUFU', 'F2I', 'I2F', 'F2F', 'POPC', 'FLO', 'BREV', 'I2FP', 'F2FP', 'F2IP'}
LSU = {'LDG', 'STG', 'LDS', 'STS', 'LD', 'ST', 'LDL', 'STL', 'SHFL', 'ATOMS', 'ATOMG', 'RED', 'LDSM'}
UNI = {'LDCU', 'UMOV', 'UIADD3', 'ULOP3', 'UISETP', 'USEL', 'ULEA', 'UIMAD', 'USHF', 'R2UR', 'UFLO', 'UPOPC', 'S2UR',
       'UPRMT', 'ULDC', 'UFMUL', 'UFFMA', 'UFADD', 'UFSETP', 'UFSEL', 'UFMNMX', 'UF2F', 'UI2FP', 'UF2FP', 'UF2IP', 'UI2F'}


def disasm(obj, name):
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r'\n\s*Function : ', txt)
    hits = [f for f in funcs[1:] if name in f.split('\n', 1)[0]]
    if not hits:
        raise SystemExit('no function matching %r; have:\n%s' % (name, '\n'.join(f.split('\n', 1)[0] for f in funcs[1:])))
    if len(hits) > 1:
        sys.stderr.write('%d matches, using the first: %s\n' % (len(hits), hits[0].split('\n', 1)[0]))
    return hits[0]


def parse(body):
    ins = []
    for line in body.split('\n'):
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        text = m.group(2).strip()
        mm = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', text)
        op = mm.group(2) if mm else '?'
        ins.append((addr, op, text))
    return ins


def loops(ins):
    out = []
    for addr, op, text in ins:
        if op == 'BRA':
            m = re.search(r'0x([0-9a-f]+)', text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= addr:
                    out.append((tgt, addr))
    return out


def main():
    obj, name = sys.argv[1], sys.argv[2]
    npx = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
    f = disasm(obj, name)
    print(f.split('\n', 1)[0])
    ins = parse(f)
    lp = loops(ins)
    print('instructions total %d, loops %s' % (len(ins), [(hex(a), hex(b), sum(1 for x in ins if a <= x[0] <= b)) for a, b in lp]))
    if not lp:
        return
    a, b = lp[int(sys.argv[4])] if len(sys.argv) > 4 else max(lp, key=lambda t: t[1] - t[0])
    body = [x for x in ins if a <= x[0] <= b]
    cnt = collections.Counter(x[1] for x in body)
    n = len(body)
    fma_slots = sum(c for o, c in cnt.items() if o in FMA)
    fma_lane = sum(c * (2 if o.endswith('2') and o[0] == 'F' else 1) for o, c in cnt.items() if o in FMA)
    xu = sum(c for o, c in cnt.items() if o in XU)
    lsu = sum(c for o, c in cnt.items() if o in LSU)
    uni = sum(c for o, c in cnt.items() if o in UNI or o.startswith('U'))
    alu = n - fma_slots - xu - lsu - uni - cnt.get('BRA', 0) - cnt.get('NOP', 0)
    print('largest loop: %d instr per iteration = %.1f per px (at %g px/iteration)' % (n, n / npx, npx))
    print('  per px: issue %.1f | fma slots %.1f (lane-ops %.1f) | alu-ish %.1f | xu %.1f | lsu %.1f | uniform %.1f' %
          (n / npx, fma_slots / npx, fma_lane / npx, alu / npx, xu / npx, lsu / npx, uni / npx))
    for o, c in cnt.most_common(40):
        print('  %-10s %5d  %5.1f%%  per-px %.2f' % (o, c, 100.0 * c / n, c / npx))


if __name__ == '__main__':
    main()
