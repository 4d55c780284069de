"""Summarise build/obj/*.ptxas.log: kernel, registers, stack, spills, smem."""
import glob, re, subprocess, sys
rows = []
for f in sorted(glob.glob('build/obj/*.ptxas.log')):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('risp::', '').replace('void ', '')
        rows.append((name, int(m.group(5)), int(m.group(2)), int(m.group(3)), m.group(7) or '0'))
flt = sys.argv[1] if len(sys.argv) > 1 else ''
for r in rows:
    if flt in r[0]:
        print('%-60s regs=%3d stack=%4d spill=%4d smem=%s' % r)
