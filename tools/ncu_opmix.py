"""Instruction mix (executed warp instructions by opcode) from `ncu -i X.ncu-rep --page source --csv`."""
import collections
import csv
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ie, src = hdr.index('Instructions Executed'), hdr.index('Source')
    tot, n = collections.Counter(), 0
    for r in rows[2:]:
        try:
            c = int(r[ie])
        except (ValueError, IndexError):
            continue
        toks = r[src].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        tot[op.split('.')[0]] += c
        n += c
    print('total warp instructions', n)
    for k, v in tot.most_common(top):
        print(f'{k:12s} {v:14d} {100 * v / n:5.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
