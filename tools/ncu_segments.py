"""Split the SASS of a profiled kernel into phases at its barrier / mbarrier / TMA
instructions and report warp-stall samples and instruction mix per phase.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_segments.py src.csv <warp-rows per launch>
"""
import collections
import csv
import sys


def main(path, units):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seg = 0
    segs = collections.OrderedDict()
    tot = 0
    for r in rows[2:]:
        if len(r) < len(hdr) - 5:
            continue
        src = r[ix['Source']].strip()
        sm = int(r[ix['# Samples']])
        ex = int(r[ix['Instructions Executed']])
        d = segs.setdefault(seg, collections.Counter())
        d['samples'] += sm
        d['inst'] += ex
        tot += sm
        toks = src.split()
        op = toks[1] if src.startswith('@') and len(toks) > 1 else toks[0]
        for key, pre in (('fp64', ('DADD', 'DMUL', 'DFMA', 'DSETP')), ('lds', ('LDS',)), ('sts', ('STS',)), ('ldg', ('LDG',)),
                         ('stg', ('STG',)), ('f2f', ('F2F',)), ('bra', ('BRA', 'BSSY', 'BSYNC')), ('shfl', ('SHFL',)),
                         ('fp32', ('FFMA', 'FMUL', 'FADD'))):
            if op.startswith(pre):
                d[key] += ex
        if op.startswith('BAR') or 'SYNCS' in op or op.startswith('UBLKCP'):
            d['end'] = src[:50]
            seg += 1
    print("total samples", tot)
    for k, d in segs.items():
        if d['samples'] * 200 > tot:
            print('%2d %5.1f%% inst %4.0f | fp64 %3.0f fp32 %3.0f lds %2.0f sts %2.0f ldg %2.0f stg %2.0f f2f %2.0f bra %2.0f shfl %2.0f | %s' % (
                k, 100 * d['samples'] / tot, d['inst'] / units, d['fp64'] / units, d['fp32'] / units, d['lds'] / units,
                d['sts'] / units, d['ldg'] / units, d['stg'] / units, d['f2f'] / units, d['bra'] / units, d['shfl'] / units,
                d.get('end', '')))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
