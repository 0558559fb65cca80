#!/usr/bin/env python
"""SASS-level hot spots of the first kernel in an .ncu-rep captured with --import-source on:
    python tools/ncu_hotspots.py rep [top_n]
prints the instructions with the most warp-stall samples and their dominant stall reasons."""
import csv
import subprocess
import sys


def main(rep, top_n=40):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'Kernel Name' and hdr is not None:
            break
        if r and r[0] == 'Address':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ix = {k: i for i, k in enumerate(hdr)}
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print('instructions', len(data), 'samples', tot)
    agg = {}
    for r in data:
        for k in hdr:
            if k.startswith('stall_') and 'Not' not in k and r[ix[k]].isdigit():
                agg[k] = agg.get(k, 0) + int(r[ix[k]])
    print(sorted(agg.items(), key=lambda x: -x[1])[:10])
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top_n]:
        n = int(r[ix['# Samples']])
        st = {k: int(r[ix[k]]) for k in hdr if k.startswith('stall_') and 'Not' not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
        print(data.index(r), n, r[ix['Source']].strip()[:64], sorted(st.items(), key=lambda x: -x[1])[:3])


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
