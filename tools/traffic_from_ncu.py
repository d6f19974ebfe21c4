"""Summarise an ncu launch list (gpu__time_duration.sum + dram__bytes_{read,write}.sum, --csv) per kernel.

    python tools/traffic_from_ncu.py profiles/r01_launches_bench_resnet50.csv > profiles/traffic.json

bench.py reads `dram_bytes_per_launch` of the dominant kernel from profiles/traffic.json for roofline.traffic.
"""
import csv
import json
import re
import sys


def kernel_key(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'\b\w+::', '', name)
    return name.replace(' ', '')


def main(path):
    rows = {}
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(r['ID'], {'kernel': kernel_key(r['Kernel Name'])})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        if r['Metric Name'] == 'gpu__time_duration.sum':
            d['ns'] = v * {'ns': 1, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 'nsecond': 1, 'second': 1e9, 's': 1e9}[unit]
        else:
            d['bytes'] = d.get('bytes', 0.0) + v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    total_ns = sum(d.get('ns', 0.0) for d in rows.values())
    out = {}
    for d in rows.values():
        o = out.setdefault(d['kernel'], {'launches': 0, 'ns': 0.0, 'bytes': 0.0})
        o['launches'] += 1
        o['ns'] += d.get('ns', 0.0)
        o['bytes'] += d.get('bytes', 0.0)
    res = {k: {'launches': o['launches'], 'time_share': o['ns'] / total_ns, 'avg_us_under_ncu': o['ns'] / o['launches'] / 1e3,
               'dram_bytes_per_launch': o['bytes'] / o['launches']}
           for k, o in sorted(out.items(), key=lambda kv: -kv[1]['ns'])}
    res['source'] = '%s: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none' % path
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main(sys.argv[1])
