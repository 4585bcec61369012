"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python profiles/launch_list.py launches.csv"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = {}
for r in rows:
    name, val, unit = r[4], float(r[14].replace(",", "")), r[13]
    us = val / 1e3 if unit in ("ns", "nsecond") else val * 1e3 if unit in ("ms", "msecond") else val
    m = re.search(r"(?:dcdf::)?\b(kb?_\w+(?:<[^(]*>)?)", name)  # this library's kernels are k_* / kb_* (with or without the namespace)
    key = "dcdf::" + m.group(1) if m else "torch kernels (synthetic generator / comparison, not part of the path)"
    a = agg.setdefault(key, [0.0, 0])
    a[0] += us; a[1] += 1
tot = sum(v[0] for k, v in agg.items() if k.startswith("dcdf::")) or 1.0
for k, (us, n) in sorted(agg.items(), key=lambda kv: (kv[0].startswith("dcdf::"), -kv[1][0])):
    share = f"{100 * us / tot:5.1f}%" if k.startswith("dcdf::") else "      "
    print(f"{us:12.1f} us  {share}  x{n:<5d} {k}")
