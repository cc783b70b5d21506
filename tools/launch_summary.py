"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel and grid:
    python tools/launch_summary.py gpurun_out/X.csv [first_kernel_regex last_kernel_regex] > profiles/X.md
With the two regexes only the launches from the first match of the first to the first later match of the second are kept
(one move: from the first libctmb kernel to the scale_kernel that ends it)."""
import csv
import re
import sys


def short(n):
    n = n.replace('void ', '').replace('ctmb::', '')
    return re.sub(r'\(.*', '', n)[:72]


def main():
    rows = []
    with open(sys.argv[1]) as f:
        for line in f:
            if line.startswith('"ID"'):
                break
        for r in csv.reader(f):
            if len(r) >= 15:
                rows.append((short(r[4]), r[7], r[8], float(r[14]), r[4]))
    a, b = 0, len(rows)
    if len(sys.argv) >= 4:
        first, last = re.compile(sys.argv[2]), re.compile(sys.argv[3])
        a = next(i for i, r in enumerate(rows) if first.search(r[4]))
        b = next(i for i, r in enumerate(rows) if i > a and last.search(r[4])) + 1
    seg = rows[a:b]
    tot = sum(r[3] for r in seg) / 1e6
    agg = {}
    for name, blk, grid, ns, _ in seg:
        s = agg.setdefault((name, grid, blk), [0, 0.0])
        s[0] += 1
        s[1] += ns / 1e6
    print(f'launches {len(seg)}, kernel time {tot:.1f} ms (cold-cache, serialised under ncu: compare shares)\n')
    print('| share | ms | launches | avg us | kernel | grid | block |')
    print('|---|---|---|---|---|---|---|')
    for (name, grid, blk), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ms / tot < 0.001:
            continue
        print(f'| {100 * ms / tot:.2f}% | {ms:.2f} | {c} | {1e3 * ms / c:.1f} | `{name}` | {grid} | {blk} |')


if __name__ == '__main__':
    main()
