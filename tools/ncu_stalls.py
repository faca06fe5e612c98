"""Per-kernel warp-stall summary of an exported ncu source page.

    ncu -i report.ncu-rep --page source --csv --print-source sass > page.csv
    python tools/ncu_stalls.py page.csv <kernel substring> [top N lines]

Prints the stall-reason totals of the first matching kernel and its N most-sampled SASS lines (offsets relative to the
kernel's first instruction, so they line up with `cuobjdump -sass`)."""
import csv, io, sys, collections


def main():
    path, pat = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = open(path).read()
    for b in txt.split('"Kernel Name",')[1:]:
        lines = b.split('\n')
        if pat not in lines[0]:
            continue
        rdr = csv.reader(io.StringIO('\n'.join(lines[1:])))
        hdr = next(rdr)
        idx = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        rows = [r for r in rdr if len(r) == len(hdr)]
        n = sum(int(r[idx['# Samples']] or 0) for r in rows)
        tot = collections.Counter()
        for r in rows:
            for s in stalls:
                tot[s] += int(r[idx[s]] or 0)
        print(lines[0][:70], '| SASS lines', len(rows), '| samples', n)
        for s, v in tot.most_common(10):
            print(f'  {s:26s} {v:8d} {100 * v / max(n, 1):5.1f}%')
        a0 = int(rows[0][idx['Address']], 16)
        for r in sorted(rows, key=lambda r: -int(r[idx['# Samples']] or 0))[:topn]:
            st = sorted(((s, int(r[idx[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
            print(f"  +{int(r[idx['Address']], 16) - a0:#07x} {int(r[idx['# Samples']]):6d} exec {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']][:64]:64s} {st}")
        return
    print('no kernel matches', pat)


if __name__ == '__main__':
    main()
