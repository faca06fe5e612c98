"""Per-kernel launch counts / total time / share from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <command>`).

    python tools/launch_shares.py launches.csv > profiles/<round>_launch_shares.txt"""
import csv, sys, collections


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        name = r[ik].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# total {total / 1e3:.3f} ms in {sum(cnt.values())} launches")
    print("kernel\tlaunches\ttotal_us\tshare")
    for k, v in tot.most_common():
        print(f"{k[:110]}\t{cnt[k]}\t{v:.1f}\t{100 * v / total:.1f}%")


if __name__ == "__main__":
    main()
