#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS mnemonics in the in-tree library (evidence for profiles/):

    python tools/sass_histogram.py [to_ued_b200/libtoued.so] > profiles/r02_sass_histogram.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA 1-D), UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be absent), LDGSTS = cp.async."""
import collections
import re
import subprocess
import sys

MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDGSTS", "MUFU", "F2FP"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "to_ued_b200/libtoued.so"
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            if op in MNEMONICS:
                counts[cur][op] += 1
    print(f"# SASS mnemonic histogram of {lib} (cuobjdump -sass, sm_100a); columns: " + " ".join(MNEMONICS) + " | total instructions")
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        if not any(c[m] for m in MNEMONICS if m not in ("MUFU", "F2FP", "LDGSTS", "SYNCS")):
            continue
        print(f"{k[:70]:70s} " + " ".join(f"{c[m]:5d}" for m in MNEMONICS) + f" | {c['_total']:6d}")
        tot.update(c)
    print(f"{'ALL LISTED KERNELS':70s} " + " ".join(f"{tot[m]:5d}" for m in MNEMONICS) + f" | {tot['_total']:6d}")
    allk = collections.Counter()
    for c in counts.values():
        allk.update(c)
    print(f"{'WHOLE LIBRARY (' + str(len(order)) + ' kernels)':70s} " + " ".join(f"{allk[m]:5d}" for m in MNEMONICS) + f" | {allk['_total']:6d}")
    assert allk["HMMA"] == 0, "legacy mma.sync found"


if __name__ == "__main__":
    main()
