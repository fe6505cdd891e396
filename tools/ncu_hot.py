#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples for one kernel of an .ncu-rep (source page, no GPU needed).

  python tools/ncu_hot.py gpurun_out/prof.ncu-rep dkdv [top_n]
"""
import csv
import subprocess
import sys


def main(path, pattern, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{pattern}",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # one section per profiled launch: a "Kernel Name" row, a header row, then one row per SASS instruction
    starts = [n for n, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    lo = starts[0]
    hi = starts[1] if len(starts) > 1 else len(rows)
    hdr = rows[lo + 1]
    ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
    total = sum(int(r[isamp] or 0) for r in body)
    print(f"kernel ~{pattern}: {len(body)} SASS instructions, {total} samples")
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
    order = sorted(range(len(body)), key=lambda n: -int(body[n][isamp] or 0))[:top]
    for n in sorted(order):
        r = body[n]
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
        sts = " ".join(f"{k}={v}" for v, k in st if v)
        print(f"{n:5d} {int(r[isamp] or 0):6d} {100.0 * int(r[isamp] or 0) / max(total, 1):5.1f}%  x{r[iexec]:>9s}  {r[isrc].strip()[:70]:70s} {sts}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
