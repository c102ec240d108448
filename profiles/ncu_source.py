#!/usr/bin/env python
"""Source-level view of one profiled launch: `python profiles/ncu_source.py x.ncu-rep [buckets]`.
Walks the SASS of the kernel in program order, cut into equal buckets, and prints per bucket the executed
instructions per warp, the share of stall samples, the top stall reasons and the dominant opcodes -- enough to
see WHERE a long straight-line kernel spends its time (the .ncu-rep files are too large to bring back)."""
import collections
import csv
import io
import subprocess
import sys


def opcode(sass):
    t = sass.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def main():
    rep = sys.argv[1]
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # one or more kernels: each starts with a "Kernel Name" row followed by a header row
    i = 0
    while i < len(rows):
        if not rows[i] or rows[i][0] != "Kernel Name":
            i += 1
            continue
        name, hdr = rows[i][1], rows[i + 1]
        j = i + 2
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            j += 1
        data = rows[i + 2:j]
        i = j
        ix = {h: k for k, h in enumerate(hdr)}
        S, IE = ix["# Samples"], ix["Instructions Executed"]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[S]) for r in data) or 1
        warps = int(data[0][IE]) or 1   # the first instruction is executed once by every warp
        sig = (name, len(data), tot)
        if sig == getattr(main, "last", None):   # ncu lists some results twice
            continue
        main.last = sig
        print(f"=== {name[:110]}")
        print(f"    {len(data)} SASS instructions, {sum(int(r[IE]) for r in data) / warps:.0f} executed per warp, {tot} stall samples")
        allst = collections.Counter()
        for r in data:
            for h in stalls:
                allst[h] += int(r[ix[h]])
        print("    stall samples: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in allst.most_common(8)))
        n = len(data)
        for b in range(nb):
            seg = data[b * n // nb:(b + 1) * n // nb]
            s = sum(int(r[S]) for r in seg)
            ie = sum(int(r[IE]) for r in seg) / warps
            st = collections.Counter()
            for r in seg:
                for h in stalls:
                    st[h] += int(r[ix[h]])
            ops = collections.Counter(opcode(r[1]) for r in seg).most_common(3)
            if ie == 0 and s == 0:
                continue
            print(f"    [{b * n // nb:5d}..] {ie:7.0f} inst/warp {100 * s / tot:5.1f}% of samples  "
                  + " ".join(f"{k[6:]}:{100 * v / max(s, 1):.0f}%" for k, v in st.most_common(3))
                  + "   " + " ".join(f"{o}x{c}" for o, c in ops))


if __name__ == "__main__":
    main()
