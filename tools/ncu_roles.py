"""Where do the warps of a warp-specialised kernel spend their time?  Reads an `ncu --set full --import-source on` report and
prints (1) the share of stall samples per window of SASS instructions together with the marker opcodes found there (LDGSTS =
gather producers, UTCHMMA = issuing warp, LDTM/UBLKCP = epilogue, ...), (2) every mbarrier try_wait with its samples, and
(3) the executed-instruction mix.  This is how the round-1 bottlenecks of conv_tc_kernel were found: the single issuing warp
(~250 dependent instructions per ring stage) and then the per-stage path of the gather warps.

Usage: python tools/ncu_roles.py report.ncu-rep [window=50] [kernel-id=:::1]"""
import collections
import csv
import subprocess
import sys

MARK = ("SYNCS", "UTCHMMA", "UTCBAR", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "LDTM", "BAR", "LDG", "STG", "STS", "LDS", "REDUX",
        "ATOMG", "ATOMS", "ELECT", "EXIT")


def opcode(src):
    sp = src.split()
    if not sp:
        return ""
    return (sp[1] if sp[0].startswith("@") and len(sp) > 1 else sp[0]).split(".")[0]


def main(rep, window=50, kid=":::1"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", kid],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    h, d = rows[1], rows[2:]
    i_s, i_src, i_e = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stalls = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[i_s] or 0) for r in d) or 1
    print(f"{len(d)} SASS instructions, {tot} stall samples\n\n# samples per window of {window} instructions")
    for w0 in range(0, len(d), window):
        seg = d[w0:w0 + window]
        s = sum(int(r[i_s] or 0) for r in seg)
        if s < tot * 0.005:
            continue
        ops = collections.Counter(o for o in (opcode(r[i_src]) for r in seg) if o in MARK)
        st = sorted(((sum(int(r[i] or 0) for r in seg), n) for i, n in stalls), reverse=True)[:3]
        ex = max(int(r[i_e] or 0) for r in seg)
        print(f"{w0:5d}  {100 * s / tot:5.1f}%  max exec {ex:9d}  {dict(ops)}  top stalls {[(n, c) for c, n in st]}")
    print("\n# mbarrier waits (samples on the try_wait and on the branch after it)")
    for k, r in enumerate(d):
        if "TRYWAIT" in r[i_src] and int(r[i_e] or 0) > 0:
            nxt = int(d[k + 1][i_s] or 0) if k + 1 < len(d) else 0
            print(f"{k:5d}  samples {int(r[i_s] or 0) + nxt:6d}  executed {int(r[i_e]):9d}  {r[i_src].strip()[:90]}")
    mix = collections.Counter()
    for r in d:
        mix[opcode(r[i_src])] += int(r[i_e] or 0)
    n = sum(mix.values()) or 1
    print("\n# executed warp-instructions by opcode")
    for op, c in mix.most_common(16):
        print(f"  {op:10s} {c:12d} {100 * c / n:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50, sys.argv[3] if len(sys.argv) > 3 else ":::1")
