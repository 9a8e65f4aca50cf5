"""Per-kernel digest of `ncu -i REP --page source --csv`: opcode mix weighted by executed count, stall reasons,
and the most-sampled instructions.   python scripts/ncu_source.py REP.ncu-rep [kernel-substring] [top-N]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def kernels(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "body": []}
            out.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["body"].append(r)
    return out


def digest(k, top=30):
    hdr, body = k["hdr"], k["body"]
    isrc, isamp, iex, ia = (hdr.index(x) for x in ("Source", "# Samples", "Instructions Executed", "Address"))
    num = lambda v: int(float(v)) if v not in ("", None) else 0
    tot, totex = sum(num(r[isamp]) for r in body), sum(num(r[iex]) for r in body)
    print("==", k["name"][:100])
    print("static instrs %d  executed warp-instrs %d  samples %d" % (len(body), totex, tot))
    c, s = Counter(), Counter()
    for r in body:
        toks = r[isrc].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        c[op] += num(r[iex])
        s[op] += num(r[isamp])
    print("opcode mix (executed, share, samples):")
    for op, v in c.most_common(22):
        print("  %-12s %10d %5.1f%% %7d" % (op, v, 100.0 * v / max(totex, 1), s[op]))
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = Counter()
    for r in body:
        for i in stall:
            try:
                agg[hdr[i]] += float(r[i] or 0)
            except ValueError:
                pass
    print("stalls:", ", ".join("%s %d" % (a[6:], b) for a, b in agg.most_common(8)))
    print("most sampled:")
    for r in sorted(body, key=lambda r: -num(r[isamp]))[:top]:
        st = sorted(((hdr[i][6:], float(r[i] or 0)) for i in stall), key=lambda x: -x[1])[0]
        print("  %s %-70s %6s %9s %s" % (r[ia][-5:], r[isrc][:70], r[isamp], r[iex], st[0]))


if __name__ == "__main__":
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    for k in kernels(sys.argv[1]):
        if sub in k["name"]:
            digest(k, top)
