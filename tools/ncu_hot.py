"""(runs here) hottest SASS instructions of one kernel in an .ncu-rep captured with --set full:
   python tools/ncu_hot.py REP KERNEL_REGEX [TOP [WHICH]]   (WHICH: index among the matching launches)   ->  address, samples, dominant stall reasons, SASS text
Used to find what bounds a kernel whose time is neither tensor- nor HBM-limited."""
import csv, io, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
chunks = out.split('"Kernel Name",')
for ch in chunks[1 + which:2 + which]:
    lines = ch.split("\n")
    print("kernel:", lines[0][:120])
    rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
    rows = [r for r in rd if r.get("Address")]
    stalls = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
    print("total samples", tot, "instructions", len(rows))
    print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    # samples between synchronisation points (mbarrier waits/arrives, named barriers, MMA issue): which role waits on which
    marks = ("SYNCS", "BAR.", "UTCHMMA", "UTCBAR", "UTMALDG", "EXIT")
    start = 0
    for i, r in enumerate(rows):
        if any(m in r["Source"] for m in marks):
            seg = sum(int(x["# Samples"] or 0) for x in rows[start:i + 1])
            if seg * 200 >= tot:
                print("  seg rows %4d-%4d %5.1f%%  ends at: %s (exec %s)" % (start, i, 100.0 * seg / max(tot, 1), r["Source"][:60], r["Instructions Executed"]))
            start = i + 1
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        s = int(r["# Samples"] or 0)
        dom = sorted(((int(r[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print("%5d %6d %5.1f%%  %-28s exec=%-8s %s" % (i, s, 100.0 * s / max(tot, 1), " ".join("%s:%d" % (k, v) for v, k in dom if v),
                                                     r["Instructions Executed"], r["Source"][:90]))
