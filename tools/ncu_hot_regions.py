"""Stall-sample profile of one kernel from an .ncu-rep (source page, SASS view): total samples and executed
instructions per contiguous region between barriers / branches, plus the top instructions.
usage: python tools/ncu_hot_regions.py report.ncu-rep 'kernel regex' [top_n]"""
import csv, io, subprocess, sys
rep, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + key], capture_output=True, text=True).stdout
lines = txt.splitlines()
# first kernel instance only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
h = rows[0]
ia, isrc, isamp, iexe = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in rows[1:])
totexe = sum(int(r[iexe]) for r in rows[1:])
print("kernel", key, "samples", tot, "warp-instructions", totexe)
reg, acc, k = [], [0, 0, 0, ""], 0
for r in rows[1:]:
    s, e, src = int(r[isamp]), int(r[iexe]), r[isrc].strip()
    acc[0] += s; acc[1] += e; acc[2] += 1
    if not acc[3]: acc[3] = "%d: %s" % (k, src[:50])
    if any(t in src for t in ("BAR.SYNC", "BRA", "EXIT", "SYNCS", "WARPSYNC")):
        reg.append(acc + ["%d: %s" % (k, src[:60])]); acc = [0, 0, 0, ""]
    k += 1
if acc[2]: reg.append(acc + ["end"])
print("regions (samples%, warp-instr%, n static) first .. last")
for s, e, n, first, last in reg:
    if s * 200 >= tot or e * 200 >= totexe:
        print("  %5.1f%%  %5.1f%%  %4d   %-58s .. %s" % (100.0 * s / tot, 100.0 * e / max(totexe, 1), n, first, last))
print("top instructions by samples")
for r in sorted(rows[1:], key=lambda r: -int(r[isamp]))[:top]:
    print("  %5.1f%%  exe %9s  %s" % (100.0 * int(r[isamp]) / tot, r[iexe], r[isrc].strip()[:100]))
