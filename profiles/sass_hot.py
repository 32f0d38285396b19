"""Print per-SASS-instruction execution counts / stall samples from an `ncu --page source --csv --print-source sass` dump.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python sass_hot.py x.csv <kernel substring> [min_exec_frac]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sub = sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for bi, b in enumerate(blocks):
    if sub not in rows[b][1]:
        continue
    end = blocks[bi + 1] if bi + 1 < len(blocks) else len(rows)
    hdr = rows[b + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[b + 2:end] if len(r) > ix["Instructions Executed"]]
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    mx = max(int(r[ix["Instructions Executed"]]) for r in data)
    print("#", rows[b][1], "total warp-instr", tot, "instructions", len(data))
    for n, r in enumerate(data):
        ex = int(r[ix["Instructions Executed"]])
        if ex < thr * mx:
            continue
        print("%4d %10d s=%5s lsb=%5s noinst=%4s  %s" % (n, ex, r[ix["# Samples"]], r[ix["stall_long_sb"]], r[ix["stall_no_inst"]], r[ix["Source"]].strip()[:100]))
    break
