"""Per-SASS-range totals (samples, executed warp instructions) of one kernel in an .ncu-rep, split at barriers.
   python tools/ncu_ranges.py rep kernel-regex [launch-skip]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > ix["# Samples"]:
        data.append(r)
ie = ix.get("Instructions Executed") or ix.get("Warp Instructions Executed")
seg, start = [], 0
def flush(end):
    s = sum(int(r[ix["# Samples"]]) for r in data[start:end]); n = sum(int(r[ie]) for r in data[start:end])
    seg.append((start, end, s, n))
for i, r in enumerate(data):
    if "BAR.SYNC" in r[ix["Source"]]:
        flush(i + 1); start = i + 1
flush(len(data))
ts = sum(s for _, _, s, _ in seg); tn = sum(n for _, _, _, n in seg)
print(f"total samples {ts}, warp instr {tn}")
for a, b, s, n in seg:
    if s * 200 > ts or n * 200 > tn:
        ops = {}
        for r in data[a:b]:
            op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[ie])
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:8]
        print(f"#{a:4d}-{b:4d}: samples {100*s/ts:5.1f}%  instr {100*n/tn:5.1f}%   " + " ".join(f"{k}:{100*v/tn:.1f}" for k, v in top))
