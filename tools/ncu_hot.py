"""Top stall-sample SASS lines of one kernel launch in an .ncu-rep (source page).
   python tools/ncu_hot.py rep kernel-regex [launch-skip] [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
print(rows[hi - 1][1][:120] if hi else "")
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data if len(r) > ix["# Samples"])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for n, r in enumerate(data):
    if len(r) <= ix["# Samples"]:
        continue
    s = int(r[ix["# Samples"]])
    st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    lines.append((s, n, r[ix["Source"]].strip(), st))
print(f"total samples {tot}, {len(data)} SASS lines")
for s, n, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{100 * s / max(tot, 1):5.1f}%  #{n:4d}  {src[:70]:70s} {st}")
