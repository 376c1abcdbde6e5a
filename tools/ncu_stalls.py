"""Top stall sites from an ncu report's source page (SASS view)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall_cols}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
rank = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:topn]
for i in sorted(rank):
    r = data[i]
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {st}")
