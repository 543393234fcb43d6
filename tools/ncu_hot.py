"""Top SASS instructions by executed count / stall samples from `ncu --page source --csv`.
usage: python tools/ncu_hot.py rep.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    try:
        data.append((n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), r))
    except ValueError:
        pass
tot_i = sum(d[2] for d in data); tot_s = sum(d[3] for d in data)
print(f"total warp-instr {tot_i:.3e}, samples {tot_s}")
# contiguous regions: print cumulative by blocks of 25 instructions
blk = 25
print("-- by block of SASS lines (line range: inst share, sample share)")
for b in range(0, len(data), blk):
    ii = sum(d[2] for d in data[b:b+blk]); ss = sum(d[3] for d in data[b:b+blk])
    if ii / max(tot_i,1) > 0.01 or ss / max(tot_s,1) > 0.01:
        print(f"  {b:5d}-{b+blk-1:5d}: inst {100*ii/tot_i:5.1f}%  samples {100*ss/max(tot_s,1):5.1f}%   first: {data[b][1][:60]}")
print("-- top by samples")
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for d in sorted(data, key=lambda d: -d[3])[:top]:
    st = sorted(((int(d[4][ix[k]] or 0), k) for k in stall_cols), reverse=True)[:2]
    print(f"  L{d[0]:5d} inst {100*d[2]/tot_i:5.2f}% samp {100*d[3]/max(tot_s,1):5.2f}%  {d[1][:70]:70s} {st}")
