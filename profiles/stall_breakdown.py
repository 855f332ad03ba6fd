import csv, collections, sys, subprocess
rep, kern, skip = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern,"--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:80])
H = rows[1]
isrc, isamp, iex = H.index('Source'), H.index('Warp Stall Sampling (All Samples)'), H.index('Instructions Executed')
stalls = [(i,h) for i,h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter(); st = collections.Counter()
perop_st = collections.defaultdict(collections.Counter)
tot=0
for r in rows[2:]:
    if len(r) < len(H) or not r[isamp].isdigit(): continue
    op = r[isrc].strip().split()
    if not op: continue
    o = op[0] if not op[0].startswith('@') else op[1]
    o = o.split('.')[0]
    s = int(r[isamp] or 0); agg[o]+=s; tot+=s; cnt[o]+=1; ex[o]+=int(r[iex] or 0)
    for i,h in stalls:
        v=int(r[i] or 0); st[h]+=v; perop_st[o][h]+=v
for o,s in agg.most_common(12):
    top = ", ".join(f"{h[6:]}={v}" for h,v in perop_st[o].most_common(3))
    print(f"{o:10s} samples={s:7d} {s/tot:6.3f} static={cnt[o]:5d} executed={ex[o]:9d}  [{top}]")
print("total samples", tot, "total executed", sum(ex.values()))
for h,v in st.most_common(8): print(" ", h, v, round(v/tot,3))
