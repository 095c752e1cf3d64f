"""Summarise an ncu raw/source CSV pair: python scripts/ncu_summary.py raw.csv src.csv [row]"""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
rowi = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rows=list(csv.reader(open(raw)))
hdr=rows[0]; r=rows[rowi]
for k in ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']:
    if k in hdr: print(k, r[hdr.index(k)][:80])
st=[(float(r[i]) if r[i] else 0,hdr[i]) for i in range(len(hdr)) if 'issue_stalled' in hdr[i] and 'ratio' in hdr[i] and 'not_issued' not in hdr[i]]
for v,n in sorted(st,reverse=True)[:8]: print(f"{v:8.2f} {n}")
rows=list(csv.reader(open(src)))
hdr=rows[1]; body=[r for r in rows[2:] if len(r)>5]
si=hdr.index("Warp Stall Sampling (All Samples)"); ii=hdr.index("Instructions Executed")
def I(x):
    try: return int(x)
    except: return 0
addrs=[r[0] for r in body]
n=addrs.index(addrs[0],1) if addrs[0] in addrs[1:] else len(body)
body=body[:n]
tot=sum(I(r[si]) for r in body); print("total samples",tot,"ninstr",n, "inst exec", sum(I(r[ii]) for r in body))
top=sorted(enumerate(body), key=lambda x:-I(x[1][si]))[:36]
for i,r in sorted(top): print(i, r[si], r[ii], r[1].strip()[:100])
for b in range(0,n,100):
    sm=sum(I(r[si]) for r in body[b:b+100]); ie=sum(I(r[ii]) for r in body[b:b+100]); print(b, f"{100*sm/max(tot,1):.1f}% samples", f"{ie/1e6:.1f}M inst")
