import csv, collections, re, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr=rows[0]; data=rows[2:]
want = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__t_sector_hit_rate.pct","smsp__inst_executed.sum","smsp__warps_eligible.avg.per_cycle_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(w, rows[1][i], [r[i] for r in data])
mixed = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(mixed.splitlines()))
src = open("/root/repo/harmonic_power_flow_b200/csrc/hpf_structured.cuh").read().splitlines()
# phase boundaries by marker comments
marks=[]
for i,l in enumerate(src,1):
    if "// =================" in l and "harm" not in l: marks.append((i,l.strip()[:40]))
    if "// t_z = u0_z" in l: marks.append((i,"C2: t_z rows"))
    if "// border system for the fundamental" in l: marks.append((i,"C3: border P5"))
    if "// ---- initial fill ----" in l: marks.append((i,"init"))
    if "post-processing (HG:547-549) + write-out of the finished" in l: marks.append((i,"C1: finalize"))
start=[m for m in marks if m[1]=="init"][-1][0]
marks=[m for m in marks if m[0]>=start]
def phase(f,ln):
    if f!='hpf_structured.cuh': return 'inlined helpers'
    cur='pre'
    for a,name in marks:
        if ln>=a: cur=name
    return cur
agg=collections.defaultdict(lambda:[0,0,collections.Counter()])
cur_file=None; cur_line=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No':
        hd=r; ii=hd.index('Instructions Executed'); si=hd.index('Warp Stall Sampling (All Samples)')
        cols=[(i,h) for i,h in enumerate(hd) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if r[0]!='': cur_line=(cur_file,int(r[0]))
    else:
        ph=phase(*cur_line)
        try: agg[ph][0]+=int(r[ii]); agg[ph][1]+=int(r[si])
        except: continue
        for i,h in cols:
            try: agg[ph][2][h[6:]]+=int(r[i])
            except: pass
ti=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
for k,(i,s,c) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    t=sum(c.values()) or 1
    print("%-42s inst %5.1f%% samples %5.1f%% | %s"%(k,100*i/ti,100*s/ts," ".join("%s %.0f%%"%(a,100*b/t) for a,b in c.most_common(4))))
tot=collections.Counter()
for v in agg.values(): tot.update(v[2])
T=sum(tot.values()); print("TOTAL stalls:", " ".join("%s %.1f%%"%(a,100*b/T) for a,b in tot.most_common(8)))
