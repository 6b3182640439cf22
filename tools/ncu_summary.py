import csv,sys,collections,subprocess,re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
d={h:(u,v) for h,u,v in zip(hdr,units,vals)}
def g(k): return d.get(k,('',''))
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','sm__cycles_elapsed.avg','lts__t_bytes.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys: print(k, g(k))
for k in sorted(d):
    if 'smsp__average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        v=float(d[k][1]); 
        if v>0.05: print('  stall', k.split('stalled_')[1].split('_per_issue')[0], v)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
cur=None; agg=collections.Counter(); samp=collections.Counter(); text={}; ops=collections.Counter()
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>=8 and r[0] not in ('','Line No','Function Name') and r[2]=='-':
        try:
            key=(cur,int(r[0])); agg[key]+=int(r[7]); samp[key]+=int(r[6]); text[key]=r[1].strip()[:80]
        except: pass
    elif len(r)>=8 and r[0]=='' and r[2].startswith('0x'):
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
        try: ops[m.group(2)]+=int(r[7])
        except: pass
tot=sum(agg.values()); ts=sum(samp.values())
print('total warp-inst (line agg)',tot)
for k,n in agg.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 30): print(f"{k[0]:16s}:{k[1]:4d} {n/tot*100:5.2f}% samp {samp[k]/ts*100:5.2f}%  {text[k]}")
to=sum(ops.values())
print(' '.join(f"{o}:{n/to*100:.1f}%" for o,n in ops.most_common(22)))
