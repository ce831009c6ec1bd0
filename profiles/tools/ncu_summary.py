"""Summarise an ncu report of the fused step kernel: key raw metrics, stall-reason shares, and dynamic instruction counts
per CUDA source line (SASS offsets from ncu --page source joined with nvdisasm -g line info of the in-tree library).

    python profiles/tools/ncu_summary.py gpurun_out/<report>.ncu-rep [mangled-symbol-after-.text.] [cubin-stem] [units-per-launch]

e.g. ... r1_loader_full.ncu-rep _ZN3phc19build_tables_kernel build_tables 2478461   (per-frame counts for the table build)
"""
import csv, collections, re, sys, subprocess, glob
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else '_ZN3phc17step_fused_kernelILb1ELb0'
cubin=sys.argv[3] if len(sys.argv)>3 else 'step_fused'
N=int(sys.argv[4]) if len(sys.argv)>4 else 65536
subprocess.run(f'ncu -i {rep} --page raw --csv > /tmp/raw.csv 2>/dev/null', shell=True)
subprocess.run(f'ncu -i {rep} --page source --csv > /tmp/src.csv 2>/dev/null', shell=True)
rows=list(csv.reader(open('/tmp/raw.csv'))); hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
r=rows[2]
for w in want:
    if w in hdr: i=hdr.index(w); print(f'{w:70s} {r[i]} {units[i]}')
# sass
subprocess.run('rm -rf /tmp/cub; mkdir -p /tmp/cub; cd /tmp/cub; cuobjdump -xelf all /root/repo/puffer_phc_b200/lib/libphc_b200.so >/dev/null 2>&1; nvdisasm -g -c '+cubin+'.sm_100a.cubin > /tmp/step_sass.txt', shell=True)
lines=open('/tmp/step_sass.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+kern)][0]
off2src={}; cur=None
for l in lines[start+1:]:
    if l.startswith('.text.') : break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: off2src[int(m.group(1),16)]=(cur, m.group(2))
rows=list(csv.reader(open('/tmp/src.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); ia=hdr.index('Source')
data=[]
for r in rows[hi+1:]:
    if r and r[0]=='Kernel Name': break
    if len(r)>ie: data.append(r)
base=int(data[0][0],16)
per=collections.Counter(); samp=collections.Counter(); tot=0
stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
st=collections.Counter()
for r in data:
    off=int(r[0],16)-base
    src=off2src.get(off,(None,''))[0]
    n=int(r[ie]); per[src]+=n; samp[src]+=int(r[isamp]); tot+=n
    for i,h in stall_cols:
        try: st[h]+=int(r[i])
        except: pass
print('total warp-instr/unit', tot/N)
ss=sum(st.values())
print('stalls:', ', '.join(f'{h[6:]} {100*c/ss:.1f}%' for h,c in st.most_common(9)))
ts=sum(samp.values())
cache={}
def srcline(k):
    if not k: return ''
    f,l=k
    for p in glob.glob('/root/repo/puffer_phc_b200/csrc/'+f):
        if p not in cache: cache[p]=open(p).read().split('\n')
        return cache[p][l-1].strip()[:95]
    return ''
print('--- top by samples')
for k,n in samp.most_common(25): print(f'{100*n/ts:5.1f}%  {per[k]/N:7.1f}/env  {k}  {srcline(k)}')
print('--- top by executed warp instructions')
for k,n in per.most_common(40): print(f'{n/N:8.1f}/unit  {100*samp[k]/ts:5.1f}% of samples  {k}  {srcline(k)}')
print('--- top sass by samples')
top=sorted(data,key=lambda r:-int(r[isamp]))[:25]
for r in top: 
    off=int(r[0],16)-base
    print(f'{100*int(r[isamp])/ts:5.2f}%  exec/env {int(r[ie])/N:6.2f}  {r[ia].strip()[:70]:70s} {off2src.get(off,(None,))[0]}')
