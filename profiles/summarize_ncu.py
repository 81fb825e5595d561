import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.max','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor.sum','lts__t_bytes.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg']
for i,h in enumerate(hdr):
    if h in want: print(f"{h:70s} {units[i]:12s} {vals[i]}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,"# Samples") for r in data)
print("total samples",tot,"instr rows",len(data))
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sec=int(sys.argv[2]) if len(sys.argv)>2 else 150
for s in range(0,len(data),sec):
    chunk=data[s:s+sec]
    n=sum(f(r,"# Samples") for r in chunk)
    if n==0: continue
    ex=sum(f(r,"Instructions Executed") for r in chunk)
    st={k:sum(f(r,k) for r in chunk) for k in stalls}
    top=sorted(st.items(), key=lambda kv:-kv[1])[:4]
    ops={}
    for r in chunk:
        t=r[ix["Source"]].split()
        if not t: continue
        op=t[1] if t[0].startswith('@') and len(t)>1 else t[0]
        op=op.split('.')[0]
        ops[op]=ops.get(op,0)+1
    topops=sorted(ops.items(), key=lambda kv:-kv[1])[:4]
    print(f"{s:5d} smp {n:6.0f} ({100*n/tot:4.1f}%) exec {ex/1e6:7.2f}M {[(k[6:],int(v)) for k,v in top]} {topops}")
