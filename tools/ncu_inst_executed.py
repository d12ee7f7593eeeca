"""Executed-instruction counts per SASS line from an ncu source-page CSV (ncu -i rep --page source --csv [| gzip]), per warp
and iteration: FP64 against other instructions, the top opcodes, and (with a fourth argument) the plateaus of the execution
count along the code -- loop bodies stand out by their trip count.
   python tools/ncu_inst_executed.py <src.csv.gz> <warps launched> <iterations per launch> [seg]
(tools/exp_src.sh makes the CSV on the GPU box.)"""
import csv,gzip,re,collections,sys
rows=list(csv.reader(gzip.open(sys.argv[1],"rt")))
hdr=rows[1] if "Source" in rows[1] else rows[0]
ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows:
    try: n=int(r[ix["Instructions Executed"]])
    except Exception: continue
    data.append((r[ix["Address"]], r[ix["Source"]].strip(), n, int(r[ix["# Samples"]] or 0)))
tot=sum(d[2] for d in data); tots=sum(d[3] for d in data)
warps=int(sys.argv[2]); iters=int(sys.argv[3])
print("total warp-instr", tot, "per warp-iteration", tot/warps/iters, "samples", tots)
def cls(op):
    op=re.sub(r"^@!?U?P\d+\s+","",op).split()[0]
    if re.match(r"D(FMA|MUL|ADD|SETP|MNMX)",op): return "fp64"
    return "other"
c=collections.Counter(); cs=collections.Counter(); ops=collections.Counter()
for a,s,n,sm in data:
    c[cls(s)]+=n; cs[cls(s)]+=sm
    ops[re.sub(r"^@!?U?P\d+\s+","",s).split()[0].split(".")[0]]+=n
for k in c: print(k, c[k]/warps/iters, "per warp-iter; samples", cs[k])
print("top opcodes per warp-iteration:")
for k,v in ops.most_common(30): print("  %-10s %.1f"%(k, v/warps/iters))
# segments: print cumulative by address region of execution count plateaus
if len(sys.argv)>4:
    prev=None; start=None; acc=0; accs=0
    for a,s,n,sm in data:
        lvl=round(n/warps/iters,2)
        if prev is None or abs(lvl-prev)>0.05*max(prev,lvl,0.01):
            if prev is not None: print("  %s..%s exec/warp-iter=%.2f  n_instr=%d samples=%d"%(start,lasta,prev,cnt,accs))
            start=a; cnt=0; accs=0; prev=lvl
        cnt+=1; accs+=sm; lasta=a
    print("  %s..%s exec/warp-iter=%.2f  n_instr=%d samples=%d"%(start,lasta,prev,cnt,accs))
