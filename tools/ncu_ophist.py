"""Opcode histogram (per epilogue warp-iteration) of a conv_tc_kernel launch in an ncu report."""
import csv, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","-k","conv_tc_kernel","-c","1","-s",skip],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
print(rows[0][1][20:100])
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr) and r[0].startswith('0x')]
data=data[:len(data)//2]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iN=hdr.index('# Samples')
ld=[k for k,r in enumerate(data) if 'LDTM' in r[iS]]
it=int(data[ld[0]][iE])
print('epilogue warp-iterations (LDTM executions)',it, ' total inst', sum(int(r[iE]) for r in data))
c=collections.Counter(); cs=collections.Counter()
for r in data:
    e=int(r[iE])
    if e<it*0.2 or e>it*2.2: continue
    op=[t for t in r[iS].split() if not t.startswith('@')][0]
    op='.'.join(op.split('.')[:2]) if op.startswith(('IMAD','LDS','STS','SHFL')) else op.split('.')[0]
    c[op]+=e/it; cs[op]+=int(r[iN])
for op,n in c.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 30): print('%-12s %6.1f per warp-iter   samples %d'%(op,n,cs[op]))
print('total per warp-iter', round(sum(c.values()),1), ' samples', sum(cs.values()), 'of', sum(int(r[iN]) for r in data))
