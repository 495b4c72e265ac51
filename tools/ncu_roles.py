"""Summarise an ncu source-page CSV of conv_tc_kernel by warp role (uses UTCHMMA/UBLKCP/LDTM anchors)."""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","-k","conv_tc_kernel","-c","1","-s",skip],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
print(rows[0][:2])
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr) and r[0].startswith('0x')]
data=data[:len(data)//2] if len(data)>1 and data[0][0]==data[len(data)//2][0] else data
iS=hdr.index('Source'); iN=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') or 'Stall' in h]
idx=lambda pat:[k for k,r in enumerate(data) if pat in r[iS]]
mma=idx('UTCHMMA'); ldtm=idx('LDTM'); blk=idx('UBLKCP')
print('n instr',len(data),'first/last UTCHMMA',mma[0],mma[-1],'LDTM',ldtm[0],ldtm[-1],'UBLKCP',blk[0],blk[-1])
tot=sum(int(r[iN]) for r in data)
def region(a,b,name):
    sel=data[a:b]; s=sum(int(r[iN]) for r in sel); e=sum(int(r[iE]) for r in sel)
    print('%-10s instr[%d:%d] samples %6d (%.0f%%) executed %d'%(name,a,b,s,100*s/tot,e))
    for r in sorted(sel,key=lambda r:-int(r[iN]))[:int(sys.argv[3]) if len(sys.argv)>3 else 8]:
        print('      %6s %9s  %s'%(r[iN],r[iE],r[iS][:90]))
# heuristic boundaries: mma region = from ~60 instr before first UTCHMMA to ~25 after last; epilogue from there to ~first UBLKCP-40 (if producer after)
b0=max(0,mma[0]-80); b1=mma[-1]+30
if blk[0] > ldtm[-1]:
    region(0,b0,'prologue'); region(b0,b1,'mma'); region(b1,blk[0]-60,'epilogue'); region(blk[0]-60,len(data),'producer+tail')
else:
    region(0,blk[0]-60,'prologue'); region(blk[0]-60,b0,'producer'); region(b0,b1,'mma'); region(b1,len(data),'epilogue+tail')
