"""Basic blocks of the first kernel of an `ncu --page source --csv --print-source cuda,sass` dump, with execution counts,
threads per instruction, stall-sample share, the source lines and the opcode mix of each — sorted by executed instructions
(or by samples with a third argument).  usage: ncu_blocks.py dump_cs.csv [top] [samp]"""
import csv, sys, collections
path=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
insts={}; cur_file=None; kern=0; seen=set(); line=None
for r in rows:
    if not r: continue
    if r[0]=="File Path":
        f=r[1].split("/")[-1]
        if f in seen and kern==1 and f=="rt_brute.cuh": kern=2
        if kern==0: kern=1
        seen.add(f); cur_file=f; continue
    if r[0] in("Function Name","Line No"): continue
    if kern!=1: continue
    if r[0].isdigit(): line=int(r[0]); continue
    if r[0]=="" and r[2].startswith("0x"):
        insts[int(r[2],16)]=(cur_file,line,r[3].strip(),int(r[7]),int(r[8]),int(r[6]))
addrs=sorted(insts)
# basic blocks: split at exec-count change or after a BRA/BSYNC/EXIT/CALL/RET
bbs=[];cur=[]
for a in addrs:
    f,l,s,e,t,sm=insts[a]
    if cur and (insts[cur[-1]][3]!=e):
        bbs.append(cur);cur=[]
    cur.append(a)
    op=s.split()[1] if s.startswith('@') else s.split()[0]
    if op.split('.')[0] in('BRA','EXIT','RET','CALL','BRX','JMP'):
        bbs.append(cur);cur=[]
if cur:bbs.append(cur)
tot=sum(i[3] for i in insts.values())
res=[]
for bb in bbs:
    e=insts[bb[0]][3]; n=len(bb)
    lines=collections.Counter((insts[a][0].replace('rt_','').replace('.cuh','').replace('.cu',''),insts[a][1]) for a in bb)
    ops=collections.Counter((insts[a][2].split()[1] if insts[a][2].startswith('@') else insts[a][2].split()[0]).split('.')[0] for a in bb)
    thr=sum(insts[a][4] for a in bb)/max(1,sum(insts[a][3] for a in bb))
    samp=sum(insts[a][5] for a in bb)
    res.append((e*n,e,n,thr,samp,lines,ops,bb[0]))
res.sort(key=lambda x:-(x[4] if len(sys.argv)>3 else x[0]))
S=sum(i[5] for i in insts.values())
cum=0
for w,e,n,thr,samp,lines,ops,a0 in res[:top]:
    cum+=w
    ls=", ".join(f"{f}:{l}x{c}" for (f,l),c in lines.most_common(5))
    os_=", ".join(f"{o}{c}" for o,c in ops.most_common(6))
    print(f"{100*w/tot:5.2f}% (cum {100*cum/tot:5.1f}) exec {e:8d} x {n:3d} instr thr {thr:4.1f} samp {100*samp/S:4.1f}% @{a0&0xfffff:05x} | {ls} | {os_}")
