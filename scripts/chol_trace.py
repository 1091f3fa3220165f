"""Summarise the per-tile time stamps of the distributed Cholesky (library built with VIPE_BA_DEFINES=VBA_CHOL_TRACE, run with
VIPE_BA_CHOL_TRACE=gpurun_out/choltr): per-rank timelines of diagonal / sub-diagonal tiles, the column period and where it goes.
Stamps per tile (us, relative to the rank's first claim): 0 claimed, 1 input tile loaded, 7 last update input in shared memory,
2 update loop done, 3 tile solve starts, 4 potrf / tile solve done, 5 result stored, 6 flag published.
Usage: python scripts/chol_trace.py [prefix]"""
import sys
PREFIX = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/choltr"
import numpy as np
def load(r):
    L=open(f'{PREFIX}_r{r}.txt').read().split('\n')
    T,rank,world,cb=map(int,L[0].split())
    rows={}
    for l in L[1:]:
        if not l.strip(): continue
        v=list(map(int,l.split())); rows[v[0]]=v[1:]
    return T,rank,world,rows
for R in (0,1):
    T,rank,world,rows=load(R)
    def tile_of(t):
        j=0
        while (j+1)*(T+1)-(j+1)*j//2 <= t: j+=1
        return j+(t-(j*(T+1)-j*(j-1)//2)),j
    t0=min(v[0] for v in rows.values() if v[0])
    diag={}; sub={}; oth={}
    allend=0
    for t,v in rows.items():
        i,j=tile_of(t); allend=max(allend,max(v))
        if i==j: diag[j]=v
        elif i==j+1: sub[j]=v
        if i<T and i!=j: oth[(i,j)]=v
    print('rank',rank,'tiles',len(rows),'span us',(allend-t0)/1e3)
    js=sorted(diag)
    for j in js[:3]+js[20:23]+js[-3:]:
        v=diag[j]; print(' diag',j,[round((x-t0)/1e3,1) if x else None for x in [v[0],v[1],v[7],v[2],v[4],v[5],v[6]]])
    for j in sorted(sub)[20:23]:
        v=sub[j]; print(' sub ',j,[round((x-t0)/1e3,1) if x else None for x in [v[0],v[7],v[2],v[3],v[4],v[5],v[6]]])
    per=[(diag[js[k+1]][6]-diag[js[k]][6])/1e3/(js[k+1]-js[k]) for k in range(len(js)-1)]
    print(' per-column period us: mean',round(np.mean(per),1),'early',np.round(per[:4],1),'mid',np.round(per[20:24],1),'late',np.round(per[-4:],1))
    print(' diag: lastinput->kloopend',round(np.mean([(diag[j][2]-diag[j][7])/1e3 for j in js if diag[j][7]]),1),'potrf',round(np.mean([(diag[j][4]-diag[j][2])/1e3 for j in js]),1),'store',round(np.mean([(diag[j][5]-diag[j][4])/1e3 for j in js]),1),'publish',round(np.mean([(diag[j][6]-diag[j][5])/1e3 for j in js]),1))
    ss=sorted(sub)
    print(' sub: lastinput->kloopend',round(np.mean([(sub[j][2]-sub[j][7])/1e3 for j in ss if sub[j][7]]),1),'wait Ljj (3->4 incl trsm)',round(np.mean([(sub[j][4]-sub[j][3])/1e3 for j in ss]),1),'store',round(np.mean([(sub[j][5]-sub[j][4])/1e3 for j in ss]),1),'publish',round(np.mean([(sub[j][6]-sub[j][5])/1e3 for j in ss]),1))
    # generic tiles: time from claim to publish, busy fraction
    dur=[(v[6]-v[0])/1e3 for v in oth.values() if v[6]]
    wait=[(v[7]-v[1])/1e3 for v in oth.values() if v[7] and v[1]]
    print(' off-diag tiles: n',len(dur),'mean claim->publish',round(np.mean(dur),1),'mean (A loaded -> last input ready)',round(np.mean(wait),1))
print()
for R,tiles in ((0,[(40,40),(41,40),(42,40),(43,40),(60,40),(93,40)]),(1,[(41,41),(42,41),(43,41),(60,41)])):
    T,rank,world,rows=load(R)
    def start(j): return j*(T+1)-j*(j-1)//2
    t0=min(v[0] for v in rows.values() if v[0])
    for (i,j) in tiles:
        v=rows.get(start(j)+(i-j))
        print('rank',R,'tile',(i,j),'claim,Aload,lastin(7),kend(2),trsm0(3),trsm1(4),stored(5),pub(6):',[round((x-t0)/1e3,1) if x else None for x in [v[0],v[1],v[7],v[2],v[3],v[4],v[5],v[6]]])
