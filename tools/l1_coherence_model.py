"""Model of the L1 data-pipe cost of the pair kernels gathers (no GPU): for a thermal fcc Cu crystal in the library s cell order, the
number of distinct 128-byte lines / 32-byte sectors touched when the 32 lanes of a warp (32 consecutive atoms) load the partner
record of their p-th list entry, under different orders of atoms inside a cell and of entries inside a row.  Result (round 2): 19.6
lines per request for the order in use; none of the alternatives goes below 16.5 even in a perfect crystal -> the gathers cost ~20
L1 wavefronts per 32 pairs whatever the order (ncu measures 24).  python tools/l1_coherence_model.py"""
import numpy as np, sys
from scipy.spatial import cKDTree
sys.path.insert(0,'/root/repo')
from pfmds_b200 import inputs
case = inputs.cu_fcc(ncell=24, seed=2)   # 55296 atoms
pos = case["pos"].copy(); L = np.array(case["box"]); N=len(pos)
rng=np.random.default_rng(1); pos=(pos+rng.normal(0,0.08,pos.shape))%L
rc=6.5; R1=5.5; R2=6.0
ncell=np.floor(L/(rc*(1+1e-9)+1e-5)).astype(int); print("ncell",ncell, "atoms/cell", N/np.prod(ncell))
ci=np.minimum((pos/L*ncell).astype(int),ncell-1)
cid=(ci[:,2]*ncell[1]+ci[:,1])*ncell[0]+ci[:,0]
tree=cKDTree(pos,boxsize=L)
pairs=tree.query_pairs(rc,output_type='ndarray')
ii=np.concatenate([pairs[:,0],pairs[:,1]]); jj=np.concatenate([pairs[:,1],pairs[:,0]])
d=pos[jj]-pos[ii]; d-=L*np.round(d/L); r=np.linalg.norm(d,axis=1)
cls=np.where(r<R1,0,np.where(r<R2,1,2))
def morton(sub,bits):
    k=np.zeros(len(sub),dtype=np.int64)
    for b in range(bits):
        for a in range(3):
            k|=((sub[:,a]>>b)&1)<<(3*b+a)
    return k
def mem_order(kind):
    frac=(pos/L*ncell)-ci   # in-cell fractional
    if kind=="file": key2=np.arange(N)
    elif kind.startswith("morton"):
        bits=int(kind[6:]); sub=np.minimum((frac*(1<<bits)).astype(int),(1<<bits)-1); key2=morton(sub,bits)*N+np.arange(N)
    elif kind=="zyx":
        sub=np.minimum((frac*4).astype(int),3); key2=((sub[:,2]*4+sub[:,1])*4+sub[:,0])*N+np.arange(N)
    order=np.lexsort((key2,cid))
    slot=np.empty(N,dtype=np.int64); slot[order]=np.arange(N)
    return slot
def lines(slot,rowkey,label):
    si=slot[ii]; sj=slot[jj]
    if rowkey=="j": k=sj
    elif rowkey=="disp":
        q=np.round(d/0.9).astype(np.int64)+16   # quantised displacement
        k=((q[:,2]*64+q[:,1])*64+q[:,0])
    elif rowkey=="disp_fine":
        q=np.round(d/0.45).astype(np.int64)+32
        k=((q[:,2]*128+q[:,1])*128+q[:,0])
    o=np.lexsort((sj,k,cls,si))
    si_s=si[o]; sj_s=sj[o]
    # slot index p within row
    start=np.searchsorted(si_s,np.arange(N)); p=np.arange(len(si_s))-start[si_s]
    warp=si_s//32
    key=warp*128+p
    # distinct lines per (warp,p)
    line=sj_s//4
    u=np.unique(np.stack([key,line],1),axis=0)
    nreq=len(np.unique(key))
    sect=np.unique(np.stack([key,sj_s],1),axis=0)
    print("%-28s lines/request %.2f  sectors/request %.2f  lanes/request %.2f"%(label,len(u)/nreq,len(sect)/nreq,len(key)/nreq))
for mk in ["file","zyx","morton1","morton2","morton3"]:
    s=mem_order(mk)
    for rk in ["j","disp","disp_fine"]:
        lines(s,rk,mk+"/"+rk)
