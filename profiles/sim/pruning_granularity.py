"""CPU simulation (numpy) of how many map points a P2P query has to visit under different exact pruning schemes, on Map-U at the
benchmark's density (1.25 M raw points in a 50 m box = 10 per voxel).  Run: python profiles/sim/pruning_granularity.py
Result of round 1 (3 000 queries):
    home voxel, then every voxel whose box is within the best distance (the kernel, phase A + B)   35.8 points, 2.06 column items
    home z-column first                                                                            46.6 points
    voxels best-first with re-pruning after each (sequential)                                      30.5 points, 3.4 voxels
    octants (0.5 m) best-first with re-pruning (sequential)                                         8.9 points, 7.8 octants
    home voxel, then every OCTANT whose box is within the best distance                            16.4 points, 4.5 non-empty octants
i.e. voxel-granular pruning is within 15 % of its limit; a sub-voxel (octant) layout could halve the candidates."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from elimaloc_b200 import synth
import elimaloc_b200 as E
rng=np.random.default_rng(0)
raw=synth.map_u(1_250_000,50.0)
m=E.VoxelHashMap(1.0,30,device=-1); m.AddPoints(raw); ex=m.export()
keys=ex['keys']; counts=ex['counts']; pts=ex['pxyz'].astype(np.float64); starts=np.concatenate([[0],np.cumsum(counts)])
vox={tuple(k):i for i,k in enumerate(keys)}
Q=rng.random((3000,3))*30+10
def box_lb(q,kk):  # squared distance from q to voxel box [kk,kk+1)
    lo=np.array(kk,float); hi=lo+1
    g=np.maximum(np.maximum(lo-q,q-hi),0); return (g*g).sum()
res={'hv':[], 'col':[], 'bestfirst':[], 'hv_items':[], 'bf_steps':[]}
for q in Q:
    k=np.floor(q).astype(int)
    def pts_of(kk):
        v=vox.get(tuple(kk)); 
        return pts[starts[v]:starts[v+1]] if v is not None else np.zeros((0,3))
    nbrs=[(dx,dy,dz) for dx in(-1,0,1) for dy in(-1,0,1) for dz in(-1,0,1)]
    # hv: home voxel, then all voxels with lb <= best
    P=pts_of(k); best=((P-q)**2).sum(1).min() if len(P) else np.inf
    vis=len(P); items=set()
    for d in nbrs:
        if d==(0,0,0): continue
        kk=k+np.array(d)
        if box_lb(q,kk)<=best:
            vis+=len(pts_of(kk)); items.add((d[0],d[1]))
    res['hv'].append(vis); res['hv_items'].append(len(items))
    # col: home column first
    vis=0; best=np.inf
    for dz in(-1,0,1):
        P=pts_of(k+np.array([0,0,dz])); vis+=len(P)
        if len(P): best=min(best,((P-q)**2).sum(1).min())
    for d in nbrs:
        if d[0]==0 and d[1]==0: continue
        kk=k+np.array(d)
        if box_lb(q,kk)<=best: vis+=len(pts_of(kk))
    res['col'].append(vis)
    # best-first with re-pruning
    order=sorted(nbrs,key=lambda d:box_lb(q,k+np.array(d)))
    best=np.inf; vis=0; steps=0
    for d in order:
        kk=k+np.array(d)
        if box_lb(q,kk)>best: break
        P=pts_of(kk); vis+=len(P); steps+=1
        if len(P): best=min(best,((P-q)**2).sum(1).min())
    res['bestfirst'].append(vis); res['bf_steps'].append(steps)
for k,v in res.items(): print(k, np.mean(v))

# ---- octant granularity (0.5 m sub-cells inside the 1 m voxels), best-first and "all sub-cells with lb <= best after home sub-cell"
sub={}
for i,p in enumerate(pts):
    sub.setdefault(tuple(np.floor(p*2).astype(int)),[]).append(i)
sub={k:np.array(v) for k,v in sub.items()}
r_bf=[]; r_steps=[]; r_hv=[]; r_hv_cells=[]
for q in Q:
    k=np.floor(q).astype(int)
    cells=[]
    for dx in range(-2,4):
        for dy in range(-2,4):
            for dz in range(-2,4):
                kk=2*k+np.array([dx,dy,dz])   # all octants of the 27 voxels
                lo=kk/2.0; hi=lo+0.5
                g=np.maximum(np.maximum(lo-q,q-hi),0); cells.append(((g*g).sum(),tuple(kk)))
    cells.sort()
    best=np.inf; vis=0; steps=0
    for lb,kk in cells:
        if lb>best: break
        idx=sub.get(kk); steps+=1
        if idx is None: continue
        P=pts[idx]; vis+=len(P); best=min(best,((P-q)**2).sum(1).min())
    r_bf.append(vis); r_steps.append(steps)
    # two-phase like the kernel: home octant, then every octant with lb <= best
    lb0,k0=cells[0]; idx=sub.get(k0); best=np.inf; vis=0; n=0
    if idx is not None: P=pts[idx]; vis=len(P); best=((P-q)**2).sum(1).min()
    for lb,kk in cells[1:]:
        if lb<=best:
            n+=1; idx=sub.get(kk)
            if idx is not None: vis+=len(idx)
    r_hv.append(vis); r_hv_cells.append(n)
print('octant best-first visited',np.mean(r_bf),'cells',np.mean(r_steps)); print('octant two-phase visited',np.mean(r_hv),'extra cells',np.mean(r_hv_cells))

# ---- hybrid: phase A = home VOXEL (as the kernel does), phase B = every OCTANT outside it whose box is within the best distance
r_h=[]; r_hc=[]; r_hnonempty=[]
for q in Q:
    k=np.floor(q).astype(int)
    v=vox.get(tuple(k)); P=pts[starts[v]:starts[v+1]] if v is not None else np.zeros((0,3))
    best=((P-q)**2).sum(1).min() if len(P) else np.inf; vis=len(P); n=0; ne=0
    for dx in range(-2,4):
        for dy in range(-2,4):
            for dz in range(-2,4):
                kk=2*k+np.array([dx,dy,dz])
                if (kk//2==k).all(): continue
                lo=kk/2.0; hi=lo+0.5
                g=np.maximum(np.maximum(lo-q,q-hi),0)
                if (g*g).sum()<=best:
                    n+=1; idx=sub.get(tuple(kk))
                    if idx is not None: vis+=len(idx); ne+=1
    r_h.append(vis); r_hc.append(n); r_hnonempty.append(ne)
print('hybrid (home voxel + octants) visited',np.mean(r_h),'octants tested',np.mean(r_hc),'non-empty',np.mean(r_hnonempty))
