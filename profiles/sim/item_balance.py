"""CPU simulation (numpy) of the work-item balance of phase B of the P2P search on Map-U at the benchmark's density.
Run: python profiles/sim/item_balance.py
Round-1 result (8 192 queries = 256 warps): 2.04 column items per query, 65.3 +- 8.4 per warp (half of the warps need a third
round); the items of a warp hold 850 points: 26.6 per lane if perfectly balanced, but one item per lane and round costs the
longest run of every round: 60.9 point steps, i.e. 44 % lane efficiency.
In warp steps of one 6-point batch per busy lane: static rounds (the kernel) 11.1, greedy pull from a warp counter 7.7, lengths
sorted + zig-zag assignment 6.8, lower bound 5.3."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from elimaloc_b200 import synth
import elimaloc_b200 as E
rng=np.random.default_rng(0)
raw=synth.map_u(1_250_000,50.0)
m=E.VoxelHashMap(1.0,30,device=-1); m.AddPoints(raw); ex=m.export()
keys=ex['keys']; counts=ex['counts']; pts=ex['pxyz'].astype(np.float64); starts=np.concatenate([[0],np.cumsum(counts)])
vox={tuple(k):i for i,k in enumerate(keys)}
Q=rng.random((8192,3))*30+10
items=[]; lens=[]
for q in Q:
    k=np.floor(q).astype(int)
    v=vox.get(tuple(k)); P=pts[starts[v]:starts[v+1]] if v is not None else np.zeros((0,3))
    best=((P-q)**2).sum(1).min() if len(P) else np.inf
    cols={}
    for dx in(-1,0,1):
        for dy in(-1,0,1):
            for dz in(-1,0,1):
                if (dx,dy,dz)==(0,0,0): continue
                kk=k+np.array([dx,dy,dz]); lo=kk.astype(float); hi=lo+1
                g=np.maximum(np.maximum(lo-q,q-hi),0)
                if (g*g).sum()<=best:
                    vv=vox.get(tuple(kk)); cols[(dx,dy)]=cols.get((dx,dy),0)+(counts[vv] if vv is not None else 0)
    items.append(len(cols)); lens.append(list(cols.values()))
items=np.array(items)
w=items.reshape(-1,32).sum(1)
print('items/query mean',items.mean(),'per warp mean',w.mean(),'std',w.std(),'rounds hist',np.bincount(np.ceil(w/32).astype(int)))
# lane time model: each round costs the longest run of the round (lanes take items in list order)
tot_pts=[]; lane_max=[]
for wi in range(len(w)):
    L=[l for q in range(wi*32,wi*32+32) for l in lens[q]]
    rounds=[L[i:i+32] for i in range(0,len(L),32)]
    lane_max.append(sum(max(r) for r in rounds)); tot_pts.append(sum(L))
print('per warp: points in items',np.mean(tot_pts),' sum over rounds of the longest run',np.mean(lane_max),' ideal (points/32)',np.mean(tot_pts)/32)

# ---- batch-granular comparison (a warp step = every busy lane consumes one batch of 6 points = three 32-byte loads)
B = 6
static_steps, dynamic_steps, sorted_steps = [], [], []
for wi in range(len(w)):
    L = [l for q in range(wi * 32, wi * 32 + 32) for l in lens[q]]
    nb = [max(1, -(-l // B)) for l in L]                       # batches per item (an empty run still costs its descriptor step)
    rounds = [nb[i:i + 32] for i in range(0, len(nb), 32)]
    static_steps.append(sum(max(r) for r in rounds))            # the kernel today: one item per lane and round
    lanes = [0] * 32                                           # greedy list scheduling: a lane that runs out pulls the next item
    for n in nb:
        i = lanes.index(min(lanes))
        lanes[i] += n
    dynamic_steps.append(max(lanes))
    s = sorted(nb, reverse=True)                               # sorted + zig-zag: lane i gets items i, 63 - i, 64 + i, ...
    lanes = [0] * 32
    for j, n in enumerate(s):
        r, k = divmod(j, 32)
        lanes[k if r % 2 == 0 else 31 - k] += n
    sorted_steps.append(max(lanes))
print('warp steps per tile (batches of 6 points): static rounds', np.mean(static_steps), ' greedy pull', np.mean(dynamic_steps),
      ' sorted zig-zag', np.mean(sorted_steps), ' lower bound', np.mean([sum(max(1, -(-l // B)) for q in range(wi*32, wi*32+32) for l in lens[q]) / 32 for wi in range(len(w))]))
