"""CPU model of the lane efficiency of K1 (development aid): which quadrature rule every (tile lane, element) pair of the
S-cube gets (distance thresholds of the rule estimator), lane-masked execution vs in-place set 0 + perfectly packed others.
Reproduces the 0.68 lane efficiency ncu measured for the lane-masked kernel and predicts 0.95 for the deferred compaction."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from multifebe_b200.host import Model, cube_mesh, cube_bcs, shape
m=40
md=Model(cube_mesh(m, shape.TRI3), cube_bcs())
X=md.node_x; conn=np.array(md.mesh.conn)
ctr=X[conn].mean(axis=1)
R=np.linalg.norm(X[conn]-ctr[:,None,:],axis=2).max(axis=1)
cl=np.sqrt(2)/m
def morton(P):
    q=np.clip(((P-P.min(0))/(P.max(0)-P.min(0)).max()*1023+0.5).astype(np.int64),0,1023)
    key=np.zeros(len(P),dtype=np.int64)
    for b in range(9,-1,-1):
        for c in range(3):
            key=(key<<1)|((q[:,c]>>b)&1)
    return key
# row nodes with multiplicity
mult=np.bincount(md.colloc_node,minlength=md.n_node)
cx=md.colloc_x; cn=md.colloc_node
def tiles_for(order_key):
    tiles=[]
    for mu in sorted(set(mult[mult>0])):
        nodes=np.where(mult==mu)[0]
        nodes=nodes[np.argsort(order_key[nodes],kind='stable')]
        for i in range(0,len(nodes),32):
            blk=nodes[i:i+32]
            for layer in range(mu):
                pts=[np.where(cn==nd)[0][layer] for nd in blk]
                tiles.append(np.array(pts))
    return tiles
ngp_of={2:4,3:7,4:15,5:19,6:28,7:40,8:54,9:66}
def rule(d):
    r=np.full(d.shape,7)   # near & regular guess
    r[d>2.0]=5; r[d>2.23]=4; r[d>4.28]=3; r[d>10.85]=2
    r[d<0.55]=0  # adaptive/singular: excluded
    return r
def evaluate(tiles, eorder):
    ideal=0; execd=0; per_rule_exec={}
    for t in tiles:
        P=cx[t]
        dist=np.linalg.norm(P[:,None,:]-ctr[None,:,:],axis=2)-R[None,:]
        d=dist/cl
        rl=rule(d)   # [lanes, elements]
        for g,n in ngp_of.items():
            msk=(rl==g)
            ideal+=msk.sum()*n
            anyl=msk.any(axis=0)
            execd+=anyl.sum()*32*n
            per_rule_exec[g]=per_rule_exec.get(g,0)+anyl.sum()*32*n
    return ideal, execd, per_rule_exec
key=morton(X)
tl=tiles_for(key)
i,e,pr=evaluate(tl,None)
print("morton tiles:",len(tl),"ideal pts",i,"executed lane-pts",e,"eff",i/e, {k:v/e for k,v in pr.items()})
# row-order (node index) tiles
tl2=tiles_for(np.arange(md.n_node))
i,e,pr=evaluate(tl2,None)
print("index-order tiles:",len(tl2),"eff",i/e, {k:round(v/e,3) for k,v in pr.items()})
def evaluate2(tiles):
    ideal2=0; exec2=0; ideal_o=0; pairs_o=0; pairs2=0
    for t in tiles:
        P=cx[t]
        d=(np.linalg.norm(P[:,None,:]-ctr[None,:,:],axis=2)-R[None,:])/cl
        rl=rule(d)
        msk=(rl==2); ideal2+=msk.sum()*4; exec2+=msk.any(axis=0).sum()*32*4; pairs2+=msk.sum()
        for g,n in ngp_of.items():
            if g!=2: ideal_o+=(rl==g).sum()*n; pairs_o+=(rl==g).sum()
    return ideal2,exec2,ideal_o,pairs2,pairs_o
i2,e2,io,p2,po=evaluate2(tl)
print("rule2 in-place eff",i2/e2,"ideal2",i2,"exec2",e2,"others ideal",io,"pairs2",p2,"pairs other",po)
print("total eff with perfect packing of others",(i2+io)/(e2+io))
