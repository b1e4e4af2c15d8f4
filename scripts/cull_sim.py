"""CPU simulation of a per-word bounding-box cull for k_tile_walk_bits (rejected).  A tile word = 32 consecutive
records of a cell in in-cell Morton order; the cull would skip a word whose box is farther than 2h from the box of a group of four own
particles.  Prints the fraction of words that survive: ~0.93 for groups of four, ~0.78 for single particles -- not worth a pass."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshless_inflatable_softbody_b200 import scenes, SceneConfig
cfg=SceneConfig(); h=cfg.h
x0,_=scenes.jittered_sphere(200000, seed=0)
x0=x0.astype(np.float32)
cw=2*h
c=np.floor(x0/cw).astype(np.int64); c-=c.min(0)
dims=c.max(0)+1
f=(x0/cw - np.floor(x0/cw)); s=np.minimum((f*8).astype(np.int64),7)
def morton(a):
    r=np.zeros(len(a),np.int64)
    for b in range(3):
        for ax in range(3):
            r|=((a[:,ax]>>b)&1)<<(3*b+ax)
    return r
lin=(c[:,2]*dims[1]+c[:,1])*dims[0]+c[:,0]
order=np.lexsort((morton(s),lin))
x=x0[order]; lin=lin[order]
starts=np.searchsorted(lin,np.arange(dims.prod())); ends=np.searchsorted(lin,np.arange(dims.prod()),side='right')
rng=np.random.default_rng(0)
cells=[cc for cc in rng.choice(np.unique(lin),40)]
tot_words=0; live_words=0; live_words_pp=0; groups=0
for cc in cells:
    cz=cc//(dims[0]*dims[1]); cy=(cc//dims[0])%dims[1]; cx=cc%dims[0]
    idx=[]
    for k in range(27):
        X,Y,Z=cx+k%3-1, cy+(k//3)%3-1, cz+k//9-1
        if 0<=X<dims[0] and 0<=Y<dims[1] and 0<=Z<dims[2]:
            l=(Z*dims[1]+Y)*dims[0]+X
            idx.append(np.arange(starts[l],ends[l]))
    idx=np.concatenate(idx); T=x[idx]
    Wc=(len(T)+31)//32
    lo=np.full((Wc,3),1e30); hi=np.full((Wc,3),-1e30)
    for w in range(Wc):
        seg=T[w*32:(w+1)*32]; lo[w]=seg.min(0); hi[w]=seg.max(0)
    own=x[starts[cc]:ends[cc]]
    lim=(2*h)**2*1.01
    for g in range(0,len(own),4):
        grp=own[g:g+4]; glo=grp.min(0); ghi=grp.max(0)
        gap=np.maximum(0,np.maximum(lo-ghi, glo-hi)); lb=(gap**2).sum(1)
        live=lb<lim
        tot_words+=Wc; live_words+=live.sum(); groups+=1
        for p in grp:
            gap=np.maximum(0,np.maximum(lo-p, p-hi)); live_words_pp+=((gap**2).sum(1)<lim).sum()/len(grp)
print('words per group',tot_words/groups,'live frac (group of 4)',live_words/tot_words,'live frac (single)',live_words_pp/tot_words)
