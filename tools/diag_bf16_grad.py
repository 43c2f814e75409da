import sys; sys.path[:0]=['.','./oracle','./tests']
import numpy as np, torch
import groove_oracle as G
from _util import build_model, grads_by_name
cfg=G.GrooveCfg(32,16,512,6,0,16,27); pen=0.38; p=0.24
for n in (4,64):
    x,y=G.det_batch(cfg,n)
    loss6,grads,_=G.train_step_oracle(G.det_params(cfg),cfg,x,y,pen,G.DropCtx(p,7,1,0,True))
    for prec in ('fp32','bf16'):
        m,P=build_model(cfg,dropout=p,precision=prec); m.set_seed(7,1,0).train()
        met,_=m.train_step(x.cuda(),y.cuda(),pen)
        gg=grads_by_name(m)
        errs=[]
        for k,w in grads.items():
            sc=float(w.abs().max())
            if sc<1e-6: continue
            errs.append((float((gg[k]-w).abs().max())/sc, float((gg[k]-w).norm()/w.norm()), k))
        errs.sort(reverse=True)
        print(n,prec,'loss',float(met[0]),loss6[0],'worst',[(round(a,4),round(b,4),k[-30:]) for a,b,k in errs[:4]], 'median', np.median([e[0] for e in errs]))
