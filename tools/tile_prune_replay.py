"""Offline replay of the sweeps' tile-level test on a dump of a real fit (debug knob ANNB_DUMP_TILES=prefix writes
prefix.<k>.bin at every threshold pair sweep: per-tile anchor-distance intervals, closest-anchor masks, the cut of
every row, the regression model, which tiles hold store entries).  Prints how many tile pairs survive the test the
kernels apply, and what-if numbers for other cut choices -- this is what motivated the reduced tile mode
(DESIGN.md section 3).      python tools/tile_prune_replay.py prefix.0.bin"""
import numpy as np, sys
def load(path):
    f=open(path,'rb')
    hdr=np.frombuffer(f.read(48),np.int64); T,KA,na,npad,has2,msz=[int(x) for x in hdr]
    m=f.read(msz)
    MAXB=8
    nb=np.frombuffer(m[:4],np.int32)[0]
    off=4
    edge=np.frombuffer(m[off:off+4*(MAXB+1)],np.float32); off+=4*(MAXB+1)
    e2=np.frombuffer(m[off:off+4*MAXB],np.float32); off+=4*MAXB
    c0=np.frombuffer(m[off:off+4*MAXB],np.float32); off+=4*MAXB
    c1=np.frombuffer(m[off:off+4*MAXB],np.float32); off+=4*MAXB
    c2=np.frombuffer(m[off:off+4*MAXB],np.float32); off+=4*MAXB
    ic=np.frombuffer(m[off:off+4*MAXB],np.float32); off+=4*MAXB
    lo=np.frombuffer(f.read(4*T*KA),np.float32).reshape(T,KA)[:,:na]
    hi=np.frombuffer(f.read(4*T*KA),np.float32).reshape(T,KA)[:,:na]
    cm=np.frombuffer(f.read(8*T),np.uint64)
    cut1=np.frombuffer(f.read(4*npad),np.float32)
    cut2=np.frombuffer(f.read(4*npad),np.float32)
    NT=T*(T+1)//2
    he=np.frombuffer(f.read(NT),np.uint8)
    return dict(T=T,na=na,nb=int(nb),edge=edge,e2=e2,c0=c0,c1=c1,c2=c2,ic=ic,lo=lo,hi=hi,cm=cm,cut1=cut1,cut2=cut2,he=he,has2=has2)
d=load(sys.argv[1])
T,na,nb=d['T'],d['na'],d['nb']
print('T',T,'na',na,'nb',nb,'has2',d['has2'])
print('edges e2',d['e2'][:nb],'c0',d['c0'][:nb],'c1',d['c1'][:nb],'c2',d['c2'][:nb],'ic',d['ic'][:nb])
c1t=d['cut1'][:T*128].reshape(T,128); c2t=d['cut2'][:T*128].reshape(T,128)
print('cut1 pct',np.percentile(d['cut1'][np.isfinite(d['cut1'])],[1,50,99,100]),'cut2 pct',np.percentile(d['cut2'][np.isfinite(d['cut2'])],[1,50,99,100]))
tmax=np.maximum(c1t.max(1),c2t.max(1)); tmed=np.maximum(np.median(c1t,1),np.median(c2t,1))
print('tile cutmax pct',np.percentile(tmax[np.isfinite(tmax)],[1,50,99]),'tile cut median pct',np.percentile(tmed,[1,50,99]), 'ninf', (~np.isfinite(tmax)).sum())
cmbits=((d['cm'][:,None]>>np.arange(na,dtype=np.uint64)[None,:])&np.uint64(1)).astype(bool)
lo,hi=d['lo'],d['hi']
print('mean width',(hi-lo).mean(),'width at own cA (min width)',(hi-lo).min(1).mean(), 'anchors per tile', cmbits.sum(1).mean())
rng=np.random.default_rng(0)
rows=rng.choice(T,200,replace=False)
tot=0;surv=0;surv_he=0;surv_med=0; gaps=[]
offs=np.concatenate([[0],np.cumsum(T-np.arange(T))])
for ti in rows:
    tj=np.arange(ti+1,T)
    li=lo[ti][None,:];hi_=hi[ti][None,:];lj=lo[tj];hj=hi[tj]
    lbmin=np.maximum(np.maximum(li-hj,lj-hi_).max(1),0);lbmax=np.maximum(hi_-lj,hj-li).max(1)
    ubmin=(li+lj).min(1);ubmax=(hi_+hj).min(1)
    cmj=cmbits[tj]
    siLo=np.where(cmj,li,np.inf).min(1);siHi=np.where(cmj,hi_,-np.inf).max(1)
    cmi=cmbits[ti][None,:]
    sjLo=np.where(cmi,lj,np.inf).min(1);sjHi=np.where(cmi,hj,-np.inf).max(1)
    slo=siLo+sjLo;shi=siHi+sjHi
    pmin=np.full(len(tj),np.inf)
    for b in range(nb):
        blo=-np.inf if b==0 else d['e2'][b]; bhi=np.inf if b+1>=nb else d['e2'][b+1]
        ov=(shi>blo)&(slo<=bhi)
        c0,c1,cz,ic=d['c0'][b],d['c1'][b],0.5*d['c2'][b],d['ic'][b]
        y=(lbmin if c0>=0 else lbmax)*c0+(ubmin if c1>=0 else ubmax)*c1+(slo if cz>=0 else shi)*cz+ic
        p=np.minimum(np.maximum(y,lbmin),ubmin)
        pmin=np.where(ov,np.minimum(pmin,p),pmin)
    cutmax=np.maximum(tmax[ti],tmax[tj]); cutmed=np.maximum(tmed[ti],tmed[tj])
    he=d['he'][offs[ti]+1:offs[ti]+1+len(tj)].astype(bool)
    s=~(pmin>cutmax)
    tot+=len(tj); surv+=(s|he).sum(); surv_he+=he.sum(); surv_med+=(~(pmin>cutmed)|he).sum()
    gaps.append((pmin-cutmax)[~he])
gaps=np.concatenate(gaps)
print('tile pairs',tot,'surviving frac %.3f'%(surv/tot),'with entries %.3f'%(surv_he/tot),'surviving if median cut %.3f'%(surv_med/tot))
print('pmin-cutmax percentiles',np.percentile(gaps[np.isfinite(gaps)],[1,5,25,50,75,95]))

# ---- what-if analysis ----
def whatif(q):
    tq=np.maximum(np.percentile(c1t,q,axis=1),np.percentile(c2t,q,axis=1)) if q<100 else tmax
    tot=0; a=0; b=0; c=0
    for ti in rows[:100]:
        tj=np.arange(ti+1,T)
        li=lo[ti][None,:];hi_=hi[ti][None,:];lj=lo[tj];hj=hi[tj]
        lbmin=np.maximum(np.maximum(li-hj,lj-hi_).max(1),0);lbmax=np.maximum(hi_-lj,hj-li).max(1)
        ubmin=(li+lj).min(1);ubmax=(hi_+hj).min(1)
        cmj=cmbits[tj]
        siLo=np.where(cmj,li,np.inf).min(1);siHi=np.where(cmj,hi_,-np.inf).max(1)
        cmi=cmbits[ti][None,:]
        sjLo=np.where(cmi,lj,np.inf).min(1);sjHi=np.where(cmi,hj,-np.inf).max(1)
        slo=siLo+sjLo;shi=siHi+sjHi
        pmin=np.full(len(tj),np.inf)
        for bb in range(nb):
            blo=-np.inf if bb==0 else d['e2'][bb]; bhi=np.inf if bb+1>=nb else d['e2'][bb+1]
            ov=(shi>blo)&(slo<=bhi)
            c0,c1,cz,ic=d['c0'][bb],d['c1'][bb],0.5*d['c2'][bb],d['ic'][bb]
            y=(lbmin if c0>=0 else lbmax)*c0+(ubmin if c1>=0 else ubmax)*c1+(slo if cz>=0 else shi)*cz+ic
            p=np.minimum(np.maximum(y,lbmin),ubmin)
            pmin=np.where(ov,np.minimum(pmin,p),pmin)
        cut=np.maximum(tq[ti],tq[tj])
        he=d['he'][offs[ti]+1:offs[ti]+1+len(tj)].astype(bool)
        s=~(pmin>cut)
        tot+=len(tj); a+=(s|he).sum(); b+=s.sum(); c+=(he&~s).sum()
    return a/tot,b/tot,c/tot
for q in (100,97,90,50):
    a,b,c=whatif(q)
    print('cut percentile %3d: surviving now-rule %.3f | if entry tiles prunable (flag-only pass) %.3f | entry tiles otherwise prunable %.3f'%(q,a,b,c))
