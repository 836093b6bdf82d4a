"""Python model of the OSD scheme of csrc/osd.cuh (only the 83 parity columns stored, systematic columns implicit, in-place
images, optional late-row pivot preference), checked trial word by trial word against the oracle's osd_candidates, with
the work statistics per call the kernel design was based on.  Test infrastructure (imports oracle/).

  python tools/osd_model.py [late_threshold|none] [n_cases]      e.g.  python tools/osd_model.py 96 200
"""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, ROOT)
import ft8_oracle as o
G0=o.G0
PC=[sum(int(G0[r,91+j])<<r for r in range(91)) for j in range(83)]   # parity column j as a 91-bit int over rows
stats=dict(calls=0,visits=0,sys_free=0,par_piv=0,par_dep=0,tsys_piv=0,tsys_dep=0)
def model(llr,S=30,D=2,late=None):
    order=o.osd_order(llr); hard=(llr>0).astype(int)
    col=list(PC); used=0; u=0; owner={}; ownrow=[None]*83; piv=[]   # piv: list of (row, orig col)
    pos={int(c):i for i,c in enumerate(order)}
    L=0
    if late is not None:
        for c in range(91):
            if pos[c]>=late: L|=1<<c
    stats['calls']+=1
    for sp in range(174):
        c=int(order[sp]); hb=int(hard[c]); stats['visits']+=1
        if c<91:
            if not (used>>c)&1:
                used|=1<<c; u|=hb<<c; piv.append((c,c)); stats['sys_free']+=1
                if len(piv)==91: break
                continue
            j=owner[c]
        else: j=c-91
        v=col[j]; f=v&~used
        if f==0:
            stats['par_dep' if c>=91 else 'tsys_dep']+=1; continue
        g=f&L if (f&L) else f
        p=(g&-g).bit_length()-1
        stats['par_piv' if c>=91 else 'tsys_piv']+=1
        used|=1<<p; u|=hb<<p; piv.append((p,c))
        if c<91: ownrow[j]=None if False else ownrow[j]
        owner[p]=j; ownrow[j]=p
        m=v&~(1<<p)
        for k in range(83):
            if k!=j and (col[k]>>p)&1: col[k]^=m
        if len(piv)==91: break
    assert len(piv)==91
    basis_sys={c for (_,c) in piv if c<91}
    # base word over systematic positions 0..90
    def word(uvec,extra=0):
        w=0
        for c in range(91):
            if c in basis_sys: continue
            j=owner[c]; assert ownrow[j]==c
            if bin(col[j]&uvec).count('1')&1: w|=1<<c
        for (row,c) in piv:
            if c<91 and (uvec>>row)&1: w|=1<<c
        return w
    out=[word(u)]
    fl=[row for (row,_) in piv[::-1][:S]]
    for i in range(S): out.append(word(u^(1<<fl[i])))
    for i in range(S):
        for j2 in range(D):
            if j2<i: out.append(word(u^(1<<fl[i])^(1<<fl[j2])))
    return out
def to_int(w):  # bit c of w -> oracle's bits_to_int of array index c
    return o.bits_to_int(np.array([(w>>c)&1 for c in range(91)],np.uint8))
rng=np.random.default_rng(5)
late=int(sys.argv[1]) if len(sys.argv)>1 and sys.argv[1]!='none' else None
n=int(sys.argv[2]) if len(sys.argv)>2 else 60
for t in range(n):
    llr=(rng.normal(0,2.83,174)).astype(np.float32)
    if t%3==1: llr=o.set_ap(llr,1+t%4)
    if t%7==0: llr[rng.integers(0,174,20)]=llr[3]   # ties
    a=model(llr,late=late); b=o.osd_candidates(llr)
    assert [to_int(x) for x in a]==list(b), t
print('ok',{k:v/stats['calls'] for k,v in stats.items()})
