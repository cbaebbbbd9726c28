#!/usr/bin/env python3
"""Design-exploration helper (not product, not test): decode a raw DEFLATE stream into tokens with
bit positions, to study sub-chunk statistics (bytes per sub-chunk, match lengths, distances)."""
import sys, zlib
import numpy as np
sys.path.insert(0, "/root/repo")

LB=[3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258]
LE=[0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0]
DB=[1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577]
DE=[0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13]
ZZ=[16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15]

class BR:
    def __init__(s,b): s.v=int.from_bytes(b,'little'); s.p=0
    def peek(s,n): return (s.v>>s.p)&((1<<n)-1)
    def get(s,n): r=s.peek(n); s.p+=n; return r

def mk(lens):
    # canonical codes -> dict (len, code_lsb_first) -> sym ; decode by trying lengths
    maxl=max(lens) if len(lens) else 0
    cnt=[0]*(maxl+2)
    for l in lens: cnt[l]+=1
    cnt[0]=0
    code=0; nxt=[0]*(maxl+2)
    for l in range(1,maxl+1):
        code=(code+cnt[l-1])<<1; nxt[l]=code
    tab={}
    for s,l in enumerate(lens):
        if l:
            c=nxt[l]; nxt[l]+=1
            r=int(bin(c)[2:].zfill(l)[::-1],2)
            tab[(l,r)]=s
    return tab,maxl
def dec(br,tab,maxl):
    for l in range(1,maxl+1):
        k=(l,br.peek(l))
        if k in tab:
            br.p+=l; return tab[k]
    raise ValueError("bad code")

def tokens(raw):
    """returns list of (bitpos_start, kind(0 lit,1 match), a, b, outpos) and block boundaries"""
    br=BR(raw); out=0; toks=[]; blocks=[]
    while True:
        last=br.get(1); typ=br.get(2)
        if typ==0:
            br.p=(br.p+7)&~7; n=br.get(16); br.get(16); br.p+=8*n; out+=n; blocks.append(('stored',br.p)); 
        else:
            if typ==1:
                ll=[8]*144+[9]*112+[7]*24+[8]*8; dl=[5]*30
            else:
                hlit=br.get(5)+257; hdist=br.get(5)+1; hclen=br.get(4)+4
                cl=[0]*19
                for i in range(hclen): cl[ZZ[i]]=br.get(3)
                ct,cm=mk(cl); lens=[]
                while len(lens)<hlit+hdist:
                    s=dec(br,ct,cm)
                    if s<16: lens.append(s)
                    elif s==16: lens+= [lens[-1]]*(3+br.get(2))
                    elif s==17: lens+=[0]*(3+br.get(3))
                    else: lens+=[0]*(11+br.get(7))
                ll=lens[:hlit]; dl=lens[hlit:]
            lt,lm=mk(ll); dt,dm=mk(dl)
            blocks.append(('data_start',br.p))
            while True:
                p0=br.p
                s=dec(br,lt,lm)
                if s<256: toks.append((p0,0,s,0,out)); out+=1
                elif s==256: blocks.append(('eob',br.p)); break
                else:
                    ln=LB[s-257]+br.get(LE[s-257]); ds=dec(br,dt,dm); d=DB[ds]+br.get(DE[ds])
                    toks.append((p0,1,ln,d,out)); out+=ln
        if last: break
    return toks,blocks,out

if __name__=="__main__":
    import workloads
    sub=int(sys.argv[1]) if len(sys.argv)>1 else 416
    allb=[]; alll=[]; alld=[]; ntok=[]
    for seed in range(3):
        comps,plains=workloads.c2_streams(1,seed0=seed,fmt="raw",workers=1)
        toks,blocks,out=tokens(comps[0])
        assert out==65536
        t=np.array(toks)
        start=[b for b in blocks if b[0]=='data_start'][0][1]
        lane=(t[:,0]-start)//sub
        nb=np.where(t[:,1]==1,t[:,2],1)
        per=np.bincount(lane,weights=nb); cnt=np.bincount(lane)
        allb.append(per); ntok.append(cnt)
        m=t[t[:,1]==1]
        alll.append(m[:,2]); alld.append(m[:,3])
    per=np.concatenate(allb); cnt=np.concatenate(ntok); L=np.concatenate(alll); D=np.concatenate(alld)
    print("sub-chunk bits",sub,"lanes",len(per),"bytes/lane mean %.1f p50 %.0f p90 %.0f p99 %.0f max %.0f"%(per.mean(),np.median(per),np.percentile(per,90),np.percentile(per,99),per.max()))
    print("tokens/lane mean %.1f p90 %.0f max %d"%(cnt.mean(),np.percentile(cnt,90),cnt.max()))
    # expected max over groups of 32 consecutive lanes
    g=per[:len(per)//32*32].reshape(-1,32); print("warp max/mean bytes: %.2f"%(g.max(axis=1).mean()/per.mean()))
    g=cnt[:len(cnt)//32*32].reshape(-1,32); print("warp max/mean tokens: %.2f"%(g.max(axis=1).mean()/cnt.mean()))
    print("matches: %d, frac of tokens %.2f, len mean %.1f p50 %d p90 %d p99 %d; bytes from matches %.2f"%(len(L),len(L)/cnt.sum(),L.mean(),np.median(L),np.percentile(L,90),np.percentile(L,99),L.sum()/per.sum()))
    for th in (8,16,32,64):
        print("  len>=%d: %.1f%% of matches, %.1f%% of bytes"%(th,100*(L>=th).mean(),100*L[L>=th].sum()/per.sum()))
    print("dist: p10 %d p50 %d p90 %d; <64: %.1f%% <256 %.1f%% <1024 %.1f%% <4096 %.1f%%"%(np.percentile(D,10),np.median(D),np.percentile(D,90),100*(D<64).mean(),100*(D<256).mean(),100*(D<1024).mean(),100*(D<4096).mean()))
