// scripts/dbg/lz_ratio_sim.c — design-exploration helper (not product, not test): the compression ratio of LZ77 parses under
// different match-finder shapes (bucket count, ways, chains, window carried or not, insertion per position or per step of P
// positions), costed with real Huffman code lengths per chunk.  It predicted the ratios of deflate_cta_kernel before the kernel
// existed.  Usage: gcc -O2 -o sim lz_ratio_sim.c -lm && ./sim text.bin
// dev tool: ratio of LZ77 parses under different match-finder shapes (dynamic Huffman cost per chunk)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
typedef struct { int chunk, hist, hb, ways, lazy, chain, nice, step; } Cfg;
static int len_sym(int len){ if(len==258) return 285; int x=len-3; int e = x<8?0:(31-__builtin_clz(x))-2; return 257+4*e+(x>>e);} 
static int len_eb(int len){ if(len==258) return 0; int x=len-3; return x<8?0:(31-__builtin_clz(x))-2; }
static int dist_sym(int d){ int D=d-1; if(D<4) return D; int n=31-__builtin_clz(D); return 2*n+((D>>(n-1))&1);} 
static int dist_eb(int d){ int D=d-1; if(D<4) return 0; return (31-__builtin_clz(D))-1; }
static double huff_cost(const uint32_t *f, int n, int maxbits){ // sum f*len using real Huffman (no limit handling beyond clamp)
  // simple O(n^2) merge
  double w[600]; int par[600]; int alive[600]; int m=0; int idx[600];
  for(int i=0;i<n;i++) if(f[i]){ w[m]=f[i]; par[m]=-1; alive[m]=1; idx[m]=i; m++; }
  if(m==0) return 0; if(m==1) return f[idx[0]];
  int total=m; int cnt=m;
  while(cnt>1){ int a=-1,b=-1; for(int i=0;i<total;i++) if(alive[i]){ if(a<0||w[i]<w[a]){b=a;a=i;} else if(b<0||w[i]<w[b]) b=i; }
    w[total]=w[a]+w[b]; par[total]=-1; alive[total]=1; par[a]=total; par[b]=total; alive[a]=alive[b]=0; total++; cnt--; }
  double cost=0; for(int i=0;i<m;i++){ int d=0; int k=i; while(par[k]>=0){k=par[k]; d++;} if(d>maxbits) d=maxbits; cost+=w[i]*d; }
  return cost; }
static int mlen(const uint8_t *d, int a, int b, int max){ int l=0; while(l<max && d[a+l]==d[b+l]) l++; return l; }
static double run(const uint8_t *data, size_t n, Cfg c){
  double total_bits=0;
  int HS=1<<c.hb;
  uint32_t *tab = malloc(sizeof(uint32_t)*HS*c.ways);
  int32_t *prev = malloc(sizeof(int32_t)*(c.chunk+c.hist+8));
  int32_t *head = malloc(sizeof(int32_t)*HS);
  for(size_t c0=0;c0<n;c0+=c.chunk){
    int len = n-c0<c.chunk? n-c0: c.chunk; int hist = c0>=c.hist? c.hist: c0; const uint8_t *d = data+c0; // positions relative: -hist..len
    uint32_t fl[288]={0}, fd[32]={0}; double extra=0;
    for(int i=0;i<HS*c.ways;i++) tab[i]=0x7fffffff; for(int i=0;i<HS;i++) head[i]=-1000000;
    #define H(p) ((( (uint32_t)d[p] | ((uint32_t)d[(p)+1]<<8) | ((uint32_t)d[(p)+2]<<16) | ((uint32_t)d[(p)+3]<<24)) * 0x9e3779b1u) >> (32-c.hb))
    #define INS(p) do{ if((p)+3<len){ uint32_t h=H(p); if(c.chain){ prev[(p)+c.hist]=head[h]; head[h]=(p);} else { uint32_t *b=tab+h*c.ways; for(int w=c.ways-1;w>0;w--) b[w]=b[w-1]; b[0]=(uint32_t)((p)+c.hist);} } }while(0)
    for(int p=-hist;p<0;p++) INS(p);
    // best match at p (before inserting p)
    int p=0; int pend_len=0,pend_dist=0; 
    // compute best for all positions lazily via function
    #define BEST(p,bl,bd) do{ bl=0;bd=0; int max=len-(p)<258?len-(p):258; if(max>=4){ uint32_t h=H(p); if(c.chain){ int q=head[h]; int k=c.chain; while(q>=-hist && (p)-q<=32768 && k--){ int l=mlen(d,q,(p),max); if(l>bl){bl=l;bd=(p)-q; if(l>=c.nice) break;} q=prev[q+c.hist]; } } else { uint32_t *b=tab+h*c.ways; for(int w=0;w<c.ways;w++){ if(b[w]==0x7fffffff) break; int q=(int)b[w]-c.hist; if(q>=(p)) continue; if((p)-q>32768) break; int l=mlen(d,q,(p),max); if(l>bl){bl=l;bd=(p)-q;} } } if(bl<3 || (bl==3&&bd>=4096)) {bl=0;} if(bl==3 && !(bd<4096)) bl=0; if (bl<4 && !(bl==3&&bd<4096)) bl=0; } }while(0)
    int ins_upto=0;
    while(p<len){
      if(c.step){ while(ins_upto<len && ins_upto < (p/c.step+1)*c.step){ INS(ins_upto); ins_upto++; } }
      int bl,bd; BEST(p,bl,bd); if(!c.step) INS(p);
      if(bl && c.lazy){ // look at p+1
        while(p+1<len){ if(c.step){ while(ins_upto<len && ins_upto < ((p+1)/c.step+1)*c.step){ INS(ins_upto); ins_upto++; } } int nl,nd; BEST(p+1,nl,nd); if(nl>bl){ fl[d[p]]++; p++; if(!c.step) INS(p); else { while(ins_upto<len && ins_upto < (p/c.step+1)*c.step){ INS(ins_upto); ins_upto++; } } bl=nl; bd=nd; } else break; if(c.lazy==1) { /* one step only: but chain of yields allowed */ } }
      }
      if(bl){ fl[len_sym(bl)]++; fd[dist_sym(bd)]++; extra+=len_eb(bl)+dist_eb(bd); if(!c.step) for(int k=1;k<bl;k++) INS(p+k); p+=bl; }
      else { fl[d[p]]++; p++; }
    }
    fl[256]=1;
    double bits = huff_cost(fl,286,15)+huff_cost(fd,30,15)+extra+ 17 + 19*3 + 60*8 /*header approx*/ + 42 /*marker*/;
    double stored = 8.0*(len+5);
    total_bits += bits<stored?bits:stored;
  }
  free(tab); free(prev); free(head);
  return n*8.0/total_bits;
}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); size_t n=ftell(f); fseek(f,0,SEEK_SET); uint8_t*data=malloc(n+16); fread(data,1,n,f); memset(data+n,0,16);
  Cfg cfgs[] = {
    {65536,0,11,2,1,0,0,0},
    {32768,32768,11,8,1,0,0,512}, {32768,32768,12,4,1,0,0,512}, {32768,32768,12,8,1,0,0,512}, {32768,32768,11,16,1,0,0,512},
    {32768,32768,12,8,1,0,0,0}, {32768,32768,12,8,1,0,0,1024}, {32768,32768,13,8,1,0,0,512}, {32768,32768,12,16,1,0,0,512},
    {65536,32768,12,8,1,0,0,1024}, {65536,32768,13,4,1,0,0,1024}, {65536,32768,13,8,1,0,0,1024},
    {32768,0,12,8,1,0,0,512},
  };
  for(unsigned i=0;i<sizeof(cfgs)/sizeof(cfgs[0]);i++){ Cfg c=cfgs[i]; printf("chunk %7d hist %5d hb %2d ways %2d lazy %d chain %3d nice %3d step %4d: ratio %.3f\n",c.chunk>1<<20?0:c.chunk,c.hist,c.hb,c.ways,c.lazy,c.chain,c.nice,c.step,run(data,n,c)); fflush(stdout);} 
}
