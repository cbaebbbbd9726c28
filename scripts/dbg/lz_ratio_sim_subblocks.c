// scripts/dbg/lz_ratio_sim_subblocks.c — design-exploration helper: the ratio of the DECOUPLED variant (every warp its own
// sub-block of the 32 KiB half, one shared bucket table, no barriers): half the entries a reader finds are in its future.
// Usage: gcc -O2 -o sim3 lz_ratio_sim_subblocks.c && ./sim3 text.bin <hash bits> <ways> <sub-blocks>
// ratio of the "16 warps, contiguous sub-blocks, shared FIFO-bucket table" scheme
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static int len_sym(int len){ if(len==258) return 285; int x=len-3; int e = x<8?0:(31-__builtin_clz(x))-2; return 257+4*e+(x>>e);} 
static int len_eb(int len){ if(len==258) return 0; int x=len-3; return x<8?0:(31-__builtin_clz(x))-2; }
static int dist_sym(int d){ int D=d-1; if(D<4) return D; int n=31-__builtin_clz(D); return 2*n+((D>>(n-1))&1);} 
static int dist_eb(int d){ int D=d-1; if(D<4) return 0; return (31-__builtin_clz(D))-1; }
static double huff_cost(const uint32_t *f, int n){ double w[600]; int par[600]; int alive[600]; int m=0;
  for(int i=0;i<n;i++) if(f[i]){ w[m]=f[i]; par[m]=-1; alive[m]=1; m++; }
  if(m<2) return m?w[0]:0; int total=m,cnt=m;
  while(cnt>1){ int a=-1,b=-1; for(int i=0;i<total;i++) if(alive[i]){ if(a<0||w[i]<w[a]){b=a;a=i;} else if(b<0||w[i]<w[b]) b=i; }
    w[total]=w[a]+w[b]; par[total]=-1; alive[total]=1; par[a]=total; par[b]=total; alive[a]=alive[b]=0; total++; cnt--; }
  double cost=0; for(int i=0;i<m;i++){ int d=0,k=i; while(par[k]>=0){k=par[k]; d++;} if(d>15)d=15; cost+=w[i]*d; } return cost; }
static int mlenf(const uint8_t *d, long a, long b, int max){ int l=0; while(l<max && d[a+l]==d[b+l]) l++; return l; }
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); size_t n=ftell(f); fseek(f,0,SEEK_SET); uint8_t*data=malloc(n+300); fread(data,1,n,f); memset(data+n,0,300);
  int HB=atoi(argv[2]), W=atoi(argv[3]), NSUB=atoi(argv[4]), HALF=32768, lazy=1; int SUB=HALF/NSUB;
  int HS=1<<HB; int64_t *tab=malloc(sizeof(int64_t)*HS*W); for(int i=0;i<HS*W;i++) tab[i]=-1;
  uint8_t *cntb=calloc(HS,1);
  int *best_l=malloc(sizeof(int)*HALF), *best_d=malloc(sizeof(int)*HALF);
  double total_bits=0; uint32_t fl[288],fd[32]; double extra=0; memset(fl,0,sizeof fl); memset(fd,0,sizeof fd);
  for(size_t h0=0; h0<n; h0+=HALF){
    int hl = n-h0<HALF? n-h0:HALF;
    // lockstep over time
    for(int t=0;t<SUB/32;t++){
      for(int w=0;w<NSUB;w++){ // insert 32 positions of each sub-block
        for(int k=0;k<32;k++){ long p=h0+(long)w*SUB+32*t+k; if(p-h0>=hl) continue; int send = (w+1)*SUB<hl?(w+1)*SUB:hl; if((p-h0)+4>send) continue;
          uint32_t v=data[p]|data[p+1]<<8|data[p+2]<<16|(uint32_t)data[p+3]<<24; uint32_t h=(v*0x9e3779b1u)>>(32-HB); tab[h*W+(cntb[h]++ % W)]=p; } }
      for(int w=0;w<NSUB;w++) for(int k=0;k<32;k++){ long p=h0+(long)w*SUB+32*t+k; int rel=p-h0; if(rel>=hl) continue; int send=(w+1)*SUB<hl?(w+1)*SUB:hl; int avail=send-rel; best_l[rel]=0; best_d[rel]=0; if(avail<4) continue;
          int max=avail<258?avail:258; uint32_t v=data[p]|data[p+1]<<8|data[p+2]<<16|(uint32_t)data[p+3]<<24; uint32_t h=(v*0x9e3779b1u)>>(32-HB);
          for(int e=0;e<W;e++){ int64_t q=tab[h*W+e]; if(q<0||q>=p||p-q>32768) continue; int l=mlenf(data,q,p,max); if(l>best_l[rel]||(l==best_l[rel]&&p-q<best_d[rel])){best_l[rel]=l;best_d[rel]=p-q;} }
          if(best_l[rel]<4) best_l[rel]=0; }
    }
    // parse per sub-block (lazy one-step as kernel: yield if next longer)
    for(int w=0;w<NSUB;w++){ int s0=w*SUB, s1=(w+1)*SUB<hl?(w+1)*SUB:hl; int p=s0; while(p<s1){ int l=best_l[p]; if(l && lazy && p+1<s1 && best_l[p+1]>l && l<16) l=0; if(l){ fl[len_sym(l)]++; fd[dist_sym(best_d[p])]++; extra+=len_eb(l)+dist_eb(best_d[p]); p+=l;} else { fl[data[h0+p]]++; p++; } } }
    if(((h0/HALF)&1)==1 || h0+HALF>=n){ fl[256]=1; double bits=huff_cost(fl,286)+huff_cost(fd,30)+extra+17+57+480+42; total_bits+=bits; memset(fl,0,sizeof fl); memset(fd,0,sizeof fd); extra=0; }
  }
  printf("hb %d ways %d nsub %d: ratio %.3f\n",HB,W,NSUB,n*8.0/total_bits); }
