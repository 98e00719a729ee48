#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
typedef struct { uint32_t d, m, l; } FD;
static FD mk(uint32_t d){ FD f; f.d=d; f.l = d>1u ? 32u-(uint32_t)__builtin_clz(d-1u) : 0u; f.m=(uint32_t)((((((uint64_t)1)<<f.l)-d)<<32)/d)+1u; return f;}
static uint32_t dv(uint32_t n, FD f){ uint32_t hi=(uint32_t)(((uint64_t)f.m*n)>>32); return (hi+n)>>f.l; }
int main(){ uint64_t bad=0; uint32_t ds[]={1,2,3,4,5,6,7,12,64,197,208,394,768,1182,4728,50432,65535,65536,65537,1000003,(1u<<30)-1,(1u<<30),(1u<<30)+1,(1u<<31)-1,(1u<<31)};
 for(unsigned i=0;i<sizeof(ds)/4;i++){ FD f=mk(ds[i]); for(uint64_t n=0;n<(1ull<<31);n+= (n<100000?1:9973)){ if(dv((uint32_t)n,f)!=(uint32_t)(n/ds[i])) {bad++; if(bad<5) printf("bad d=%u n=%llu\n",ds[i],(unsigned long long)n);} } uint32_t n=(1u<<31)-1; if(dv(n,f)!=n/ds[i]) {bad++; printf("bad top d=%u\n",ds[i]);} }
 srand(1); for(int i=0;i<2000000;i++){ uint32_t d=(uint32_t)(rand()%((i&1)?100000:2000000000))+1; uint32_t n=(uint32_t)(((uint64_t)rand()<<16 ^ rand()) & 0x7fffffff); FD f=mk(d); if(dv(n,f)!=n/d){bad++; if(bad<10)printf("bad d=%u n=%u\n",d,n);} }
 printf("bad=%llu\n",(unsigned long long)bad); return 0; }
