// Exhaustive check behind mlp_act / mlp_act2 (csrc/mlp.cuh): for every f32 v < -17.5 the clamped ELU evaluation returns exactly -1.0f,
// so the saturation select of det::expm1f_ is redundant there.  gcc -O2 -ffp-contract=off tools/elu_check.c -lm && ./a.out
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static float poly(float r){float p=1.9875691500E-4f;p=fmaf(p,r,1.3981999507E-3f);p=fmaf(p,r,8.3334519073E-3f);p=fmaf(p,r,4.1665795894E-2f);p=fmaf(p,r,1.6666665459E-1f);p=fmaf(p,r,5.0000001201E-1f);volatile float z=r*r;return fmaf(p,z,r);}
static float e_of(float v){
  const float MAGIC=12582912.0f;
  float xx=fmaxf(v,-20.0f); volatile float m=xx*1.44269504088896341f; volatile float tm=m+MAGIC; volatile float n=tm-MAGIC;
  float r=fmaf(n,-0.693359375f,xx); r=fmaf(n,2.12194440e-4f,r);
  float pl=poly(r); float t=u2f((f2u(tm)<<23)+0x3F800000u); volatile float tm1=t-1.0f; return fmaf(pl,t,tm1);
}
int main(){
  // all negative floats from -17.5 down to -inf
  uint32_t lo=f2u(-17.5f), hi=f2u(-INFINITY); long bad=0,cnt=0;
  for(uint32_t u=lo+1;u<=hi;++u){float v=u2f(u); float e=e_of(v); ++cnt; if(e!=-1.0f){ if(bad<5)printf("v=%a e=%a\n",v,e); ++bad;}}
  printf("checked %ld values below -17.5: %ld differ from -1\n",cnt,bad);
  // v == -17.5 itself is NOT < -17.5: takes e; fine either way
  return bad!=0;
}
