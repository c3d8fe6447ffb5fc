/* ORACLE tooling (test infrastructure): checks that the division sequences the CUDA kernels use
 * (se3ds_b200/csrc/canon_math.cuh: div_const = one Markstein correction for the constants 2*pi, pi,
 * 255; div_rcp = two corrections for run-time denominators) return exactly the IEEE quotient.
 * Exhaustive over all 2^23 mantissas (times a few binades) for the constants, random for a/R.
 * usage: divcheck [binades] [random_millions]; prints the mismatch counts, exit code 1 on any. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float div3(float a,float b,float y){float q0=a*y;float r=fmaf(-b,q0,a);return fmaf(r,y,q0);}
static inline float div5(float a,float b,float y){float q0=a*y;float r0=fmaf(-b,q0,a);float q1=fmaf(r0,y,q0);float r1=fmaf(-b,q1,a);return fmaf(r1,y,q1);}
static uint64_t s=88172645463325252ull; static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
int main(int argc,char**argv){int nb=argc>1?atoi(argv[1]):31; long nr=(argc>2?atol(argv[2]):400)*1000000L; long total_bad=0;
  // exhaustive: constants, all mantissas over several binades
  float consts[4]={(float)(2*3.141592653589793),(float)3.141592653589793,255.0f,20.0f};
  for(int c=0;c<4;c++){float b=consts[c];float y=1.0f/b;long bad3=0,bad5=0;
    for(int e=130-nb+1;e<=130;e++) for(uint32_t m=0;m<(1u<<23);m++){float a=u2f(((uint32_t)e<<23)|m);float t=a/b;
      if(f2u(div3(a,b,y))!=f2u(t))bad3++; if(f2u(div5(a,b,y))!=f2u(t))bad5++;}
    printf("const %.9g y=%a div3_bad %ld div5_bad %ld binades %d\n",b,y,bad3,bad5,nb); total_bad+=bad3+bad5;}
  // small ints / 255
  {float b=255.f,y=1.0f/b;int bad=0;for(int i=-1;i<=255;i++){if(f2u(div3((float)i,b,y))!=f2u((float)i/b))bad++;}printf("ints/255 div3_bad %d\n",bad); total_bad+=bad;}
  // random: a/R with a=R*u  and a=Z (|Z|<=R)
  long bad3=0,bad5=0,n=0;
  for(long i=0;i<nr;i++){uint64_t r=rnd();float R=u2f((uint32_t)((118u+((r>>23)&15))<<23)|(uint32_t)(r&0x7fffff));
    float u=((float)(int32_t)(r>>32))*(1.0f/2147483648.0f); float a=(i&1)?R*u:R*u*0.999f; float y=1.0f/R; float t=a/R;
    if(f2u(div3(a,R,y))!=f2u(t))bad3++; if(f2u(div5(a,R,y))!=f2u(t))bad5++; n++;}
  printf("random a/R n=%ld div3_bad %ld div5_bad %ld\n",n,bad3,bad5); total_bad+=bad5;
  return total_bad?1:0;}
