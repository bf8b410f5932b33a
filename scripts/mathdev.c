#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline double u2d(uint64_t u){double d; memcpy(&d,&u,8); return d;}
static inline uint64_t d2u(double d){uint64_t u; memcpy(&u,&d,8); return u;}
static const double EC[] = {
  /* c11..c2 */
  0,0
};
double my_exp(double x){
  const double L2E=u2d(0x3ff71547652b82feULL), SH=6755399441055744.0;
  const double LN2H=u2d(0x3fe62e42fefa39efULL), LN2L=u2d(0x3c7abc9e3b39803fULL);
  if (!(fabs(x) < 708.0)) return exp(x);
  double t=fma(x,L2E,SH); double n=t-SH;
  double r=fma(n,-LN2H,x); r=fma(n,-LN2L,r);
  double p=u2d(0x3e5ade1569ce2bdfULL);
  p=fma(p,r,u2d(0x3e928af3fca213eaULL));
  p=fma(p,r,u2d(0x3ec71dee62401315ULL));
  p=fma(p,r,u2d(0x3efa01997c89eb71ULL));
  p=fma(p,r,u2d(0x3f2a01a014761f65ULL));
  p=fma(p,r,u2d(0x3f56c16c1852b7afULL));
  p=fma(p,r,u2d(0x3f81111111122322ULL));
  p=fma(p,r,u2d(0x3fa55555555502a1ULL));
  p=fma(p,r,u2d(0x3fc5555555555511ULL));
  p=fma(p,r,u2d(0x3fe000000000000bULL));
  p=fma(p,r,1.0);
  p=fma(p,r,1.0);
  int32_t k=(int32_t)(d2u(t)&0xffffffffu);
  uint64_t u=d2u(p); u += ((uint64_t)(uint32_t)k)<<52; /* hi word += k<<20 */
  return u2d(u);
}
static double qdiv(double a,double b){ /* stand-in for rcp-seeded quotient: two newton on 1/b from an approximate seed + residual */
  double r=(double)(float)(1.0/b); r=fma(r,fma(-b,r,1.0),r); r=fma(r,fma(-b,r,1.0),r);
  double q=a*r; return fma(fma(-b,q,a),r,q);
}
double my_log(double x){
  const double ln2_hi=6.93147180369123816490e-01, ln2_lo=1.90821492927058770002e-10;
  const double Lg1=6.666666666666735130e-01,Lg2=3.999999999940941908e-01,Lg3=2.857142874366239149e-01,Lg4=2.222219843214978396e-01,Lg5=1.818357216161805012e-01,Lg6=1.531383769920937332e-01,Lg7=1.479819860511658591e-01;
  uint64_t u=d2u(x); int32_t hx=(int32_t)(u>>32);
  if (hx < 0x00100000 || hx >= 0x7ff00000) return log(x);  /* zero, denormal, negative, inf, nan */
  int k=(hx>>20)-1023; hx&=0x000fffff;
  int i=(hx+0x95f64)&0x100000;
  u=((uint64_t)(uint32_t)(hx|(i^0x3ff00000))<<32)|(u&0xffffffffu); k+=(i>>20);
  double m=u2d(u); double f=m-1.0;
  double s=qdiv(f,2.0+f); double dk=(double)k; double z=s*s; double w=z*z;
  double t1=w*fma(w,fma(w,Lg6,Lg4),Lg2);
  double t2=z*fma(w,fma(w,fma(w,Lg7,Lg5),Lg3),Lg1);
  double R=t2+t1; double hfsq=0.5*f*f;
  return dk*ln2_hi-((hfsq-(s*(hfsq+R)+dk*ln2_lo))-f);
}
static double ulp_err(double got,long double ref){ if (got==(double)ref) return 0; double u=fabs(nextafter((double)ref,INFINITY)-(double)ref); return (double)fabsl(((long double)got-ref)/u); }
int main(){
  srand48(1); double me=0,ml=0, ge=0, gl=0; int ne=0;
  for (long i=0;i<20000000;i++){
    double x=(drand48()*2-1)*700.0; if (i&1) x=(drand48()*2-1)*2.0; if ((i&7)==3) x=-drand48()*50;
    double e=ulp_err(my_exp(x),expl((long double)x)); if (e>me) me=e; if (my_exp(x)!=exp(x)) ne++;
    double g=ulp_err(exp(x),expl((long double)x)); if (g>ge) ge=g;
    double y=exp((drand48()*2-1)*(i&2?1.0:600.0)); if ((i&15)==5) y=1.0+(drand48()-0.5)*1e-3;
    double l=ulp_err(my_log(y),logl((long double)y)); if (l>ml) ml=l;
    g=ulp_err(log(y),logl((long double)y)); if (g>gl) gl=g;
  }
  printf("exp max ulp err %.3f (glibc %.3f) differs from glibc in %d; log max ulp err %.3f (glibc %.3f)\n",me,ge,ne,ml,gl);
  printf("%a %a %a\n", my_exp(0.0), my_exp(-745.0), my_exp(710.0)); printf("%a %a %a\n", my_log(1.0), my_log(0.0), my_log(1e-310));
  return 0;
}
