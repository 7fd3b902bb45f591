#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
typedef struct { double sign[4]; double hpi_inv, hpi, c0,c1,c2,c3,c4, s1,s2,s3; } sincos_t;
static const sincos_t T[2] = {
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
static uint32_t fw(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static uint32_t top12(float x){return (fw(x)>>20)&0x7ff;}
static float poly(double x,double x2,const sincos_t*p,int n){
  if((n&1)==0){ double x3=x*x2; double s1=p->s2+x2*p->s3; double x7=x3*x2; double s=x+x3*p->s1; return (float)(s+x7*s1);}
  else { double x4=x2*x2; double c2=p->c3+x2*p->c4; double c1=p->c0+x2*p->c1; double x6=x4*x2; double c=c1+x4*p->c2; return (float)(c+x6*c2);} }
static double reduce_fast(double x,const sincos_t*p,int*np){ double r=x*p->hpi_inv; int n=((int32_t)r+0x800000)>>24; *np=n; return x-n*p->hpi; }
static float my_sinf(float y){ double x=y; const sincos_t*p=&T[0]; int n;
  if(top12(y)<top12(0x1.921FB6p-1f)){ double s=x*x; if(top12(y)<top12(0x1p-12f)) return y; return poly(x,s,p,0);} 
  x=reduce_fast(x,p,&n); double s=p->sign[n&3]; if(n&2)p=&T[1]; return poly(x*s,x*x,p,n); }
static float my_cosf(float y){ double x=y; const sincos_t*p=&T[0]; int n;
  if(top12(y)<top12(0x1.921FB6p-1f)){ double x2=x*x; if(top12(y)<top12(0x1p-12f)) return 1.0f; return poly(x,x2,p,1);} 
  x=reduce_fast(x,p,&n); double s=p->sign[n&3]; if(n&2)p=&T[1]; return poly(x*s,x*x,p,n^1); }
int main(){ srand(2); long ds=0,dc=0; 
  for(long it=0;it<30000000;it++){ float a=((float)rand()/RAND_MAX*2-1)*3.14159274f; if(it%3==0) a*=1e-3f;
    if(fw(sinf(a))!=fw(my_sinf(a))){ds++; if(ds<4)printf("sin %a: %a vs %a\n",a,sinf(a),my_sinf(a));}
    if(fw(cosf(a))!=fw(my_cosf(a))){dc++; if(dc<4)printf("cos %a: %a vs %a\n",a,cosf(a),my_cosf(a));} }
  printf("replica mismatches: sin %ld cos %ld of 30M\n",ds,dc); }
