#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static const float atanhi[] = { 4.6364760399e-01, 7.8539812565e-01, 9.8279368877e-01, 1.5707962513e+00 };
static const float atanlo[] = { 5.0121582440e-09, 3.7748947079e-08, 3.4473217170e-08, 7.5497894159e-08 };
static const float aT[] = { 3.3333334327e-01, -2.0000000298e-01, 1.4285714924e-01, -1.1111110449e-01, 9.0908870101e-02, -7.6918758452e-02, 6.6610731184e-02, -5.8335702866e-02, 4.9768779427e-02, -3.6531571299e-02, 1.6285819933e-02 };
static uint32_t fw(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static float fd_atanf(float x){
  float w,s1,s2,z; int32_t ix,hx,id; hx=(int32_t)fw(x); ix=hx&0x7fffffff;
  if(ix>=0x4c000000){ if(hx>0) return atanhi[3]+atanlo[3]; else return -atanhi[3]-atanlo[3]; }
  if(ix<0x3ee00000){ if(ix<0x31000000) return x; id=-1; }
  else { x=fabsf(x);
    if(ix<0x3f980000){ if(ix<0x3f300000){id=0;x=(2.0f*x-1.0f)/(2.0f+x);} else {id=1;x=(x-1.0f)/(x+1.0f);} }
    else { if(ix<0x401c0000){id=2;x=(x-1.5f)/(1.0f+1.5f*x);} else {id=3;x=-1.0f/x;} } }
  z=x*x; w=z*z;
  s1=z*(aT[0]+w*(aT[2]+w*(aT[4]+w*(aT[6]+w*(aT[8]+w*aT[10])))));
  s2=w*(aT[1]+w*(aT[3]+w*(aT[5]+w*(aT[7]+w*aT[9]))));
  if(id<0) return x-x*(s1+s2);
  z=atanhi[id]-((x*(s1+s2)-atanlo[id])-x);
  return (hx<0)?-z:z;
}
static float fd_atan2f(float y,float x){
  const float pi=3.1415927410e+00f, pi_lo=-8.7422776573e-08f, pi_o_2=1.5707963705e+00f;
  int32_t hx=(int32_t)fw(x),hy=(int32_t)fw(y),ix=hx&0x7fffffff,iy=hy&0x7fffffff; float z;
  if(hx==0x3f800000) return fd_atanf(y);
  int m=((hy>>31)&1)|((hx>>30)&2);
  if(iy==0){ switch(m){case 0: case 1: return y; case 2: return pi+1e-30f; default: return -pi-1e-30f;} }
  if(ix==0) return (hy<0)? -pi_o_2-1e-30f: pi_o_2+1e-30f;
  int k=(iy-ix)>>23;
  if(k>60) z=pi_o_2+0.5f*pi_lo; else if(hx<0&&k<-60) z=0.0f; else z=fd_atanf(fabsf(y/x));
  switch(m){case 0:return z; case 1:return -z; case 2:return pi-(z-pi_lo); default:return (z-pi_lo)-pi;}
}
int main(){
  srand(1); long n=0,da=0,db=0,maxa=0,maxb=0; 
  for(long it=0;it<20000000;it++){
    int R=(it&1)?1900000:20000;
    float y=(float)((rand()%(2*R+1))-R), x=(float)((rand()%(2*R+1))-R);
    float g=atan2f(y,x); float a=(float)atan2((double)y,(double)x); float b=fd_atan2f(y,x);
    n++; if(fw(g)!=fw(a)){da++;} if(fw(g)!=fw(b)){db++; if(db<5)printf("b mismatch y=%g x=%g g=%a b=%a\n",y,x,g,b);}
  }
  printf("n=%ld  glibc!=double-rounded: %ld   glibc!=fdlibm-replica: %ld\n",n,da,db);
  // sin/cos
  long ds=0,dc=0; 
  for(long it=0;it<20000000;it++){ float a=((float)rand()/RAND_MAX*2-1)*3.14159274f; 
    if(fw(sinf(a))!=fw((float)sin((double)a)))ds++; if(fw(cosf(a))!=fw((float)cos((double)a)))dc++; }
  printf("sinf mismatches %ld cosf mismatches %ld of 20M\n",ds,dc);
  // lround trick exhaustive on [0,64)
  long bad=0; for(uint32_t u=0;u<0x42800000u;u++){ float v; memcpy(&v,&u,4); float t=v+0x1.fffffep-2f; int r=(int)t; if(r!=(int)lroundf(v)){bad++; if(bad<5)printf("bad %a\n",v);} float nv=-v; t=nv-0x1.fffffep-2f; r=(int)t; if(r!=(int)lroundf(nv)) bad++; }
  printf("lround trick bad=%ld\n",bad);
}
