// TEST CODE: accuracy of inflx_atan_tan (csrc/inflx_device.cuh, the epilogue's delta = atan(y) and
// tan(delta)) against libquadmath, on the HOST build of the device header (same fma / add / mul
// sequence as the device; the two hardware seeds are IEEE values here).  argv[1] = samples.
#include "cuda_host_shim.h"
#include INFLX_DEVICE_HEADER
#include <quadmath.h>
#include <stdio.h>
#include <stdlib.h>
#include <random>
int main(int argc, char** argv){
  std::mt19937_64 rng(4);
  std::uniform_real_distribution<double> u(0,1);
  double worst_d=0, worst_t=0; 
  const int n = argc > 1 ? atoi(argv[1]) : 2000000;
  for(int k=0;k<n;k++){
    double y;
    int f=k%4;
    if(f==0) y=pow(10.0, -2+4*u(rng)); else if(f==1) y=u(rng); else if (f==2) y=1.0/(u(rng)+1e-9); else y=pow(10.0,-30+60*u(rng));
    double yinv=1.0/y;
    double d,t; inflx_atan_tan(y,yinv,d,t,inflx_exact());
    __float128 td=atanq((__float128)y);
    double ulp=fabs(nextafter((double)td, INFINITY)-(double)td);
    double e=fabs((double)(((__float128)d-td)/ulp)); if(e>worst_d) worst_d=e;
    __float128 tt=tanq((__float128)d);
    double ulpt=fabs(nextafter((double)tt, INFINITY)-(double)tt);
    double et=fabs((double)(((__float128)t-tt)/ulpt)); if(et>worst_t && y<1e15) worst_t=et;
  }
  printf("chains=%d n=%d worst_atan_ulp=%.4f worst_tan_ulp=%.4f\n", INFLX_ATAN_CHAINS, n, worst_d, worst_t);
}
