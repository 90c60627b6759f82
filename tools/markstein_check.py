import numpy as np
rng=np.random.default_rng(0)
def check(a,b):
    import math
    y=1.0/b
    q=a*y
    # exact fma via math.fma (py3.13?) fallback to numpy longdouble? use fractions for exactness on a sample
    return y,q
# use C via ctypes for fma
import ctypes, subprocess, os, tempfile
src=r'''
#include <math.h>
#include <stdint.h>
long check(const double* a, const double* b, long n){
  long bad=0;
  for(long i=0;i<n;i++){
    double y=1.0/b[i];
    double q=a[i]*y;
    double r=fma(-q,b[i],a[i]);
    double q1=fma(r,y,q);
    double ref=a[i]/b[i];
    if(!(q1==ref) && !(q1!=q1 && ref!=ref)) bad++;
  }
  return bad;
}
'''
d=tempfile.mkdtemp()
open(d+'/m.c','w').write(src)
subprocess.check_call(['gcc','-O2','-ffp-contract=off','-shared','-fPIC','-o',d+'/m.so',d+'/m.c','-lm'])
L=ctypes.CDLL(d+'/m.so'); L.check.restype=ctypes.c_long
P=ctypes.POINTER(ctypes.c_double)
n=20_000_000
for name,(a,b) in {
 'fd-like': (rng.standard_normal(n)*10.0**rng.uniform(-14,2,n), 1.4901161193847656e-08*(1+rng.uniform(-1e-7,1e-7,n))),
 'general': (rng.standard_normal(n)*10.0**rng.uniform(-30,30,n), rng.standard_normal(n)*10.0**rng.uniform(-30,30,n)),
 'neg-dx': (rng.standard_normal(n), -1.4901161193847656e-08*(1+rng.uniform(-1e-6,1e-6,n))),
 'zeros': (np.zeros(n), rng.standard_normal(n)*1e-8),
}.items():
    a=np.ascontiguousarray(a); b=np.ascontiguousarray(b)
    print(name, L.check(a.ctypes.data_as(P), b.ctypes.data_as(P), n))
