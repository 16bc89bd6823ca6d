"""device acme_exp vs the CUDA library exp: exercised through the diode law: run the TPI kernel
(acme_exp); here: a direct sweep through a tiny nvcc-built test kernel."""
import subprocess, os, sys
src = r'''
#include <cstdio>
#include <cmath>
#include <cstring>
#include "ELEMENTS_CUH"
__global__ void k(const double* x, double* a, double* b, int n){int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n){a[i]=acme::acme_exp(x[i], acme::ACME_EXPC); b[i]=exp(x[i]);}}
int main(){ const int n=1<<22; double *x,*a,*b; cudaMallocManaged(&x,n*8); cudaMallocManaged(&a,n*8); cudaMallocManaged(&b,n*8);
 for(int i=0;i<n;i++){ double t=(double)i/n; x[i]= (i%4==0)? -760+1520*t : (i%4==1)? -40+80*t : (i%4==2)? 700+50*t : -700-50*t; }
 x[0]=NAN; x[1]=INFINITY; x[2]=-INFINITY; x[3]=0.0; x[4]=-0.0; x[5]=709.782712893384; x[6]=709.7827128933841; x[7]=-745.1332191019412; x[8]=-745.1332191019411; x[9]=1e-320;
 k<<<(n+255)/256,256>>>(x,a,b,n); cudaDeviceSynchronize();
 long bad=0; long long maxulp=0; for(int i=0;i<n;i++){ bool same = (a[i]==b[i]) || (a[i]!=a[i] && b[i]!=b[i]); if(!same){ long long ia,ib; memcpy(&ia,&a[i],8); memcpy(&ib,&b[i],8); long long d = ia>ib? ia-ib: ib-ia; if (d>maxulp) maxulp=d; if(d>1){ if(bad<5) printf("x=%.17g acme=%.17g lib=%.17g ulps=%lld\n",x[i],a[i],b[i],d); bad++; } } }
 printf("exp: max ulp distance to the library exp %lld; > 1 ulp: %ld of %d\n", maxulp, bad, n); return bad!=0; }
'''
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
open('/tmp/exp_test.cu','w').write(src.replace('ELEMENTS_CUH', os.path.join(root, 'acme.jl_b200', 'csrc', 'elements.cuh')))
subprocess.check_call(['nvcc','-std=c++17','-O2','-gencode','arch=compute_100a,code=sm_100a','-o','/tmp/exp_test','/tmp/exp_test.cu'])
sys.exit(subprocess.call(['/tmp/exp_test']))
