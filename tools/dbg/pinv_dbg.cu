#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../mjpl_b200/csrc/vk_core.cuh"
using namespace vk;
__global__ void k(const double* Ain, const double* x, double* y){ double A[6][6]; for(int i=0;i<6;i++)for(int j=0;j<6;j++)A[i][j]=Ain[i*6+j]; sym6_pinv_apply(A,x,y); }
int main(){
  srand(1); double J[6][6], A[36], x[6], yh[6], yd[6];
  for(int i=0;i<6;i++){x[i]=rand()/(double)RAND_MAX-0.5; for(int j=0;j<6;j++)J[i][j]=rand()/(double)RAND_MAX-0.5;}
  for(int i=0;i<6;i++)for(int j=0;j<6;j++){double a=0;for(int c=0;c<6;c++)a+=J[i][c]*J[j][c];A[i*6+j]=a;}
  double Ah[6][6]; for(int i=0;i<6;i++)for(int j=0;j<6;j++)Ah[i][j]=A[i*6+j];
  sym6_pinv_apply(Ah,x,yh);
  double *dA,*dx,*dy; cudaMalloc(&dA,288);cudaMalloc(&dx,48);cudaMalloc(&dy,48);
  cudaMemcpy(dA,A,288,cudaMemcpyHostToDevice);cudaMemcpy(dx,x,48,cudaMemcpyHostToDevice);
  k<<<1,1>>>(dA,dx,dy); printf("err %s\n",cudaGetErrorString(cudaDeviceSynchronize()));
  cudaMemcpy(yd,dy,48,cudaMemcpyDeviceToHost);
  // residual check: A*y should equal x
  for(int t=0;t<2;t++){ double* y=t?yd:yh; double r=0; for(int i=0;i<6;i++){double a=0;for(int j=0;j<6;j++)a+=A[i*6+j]*y[j]; r=fmax(r,fabs(a-x[i]));} printf("%s residual %g  y0 %g\n",t?"dev":"host",r,y[0]); }
  return 0; }
