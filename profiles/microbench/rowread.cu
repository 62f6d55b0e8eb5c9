// Memory-side ceiling of the ring kernel's access pattern: CTAs of 256 threads read 2 KB row segments
// (element t+1 of each 257-element complex64 row) in (a) natural, (b) r2-sorted, (c) random row order,
// with no arithmetic beyond a running sum.  B200, 512^3 half-spectrum (535 MB useful).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
template<int U>
__global__ void __launch_bounds__(256) rowread(const float2* __restrict__ dk, const int* __restrict__ rows, int nrows, int rows_per_cta, int nz, float* out){
  int i0=blockIdx.x*rows_per_cta, i1=min(nrows,i0+rows_per_cta); float acc=0; int t=threadIdx.x;
  for(int i=i0;i<i1;i+=U){ float2 v[U];
    #pragma unroll
    for(int u=0;u<U;u++){ int r = (i+u<i1)? rows[i+u] : rows[i0]; v[u]=dk[(long)r*nz + 1 + t]; }
    #pragma unroll
    for(int u=0;u<U;u++) acc+=v[u].x+v[u].y; }
  if(acc==1.2345f) out[0]=acc; }
template<typename F> float timeit(F f,int reps=5){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; } return best; }
int main(){ const int N=512, nz=N/2+1, nrows=N*N; float2* dk; int* d_rows; float* out;
  CK(cudaMalloc(&dk,(size_t)nrows*nz*8)); CK(cudaMemset(dk,0,(size_t)nrows*nz*8)); CK(cudaMalloc(&d_rows,nrows*4)); CK(cudaMalloc(&out,64));
  std::vector<int> nat(nrows), srt(nrows), rnd(nrows); std::iota(nat.begin(),nat.end(),0); srt=nat; rnd=nat;
  auto r2=[&](int r){int kx=r/N, ky=r%N; if(kx>N/2)kx-=N; if(ky>N/2)ky-=N; return kx*kx+ky*ky;};
  std::stable_sort(srt.begin(),srt.end(),[&](int a,int b){return r2(a)<r2(b);});
  srand(1); for(int i=nrows-1;i>0;i--){int j=rand()%(i+1); std::swap(rnd[i],rnd[j]);}
  double gb=8.0*nrows*255/1e9;
  const char* names[3]={"natural","r2-sorted","random"}; std::vector<int>* ord[3]={&nat,&srt,&rnd};
  for(int o=0;o<3;o++){ CK(cudaMemcpy(d_rows,ord[o]->data(),nrows*4,cudaMemcpyHostToDevice));
    for(int ctas: {444, 888, 1776, 3552}){ int rpc=(nrows+ctas-1)/ctas;
      float a=timeit([&]{rowread<8><<<ctas,256>>>(dk,d_rows,nrows,rpc,nz,out);}); float b=timeit([&]{rowread<16><<<ctas,256>>>(dk,d_rows,nrows,rpc,nz,out);});
      printf("%-10s ctas=%4d  U=8: %.4f ms %5.0f GB/s | U=16: %.4f ms %5.0f GB/s\n",names[o],ctas,a,gb/a*1e3,b,gb/b*1e3); } }
  return 0; }
