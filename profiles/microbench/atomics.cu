// Micro-benchmarks that decide the deposit design on B200 (run under gpurun; results in profiles/).
//   A  red.global.add.u32 on M random counters            (histogram pass of a counting sort)
//   B  atom.global.add.u32 (with return) on M counters    (slot claim of a counting sort)
//   C  shared-memory atomicAdd(float) (ATOMS.CAST.SPIN loop) vs atomicAdd(int) (native ATOMS.ADD)
//   D  random 12-byte gathers from a 1.6 GB array         (deposit reading particles through a sorted index)
//   E  streaming 12-byte reads                             (deposit reading physically sorted particles)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ unsigned hash(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

__global__ void kA(unsigned* c, unsigned M, long n){ long i=(long)blockIdx.x*blockDim.x+threadIdx.x, s=(long)gridDim.x*blockDim.x;
  for(;i<n;i+=s) atomicAdd(&c[hash((unsigned)i)%M],1u); }
__global__ void kB(unsigned* c, unsigned M, long n, unsigned* out){ long i=(long)blockIdx.x*blockDim.x+threadIdx.x, s=(long)gridDim.x*blockDim.x; unsigned acc=0;
  for(;i<n;i+=s) acc+=atomicAdd(&c[hash((unsigned)i)%M],1u); if(acc==0x12345678u) out[0]=acc; }
// F: same update stream, but the 32 lanes of a warp always fall into 32 distinct banks (row = random, column = lane):
// what a deposit would see if every warp's particles were regrouped by bank first
template<typename T> __global__ void kF(T* out, int cells, int iters){
  extern __shared__ unsigned char raw[]; T* s=(T*)raw;
  for(int i=threadIdx.x;i<cells;i+=blockDim.x) s[i]=0; __syncthreads();
  unsigned h=hash(blockIdx.x*blockDim.x+threadIdx.x+1); const int rows=cells/32-8, lane=threadIdx.x&31;
  for(int it=0;it<iters;it++){ h=hash(h+it); int base=(h%rows)*32+lane;
    #pragma unroll
    for(int q=0;q<8;q++) atomicAdd(&s[base+((q*37)%8)*32], (T)1); }
  __syncthreads(); T acc=0; for(int i=threadIdx.x;i<cells;i+=blockDim.x) acc+=s[i]; if(acc==(T)-1) out[0]=acc; }
template<typename T> __global__ void kC(T* out, int cells, int iters){
  extern __shared__ unsigned char raw[]; T* s=(T*)raw;
  for(int i=threadIdx.x;i<cells;i+=blockDim.x) s[i]=0; __syncthreads();
  unsigned h=hash(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0;it<iters;it++){ h=hash(h+it); int base=h%(cells-8);
    #pragma unroll
    for(int q=0;q<8;q++) atomicAdd(&s[base+((q*37)%8)], (T)1); }
  __syncthreads(); T acc=0; for(int i=threadIdx.x;i<cells;i+=blockDim.x) acc+=s[i]; if(acc==(T)-1) out[0]=acc; }
__global__ void kD(const float* p, const unsigned* idx, long n, float* out){ long i=(long)blockIdx.x*blockDim.x+threadIdx.x, s=(long)gridDim.x*blockDim.x; float acc=0;
  for(;i<n;i+=s){ const float* q=p+3l*idx[i]; acc+=q[0]+q[1]+q[2]; } if(acc==1.2345f) out[0]=acc; }
__global__ void kE(const float* p, long n, float* out){ long i=(long)blockIdx.x*blockDim.x+threadIdx.x, s=(long)gridDim.x*blockDim.x; float acc=0;
  for(;i<n;i+=s){ const float* q=p+3l*i; acc+=q[0]+q[1]+q[2]; } if(acc==1.2345f) out[0]=acc; }
__global__ void kIdx(unsigned* idx, long n){ long i=(long)blockIdx.x*blockDim.x+threadIdx.x; if(i<n) idx[i]=(unsigned)((hash((unsigned)i)*2654435761ull)%n); }

template<typename F> float timeit(F f,int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; } return best; }

int main(){
  const long n=134217728; unsigned *c,*out,*idx; float *p,*fo;
  CK(cudaMalloc(&c,4u<<20)); CK(cudaMalloc(&out,64)); CK(cudaMalloc(&idx,n*4)); CK(cudaMalloc(&p,n*12)); CK(cudaMalloc(&fo,64));
  CK(cudaMemset(c,0,4u<<20)); CK(cudaMemset(p,0,n*12));
  kIdx<<<(n+255)/256,256>>>(idx,n); CK(cudaDeviceSynchronize());
  int sms=148; 
  for(unsigned M: {16384u,131072u,1048576u}){
    float a=timeit([&]{kA<<<sms*16,256>>>(c,M,n);}); float b=timeit([&]{kB<<<sms*16,256>>>(c,M,n,out);});
    printf("A red.global.u32  M=%8u: %7.3f ms  %7.1f Gops/s\n",M,a,n/a/1e6);
    printf("B atom.global.u32 M=%8u: %7.3f ms  %7.1f Gops/s\n",M,b,n/b/1e6);
  }
  for(int cells: {9537, 12635}){
    int iters=2048; size_t sm=cells*4; 
    CK(cudaFuncSetAttribute(kC<float>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm)); CK(cudaFuncSetAttribute(kC<int>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm));
    for(int cps: {1,2,4}){
      int blocks=sms*cps; double ops=(double)blocks*256*iters*8;
      float f=timeit([&]{kC<float><<<blocks,256,sm>>>((float*)fo,cells,iters);}); float i=timeit([&]{kC<int><<<blocks,256,sm>>>((int*)out,cells,iters);});
      printf("C smem atomics cells=%5d CTAs/SM=%d: float(CAS) %7.3f ms %7.1f Gops/s (%.2f ops/clk/SM @1.9GHz) | int(native) %7.3f ms %7.1f Gops/s\n",cells,cps,f,ops/f/1e6,ops/f/1e6/148/1.9,i,ops/i/1e6);
    }
  }
  { int cells=12635, iters=2048; size_t sm=cells*4;
    CK(cudaFuncSetAttribute(kF<float>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm)); CK(cudaFuncSetAttribute(kF<int>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm));
    for(int cps: {2,4}){
      int blocks=sms*cps; double ops=(double)blocks*256*iters*8;
      float f=timeit([&]{kF<float><<<blocks,256,sm>>>((float*)fo,cells,iters);}); float i=timeit([&]{kF<int><<<blocks,256,sm>>>((int*)out,cells,iters);});
      printf("F smem atomics, lanes in distinct banks, cells=%5d CTAs/SM=%d: float(CAS) %7.3f ms %7.1f Gops/s (%.2f ops/clk/SM @1.9GHz) | int(native) %7.3f ms %7.1f Gops/s\n",cells,cps,f,ops/f/1e6,ops/f/1e6/148/1.9,i,ops/i/1e6);
    }
  }
  float d=timeit([&]{kD<<<sms*16,256>>>(p,idx,n,fo);}); float e=timeit([&]{kE<<<sms*16,256>>>(p,n,fo);});
  printf("D random 12B gather: %7.3f ms  (%.1f GB/s useful, idx+payload)\n",d,(n*16.0)/d/1e6);
  printf("E stream 12B read  : %7.3f ms  (%.1f GB/s)\n",e,(n*12.0)/e/1e6);
  return 0;
}
