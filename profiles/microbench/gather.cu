// Random 12-byte gathers through a sorted index: effect of cudaLimitMaxL2FetchGranularity and of the
// access shape (3 scalar loads vs. one 16-byte-aligned window + shuffle-free select).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ unsigned hash(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }
__global__ void kIdx(unsigned* idx, long n){ long i=(long)blockIdx.x*blockDim.x+threadIdx.x; if(i<n) idx[i]=(unsigned)((hash((unsigned)i)*2654435761ull)%n); }
// gather + write sorted copy (what a "physical sort" of the payload costs)
__global__ void kGather(const float* __restrict__ p, const unsigned* __restrict__ idx, long n, float* __restrict__ out){
  long i=(long)blockIdx.x*blockDim.x+threadIdx.x, s=(long)gridDim.x*blockDim.x;
  for(;i<n;i+=s){ const float* q=p+3l*idx[i]; float x=__ldg(q),y=__ldg(q+1),z=__ldg(q+2); out[3*i]=x; out[3*i+1]=y; out[3*i+2]=z; } }
// same, 4 particles in flight per thread
__global__ void kGather4(const float* __restrict__ p, const unsigned* __restrict__ idx, long n, float* __restrict__ out){
  long s=(long)gridDim.x*blockDim.x; long i=(long)blockIdx.x*blockDim.x+threadIdx.x;
  for(;i+3*s<n;i+=4*s){ float v[4][3];
    #pragma unroll
    for(int u=0;u<4;u++){ const float* q=p+3l*idx[i+u*s]; v[u][0]=__ldg(q); v[u][1]=__ldg(q+1); v[u][2]=__ldg(q+2); }
    #pragma unroll
    for(int u=0;u<4;u++){ long o=3*(i+u*s); out[o]=v[u][0]; out[o+1]=v[u][1]; out[o+2]=v[u][2]; } }
  for(;i<n;i+=s){ const float* q=p+3l*idx[i]; out[3*i]=__ldg(q); out[3*i+1]=__ldg(q+1); out[3*i+2]=__ldg(q+2);} }
template<typename F> float timeit(F f,int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; } return best; }
int main(){
  const long n=134217728; unsigned *idx; float *p,*o;
  CK(cudaMalloc(&idx,n*4)); CK(cudaMalloc(&p,n*12)); CK(cudaMalloc(&o,n*12)); CK(cudaMemset(p,0,n*12));
  kIdx<<<(n+255)/256,256>>>(idx,n); CK(cudaDeviceSynchronize());
  size_t g=0; cudaDeviceGetLimit(&g,cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n",g);
  for(size_t gran: {128,64,32}){
    cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,gran); cudaDeviceGetLimit(&g,cudaLimitMaxL2FetchGranularity);
    float a=timeit([&]{kGather<<<148*16,256>>>(p,idx,n,o);}); float b=timeit([&]{kGather4<<<148*8,256>>>(p,idx,n,o);});
    printf("granularity req %zu (%s) got %zu: gather+write %7.3f ms | 4-deep %7.3f ms\n",gran,cudaGetErrorString(e),g,a,b);
  }
  return 0; }
