// Micro-benchmarks that set the design numbers for the B200 build:
//   DMMA.8x8x4 issue rate, DFMA rate, REDG.F64 throughput, cuBLAS DGEMM (the Hankel bar), HBM copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb mb.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__global__ void dmma_rate(double* out, int iters){
  double c[8][2]; for(int i=0;i<8;i++){c[i][0]=0;c[i][1]=0;}
  double a=threadIdx.x*1e-3, b=threadIdx.x*2e-3;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<8;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0; for(int i=0;i<8;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void dfma_rate(double* out, int iters){
  double c[16]; for(int i=0;i<16;i++) c[i]=i;
  double a=1.0000001, b=threadIdx.x*1e-9;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<16;i++) c[i]=fma(c[i],a,b);
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// each thread does n atomics to distinct addresses (stride pattern: coalesced across warp)
__global__ void red_rate(double* g, size_t n_addr, int per_thread){
  size_t t = blockIdx.x*(size_t)blockDim.x+threadIdx.x;
  for(int i=0;i<per_thread;i++){ size_t a=(t + (size_t)i*gridDim.x*blockDim.x) % n_addr; atomicAdd(&g[a], 1.0); }
}
__global__ void copyk(const double2* __restrict__ a, double2* __restrict__ b, size_t n){
  for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; i<n; i+=(size_t)gridDim.x*blockDim.x) b[i]=a[i];
}
template<class F> float timeit(F f, int reps=5){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize()); float best=1e30;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  for(int wps: {4,8,16,32}){ // warps per SM
    int iters=20000; int thr=wps*32;
    float ms=timeit([&]{dmma_rate<<<sms,thr>>>(out,iters);});
    double flops=(double)sms*wps*iters*8*512; printf("DMMA  warps/SM=%2d : %.2f TFLOP/s\n", wps, flops/ms*1e-9);
    ms=timeit([&]{dfma_rate<<<sms,thr>>>(out,iters);});
    flops=(double)sms*thr*(double)iters*16*2; printf("DFMA  warps/SM=%2d : %.2f TFLOP/s\n", wps, flops/ms*1e-9);
  }
  // RED.F64
  for(size_t naddr: {(size_t)1<<20, (size_t)1<<24}){
    double* g; CK(cudaMalloc(&g, naddr*8)); CK(cudaMemset(g,0,naddr*8));
    int blocks=sms*16, thr=256, per=64;
    float ms=timeit([&]{red_rate<<<blocks,thr>>>(g,naddr,per);});
    printf("REDG.F64 over %zu addrs: %.1f G atomics/s\n", naddr, (double)blocks*thr*per/ms*1e-6);
    cudaFree(g);
  }
  // HBM copy
  { size_t n=(size_t)1<<27; double2 *a,*b; CK(cudaMalloc(&a,n*16)); CK(cudaMalloc(&b,n*16)); CK(cudaMemset(a,1,n*16));
    float ms=timeit([&]{copyk<<<sms*8,512>>>(a,b,n);});
    printf("copy kernel 2x%zu MB: %.1f GB/s\n", n*16>>20, 2.0*n*16/ms*1e-6);
    ms=timeit([&]{cudaMemcpyAsync(b,a,n*16,cudaMemcpyDeviceToDevice);});
    printf("cudaMemcpy D2D     : %.1f GB/s\n", 2.0*n*16/ms*1e-6);
    cudaFree(a); cudaFree(b);}
  // cuBLAS DGEMM: the Hankel shapes (row-major out[2Nz,Nr]=in[2Nz,Nr]@M[Nr,Nr] == col-major C[Nr,2Nz]=M'[Nr,Nr] in[Nr,2Nz])
  cublasHandle_t h; cublasCreate(&h);
  int shapes[][3]={{256,8192,256},{256,49152,256},{512,4096,512},{512,32768,512},{4096,4096,4096}};
  for(auto& s: shapes){ int m=s[0],n=s[1],k=s[2]; double *A,*B,*C; CK(cudaMalloc(&A,(size_t)m*k*8)); CK(cudaMalloc(&B,(size_t)k*n*8)); CK(cudaMalloc(&C,(size_t)m*n*8));
    CK(cudaMemset(A,0,(size_t)m*k*8)); CK(cudaMemset(B,0,(size_t)k*n*8)); double al=1,be=0;
    float ms=timeit([&]{cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,m,n,k,&al,A,m,B,k,&be,C,m);},10);
    printf("cuBLAS DGEMM m=%d n=%d k=%d : %.3f ms  %.2f TFLOP/s\n", m,n,k,ms, 2.0*m*n*k/ms*1e-9);
    cudaFree(A);cudaFree(B);cudaFree(C);}
  return 0;
}
