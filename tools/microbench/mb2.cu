// DMMA.8x8x4 throughput with DISTINCT operands (the first microbench, mb.cu, reused one register pair for every
// instruction): accumulator tiles of MI x NI MMA blocks per warp, operands from registers or re-read from shared
// memory every k-step, as the Hankel kernels do.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb2 mb2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// MODE 0: operands in registers (rotated so every instruction names different registers)
// MODE 1: operands re-loaded from shared memory each k-step (LDS.64 each, conflict-free)
// MODE 2: as 1 but A loaded as LDS.128 pairs (re, im planes) like the Hankel kernel
template <int MI, int NI, int MODE>
__global__ void k(double *out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[MI][NI][2];
    for (int i = 0; i < MI; i++) for (int n = 0; n < NI; n++) acc[i][n][0] = acc[i][n][1] = 0.;
    double a[MI], b[NI];
    for (int i = 0; i < MI; i++) a[i] = 1e-3 * (lane + i);
    for (int n = 0; n < NI; n++) b[n] = 2e-3 * (lane + n);
    const double *base = sm + (warp & 3) * 1024;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < MI; i++) a[i] = base[((it & 3) * MI + i) * 32 + lane];
#pragma unroll
            for (int n = 0; n < NI; n++) b[n] = base[512 + ((it & 3) * NI + n) * 32 + lane];
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < MI; i += 2) {
                double2 v = *(const double2 *)(base + (((it & 3) * MI + i) * 32 + 2 * lane) % 1024);
                a[i] = v.x; a[i + 1] = v.y;
            }
#pragma unroll
            for (int n = 0; n < NI; n++) b[n] = base[512 + ((it & 3) * NI + n) * 32 + lane];
        }
#pragma unroll
        for (int n = 0; n < NI; n++)
#pragma unroll
            for (int i = 0; i < MI; i++) dmma(acc[i][n][0], acc[i][n][1], a[i], b[n]);
        if (MODE == 0) { double t = a[0]; for (int i = 0; i + 1 < MI; i++) a[i] = a[i + 1]; a[MI - 1] = t; }
    }
    double s = 0;
    for (int i = 0; i < MI; i++) for (int n = 0; n < NI; n++) s += acc[i][n][0] + acc[i][n][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MI, int NI, int MODE>
void run(double *out, int sms, int wps) {
    int iters = 4000, thr = wps * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(k<MI, NI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k<MI, NI, MODE><<<sms, thr, 65536>>>(out, 10);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<MI, NI, MODE><<<sms, thr, 65536>>>(out, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)sms * wps * iters * MI * NI * 512.;
    printf("DMMA tile %dx%d mode %d warps/SM=%2d : %.2f TFLOP/s\n", MI, NI, MODE, wps, flops / ms * 1e-9);
}
int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev)); int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    for (int wps : {4, 8, 16}) {
        run<4, 4, 0>(out, sms, wps); run<4, 4, 1>(out, sms, wps); run<4, 4, 2>(out, sms, wps);
        run<8, 4, 0>(out, sms, wps); run<8, 4, 1>(out, sms, wps); run<8, 4, 2>(out, sms, wps);
        run<2, 4, 1>(out, sms, wps); run<2, 8, 1>(out, sms, wps);
    }
    return 0;
}
