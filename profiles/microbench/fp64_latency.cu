#include <cstdio>
#include <cuda_runtime.h>
// dependent DFMA chain: latency; independent chains: throughput
template<int ILP> __global__ void k_dfma(double* out, int iters, long long* cyc) {
    double a[ILP]; for (int j=0;j<ILP;j++) a[j] = threadIdx.x*1e-9 + j;
    double b = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int i=0;i<iters;i++) {
        #pragma unroll
        for (int j=0;j<ILP;j++) a[j] = fma(a[j], b, c);
    }
    long long t1 = clock64();
    double s=0; for (int j=0;j<ILP;j++) s+=a[j];
    out[blockIdx.x*blockDim.x+threadIdx.x] = s;
    if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
__global__ void k_dmul_dadd(double* out, int iters, long long* cyc) {
    double a = threadIdx.x*1e-9+1.0, b = 1.0000001;
    long long t0 = clock64();
    for (int i=0;i<iters;i++) { a = a*b; a = a + 1e-9; }
    long long t1 = clock64();
    out[blockIdx.x*blockDim.x+threadIdx.x] = a;
    if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1<<24); cudaMalloc(&cyc, 8);
    long long h; int iters = 4096;
    // latency: 1 warp
    k_dfma<1><<<1,32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.2f cycles\n", (double)h/iters);
    k_dfma<2><<<1,32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP2 1 warp: %.2f cycles/iter (2 fma)\n", (double)h/iters);
    k_dfma<4><<<1,32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP4 1 warp: %.2f cycles/iter (4 fma)\n", (double)h/iters);
    k_dfma<8><<<1,32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP8 1 warp: %.2f cycles/iter (8 fma)\n", (double)h/iters);
    k_dmul_dadd<<<1,32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMUL+DADD dependent pair: %.2f cycles\n", (double)h/iters);
    // throughput: fill one SM with warps
    for (int warps : {4, 8, 16, 32, 64}) {
        k_dfma<4><<<1, 32*warps>>>(out, iters, cyc); if (warps>32) { } cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (warps <= 32) printf("1 SM, %d warps, ILP4: %.2f cycles/iter -> %.1f DFMA lane-ops/cycle/SM\n", warps, (double)h/iters, 4.0*32*warps/((double)h/iters));
    }
    // whole chip throughput
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma<8><<<148*8, 256>>>(out, 1<<14, cyc);
    cudaEventRecord(e0); k_dfma<8><<<148*8, 256>>>(out, 1<<14, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0*8*(1<<14)*148*8*256;
    printf("chip DFMA: %.2f TFLOP/s (%.3f ms)\n", flops/ms/1e9, ms);
    return 0;
}
