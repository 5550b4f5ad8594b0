// Micro-probe: throughput of mma.sync.aligned.m8n8k4 f64 (DMMA) against register-resident DFMA chains on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_kernel(long iters, double *out) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c[8][2];
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    for (long it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(long iters, double *out) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    double c[16];
    for (int i = 0; i < 16; i++) c[i] = i;
    for (long it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 256;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int which = 0; which < 2; which++) {
        const long iters = 1 << 15;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (which == 0) dmma_kernel<<<blocks, threads>>>(iters, out);
            else dfma_kernel<<<blocks, threads>>>(iters, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            // DMMA m8n8k4: 8*8*4 = 256 FMA per warp instruction; DFMA: 32 FMA per warp instruction
            const double warps = (double)blocks * threads / 32;
            const double fma = which == 0 ? warps * iters * 8 * 256.0 : warps * iters * 16 * 32.0;
            if (rep == 2) printf("%s: %.2f TFLOP/s (%.3f ms)\n", which == 0 ? "DMMA m8n8k4" : "DFMA", 2 * fma / (ms * 1e-3) / 1e12, ms);
        }
    }
    return 0;
}
