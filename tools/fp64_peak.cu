// FP64 peak microbenchmark for B200 (sm_100a): DFMA (CUDA-core) and DMMA (mma.sync m8n8k4 f64).
// The measured numbers are the FP64 roofline denominators quoted in DESIGN.md / bench.py
// (MEASURED_PEAKS.json carries no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < NACC; ++u) dmma(c[u][0], c[u][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d\n", p.name, sms);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int iters = 20000; float ms;
            dfma_kernel<<<sms * bps, threads>>>(out, 100);
            cudaDeviceSynchronize();
            cudaEventRecord(e0); dfma_kernel<<<sms * bps, threads>>>(out, iters); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 64 * iters * (double)threads * sms * bps;
            printf("DFMA threads %4d bps %d: %.2f TFLOP/s (%.2f ms)\n", threads, bps, fl / ms * 1e-9, ms);
            dmma_kernel<8><<<sms * bps, threads>>>(out, 100);
            cudaDeviceSynchronize();
            cudaEventRecord(e0); dmma_kernel<8><<<sms * bps, threads>>>(out, iters); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)(threads / 32) * sms * bps;
            printf("DMMA threads %4d bps %d: %.2f TFLOP/s (%.2f ms)\n", threads, bps, fl / ms * 1e-9, ms);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
