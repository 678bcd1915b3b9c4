// Micro-probe of the fp64 path on sm_100a: dependent-issue latency of DFMA / DMUL / MUFU.RCP64H, DFMA throughput
// against the number of resident warps, LDS.128 and bar.sync latency.  Build and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_probe tools/fp64_probe.cu && /tmp/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dep_dfma(double *out, long long *clk, double a, double b, int iters) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x = fma(x, b, a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
}

__global__ void dep_rcp(double *out, long long *clk, double a, int iters) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
}

template <int ILP>
__global__ void thr_dfma(double *out, long long *clk, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
}

__global__ void dep_lds(double *out, long long *clk, int iters) {
    __shared__ double2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2((double)((i * 7 + 1) & 1023), 0.0);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) idx = (int)buf[idx].x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    out[threadIdx.x] = idx;
}

__global__ void bar_lat(long long *clk, int iters) {
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
    double *out;
    long long *clk, h;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&clk, 8);
    const int iters = 256;
    dep_dfma<<<1, 32>>>(out, clk, 1.0, 0.999, iters);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency (1 warp): %.2f clk\n", (double)h / (iters * 32));
    dep_rcp<<<1, 32>>>(out, clk, 1.3, iters);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("dependent MUFU.RCP64H latency (1 warp): %.2f clk\n", (double)h / (iters * 32));
    dep_lds<<<1, 32>>>(out, clk, iters);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("dependent LDS.128 + cvt latency (1 warp): %.2f clk\n", (double)h / (iters * 16));
    for (int th : {64, 128, 256, 512, 1024}) {
        bar_lat<<<1, th>>>(clk, iters);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("bar.sync, %4d threads: %.2f clk\n", th, (double)h / (iters * 16));
    }
    for (int th : {32, 64, 128, 256, 512, 1024}) {
        thr_dfma<1><<<1, th>>>(out, clk, 1.0, 0.999, iters);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        double c1 = (double)h;
        thr_dfma<4><<<1, th>>>(out, clk, 1.0, 0.999, iters);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        double c4 = (double)h;
        thr_dfma<8><<<1, th>>>(out, clk, 1.0, 0.999, iters);
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        double c8 = (double)h;
        printf("DFMA/clk/SM at %4d threads: ILP1 %.1f  ILP4 %.1f  ILP8 %.1f\n", th, th * iters * 8.0 / c1,
               th * iters * 8.0 * 4 / c4, th * iters * 8.0 * 8 / c8);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
