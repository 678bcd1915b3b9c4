// Dev probe (tools/, not product): is tensor memory usable as a parking space for registers next to warp roles with
// re-balanced register budgets?  Measures, per SM and with 8 warps moving 100 columns per thread at once:
// tcgen05.st throughput, tcgen05.ld throughput, round trip with verification; setmaxnreg inc/dec on a 384-thread CTA;
// bar.arrive / bar.sync hand-over between warp groups.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tools/tmem_probe.cu && ./tmem_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../fem_helmholtz_eqn_b200/csrc/fh_tmem.cuh"
using namespace fh;

constexpr int kCols = 100;  // columns parked per thread and system
__device__ __forceinline__ void park(uint32_t ta, const uint32_t *v) {
    tmem_st32(ta, v), tmem_st32(ta + 32, v + 32), tmem_st32(ta + 64, v + 64), tmem_st4(ta + 96, v + 96);
}
__device__ __forceinline__ void unpark(uint32_t ta, uint32_t *v) {
    tmem_ld32(ta, v), tmem_ld32(ta + 32, v + 32), tmem_ld32(ta + 64, v + 64), tmem_ld4(ta + 96, v + 96);
}

__global__ void __launch_bounds__(384, 1) probe(long long *out, int reps) {
    __shared__ uint32_t tbase;
    __shared__ volatile int mailbox;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    if (w == 0) tmem_alloc(&tbase, 512);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    long long res[6] = {0, 0, 0, 0, 0, 0};
    if (w < 8) {
        reg_inc<208>();
        // lane quadrant of this warp; warps w and w + 4 share a quadrant and split the columns
        const uint32_t ta = base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(w >> 2) * 256u;
        uint32_t v[kCols], r[kCols];
#pragma unroll
        for (int i = 0; i < kCols; ++i) v[i] = (uint32_t)(t * 1000 + i) * 2654435761u;
        asm volatile("bar.sync 1, 256;");
        long long c0 = clock64();
        for (int k = 0; k < reps; ++k) {
            park(ta + (k & 1) * 128, v);
#pragma unroll
            for (int i = 0; i < kCols; ++i) v[i] += 1;   // (keeps the stores from being hoisted / merged)
        }
        tmem_wait_st();
        asm volatile("bar.sync 1, 256;");
        long long c1 = clock64();
        uint32_t acc = 0;
        for (int k = 0; k < reps; ++k) {
            unpark(ta + (k & 1) * 128, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < kCols; ++i) acc += r[i];
        }
        asm volatile("bar.sync 1, 256;");
        long long c2 = clock64();
        // ld issued early, waited for late (what the back-substitution would do)
        long long bad = 0;
        for (int k = 0; k < reps; ++k) {
#pragma unroll
            for (int i = 0; i < kCols; ++i) v[i] = (uint32_t)(t * 1000 + i + k) * 2654435761u;
            park(ta + (k & 1) * 128, v);
            tmem_wait_st();
            unpark(ta + (k & 1) * 128, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < kCols; ++i) bad += (r[i] != v[i]);
        }
        asm volatile("bar.sync 1, 256;");
        long long c3 = clock64();
        // alignment: do xN moves need N-aligned columns?  (x8 at 4 mod 8, x16 at 12 mod 16, x4 at 20)
        long long bad_al = 0;
        {
#pragma unroll
            for (int i = 0; i < kCols; ++i) v[i] = (uint32_t)(t * 77 + i) * 2246822519u;
            const uint32_t tb = ta + 128;
            tmem_st8(tb + 4, v), tmem_st16(tb + 12, v + 8), tmem_st4(tb + 28, v + 24), tmem_st32(tb + 36, v + 28);
            tmem_wait_st();
            tmem_ld4(tb + 4, r), tmem_ld8(tb + 8, r + 4), tmem_ld16(tb + 16, r + 12), tmem_ld32(tb + 32, r + 28), tmem_ld8(tb + 60, r + 56), tmem_ld4(tb + 64, r + 60);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 56; ++i) bad_al += (r[i] != v[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) bad_al += (r[56 + i] != v[56 + i]) + (r[60 + i] != v[60 + i]);
        }
        atomicAdd((unsigned long long *)&out[blockIdx.x * 8 + 5], (unsigned long long)bad_al);
        // hand-over with the other warp group: this group produces, the other consumes
        for (int k = 0; k < 64; ++k) {
            asm volatile("bar.sync 4, 384;");          // mailbox consumed
            if (t == 0) mailbox = k;
            asm volatile("bar.arrive 3, 384;");        // mailbox full
        }
        res[0] = c1 - c0, res[1] = c2 - c1, res[2] = c3 - c2, res[3] = bad, res[4] = acc;
        if (lane == 0 && w == 0) {
            out[blockIdx.x * 8 + 0] = res[0], out[blockIdx.x * 8 + 1] = res[1], out[blockIdx.x * 8 + 2] = res[2];
        }
        atomicAdd((unsigned long long *)&out[blockIdx.x * 8 + 3], (unsigned long long)bad);
        if (acc == 12345) out[blockIdx.x * 8 + 7] = acc;
    } else {
        reg_dec<88>();
        long long bad = 0;
        for (int k = 0; k < 64; ++k) {
            asm volatile("bar.arrive 4, 384;");
            asm volatile("bar.sync 3, 384;");
            bad += (mailbox != k);
        }
        atomicAdd((unsigned long long *)&out[blockIdx.x * 8 + 4], (unsigned long long)bad);
    }
    __syncthreads();
    if (w == 0) tmem_dealloc(base, 512);
}

int main() {
    long long *out, h[148 * 8];
    cudaMalloc(&out, sizeof(h));
    const int reps = 64;
    for (int pass = 0; pass < 2; ++pass) {
        cudaMemset(out, 0, sizeof(h));
        probe<<<148, 384>>>(out, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = 256.0 * kCols * 4;
    for (int b = 0; b < 148; b += 49)
        printf("SM-block %3d: st %.0f cyc/park (%.1f B/cyc)  ld %.0f cyc/unpark (%.1f B/cyc)  round trip %.0f cyc  "
               "mismatches %lld  hand-over errors %lld\n", b, (double)h[b * 8] / reps, bytes * reps / h[b * 8],
               (double)h[b * 8 + 1] / reps, bytes * reps / h[b * 8 + 1], (double)h[b * 8 + 2] / reps, h[b * 8 + 3], h[b * 8 + 4]);
    long long bad = 0, bad_al = 0;
    for (int b = 0; b < 148; ++b) bad += h[b * 8 + 3] + h[b * 8 + 4], bad_al += h[b * 8 + 5];
    printf("unaligned xN moves (x8 at 4 mod 8, x16 at 12 mod 16, x32 at 36): %lld mismatches\n", bad_al);
    printf("total errors over 148 CTAs: %lld\n", bad);
    return bad != 0;
}
