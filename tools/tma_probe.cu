// Dev probe: which out-of-bounds TMA box placements are legal on sm_100a?  (tools/, not product)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, int c1, int c2, double *out, int boxrows, int do_store) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *base = raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(raw) & 1023u)) & 1023u);
    __shared__ unsigned long long bar;
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(base);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(boxrows * 128));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sd), "l"(&map), "r"(0), "r"(c1), "r"(c2), "r"(sb) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(sb) : "memory");
    double *s = (double *)base;
    for (int i = threadIdx.x; i < boxrows * 16; i += blockDim.x) out[i] = s[i];
    __syncthreads();
    if (do_store && threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map), "r"(0), "r"(c1), "r"(c2), "r"(sd) : "memory");
        asm volatile("cp.async.bulk.commit_group;");
        asm volatile("cp.async.bulk.wait_group 0;");
    }
}
int main(int argc, char **argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1; int idx = -1;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    struct Case { int chunks, batch, box, c1, c2, store, swz, promo; const char *name; } cases[] = {
        {300, 2, 256, 0, 1, 1, 1, 2, "in-bounds"},
        {300, 2, 256, 100, 1, 1, 1, 2, "tail beyond dim (box start 100, dim 300)"},
        {2, 2, 256, 0, 1, 1, 1, 2, "box 256 > dim 2, start 0"},
        {300, 2, 256, -10, 1, 1, 1, 2, "negative start -10"},
        {2, 2, 256, -254, 1, 1, 1, 2, "box>dim and negative start"},
        {2, 2, 256, -254, 1, 0, 1, 2, "same, load only"},
        {2, 2, 256, -254, 1, 1, 0, 2, "same, no swizzle"},
        {2, 2, 256, -254, 1, 1, 1, 0, "same, no L2 promotion"},
        {2, 2, 2, 0, 1, 1, 1, 2, "box == dim == 2"},
    };
    for (auto &c : cases) {
        ++idx; if (only >= 0 && idx != only) continue;
        int rowsPerSys = c.chunks * 8 + 3;
        size_t n = (size_t)rowsPerSys * c.batch * 2;
        double *g, *out; cudaMalloc(&g, n * 8 + 4096); cudaMalloc(&out, 256 * 128);
        double *h = (double *)malloc(n * 8); for (size_t i = 0; i < n; ++i) h[i] = (double)i;
        cudaMemcpy(g, h, n * 8, cudaMemcpyHostToDevice);
        CUtensorMap map; cuuint64_t dims[3] = {16, (cuuint64_t)c.chunks, (cuuint64_t)c.batch};
        cuuint64_t strides[2] = {128, (cuuint64_t)rowsPerSys * 16}; cuuint32_t box[3] = {16, (cuuint32_t)c.box, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
        probe<<<1, 128, 40000>>>(map, c.c1, c.c2, out, c.box, c.store);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%-45s encode=%d run=%s\n", c.name, (int)r, cudaGetErrorName(e));
        if (e != cudaSuccess) { printf("  (context poisoned, stopping)\n"); return 0; }
        free(h); cudaFree(g); cudaFree(out);
    }
    return 0;
}
