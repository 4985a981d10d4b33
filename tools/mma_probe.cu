// Micro-benchmark (development tool, not part of the product): cycles per tcgen05.mma as a function of the shared-memory
// operand layout (no swizzle vs 128-byte swizzle), operand start alignment, SBO and N. One CTA per SM, one thread issues
// a train of MMAs on zeroed shared memory; the train is timed with clock64 between the first issue and the commit's
// mbarrier completion. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define NSLOT 8
struct Probe {
    uint64_t a[NSLOT], b[NSLOT];   // descriptors with start address relative to the smem base
    uint32_t idesc[NSLOT];
    int nslot, iters;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Probe pr, long long *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) ((uint4 *)smem)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    // the issuing thread is elected inside a warp-uniform branch: a divergent `threadIdx.x == 0` branch makes nvcc wrap
    // every MMA in an ELECT/branch loop (~50 cycles of issue overhead per MMA, which is what the first version measured)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    uint32_t elected = 0;
    if (warp == 0) asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    if (warp == 0 && elected) {
        const uint64_t base = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
        uint64_t da[NSLOT], db[NSLOT];
        for (int j = 0; j < NSLOT; j++) { da[j] = pr.a[j] + base; db[j] = pr.b[j] + base; }
        for (int rep = 0; rep < 2; rep++) {   // rep 0 = warm-up
            const long long t0 = clock64();
            for (int it = 0; it < pr.iters; it++) {
#pragma unroll
                for (int j = 0; j < NSLOT; j++) {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                            "l"(da[j]), "l"(db[j]), "r"(pr.idesc[j]), "r"(1)
                            : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(b), "r"((uint32_t)rep) : "memory");
            } while (!done);
            const long long t1 = clock64();
            if (rep == 1) out[blockIdx.x] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static uint64_t desc(uint32_t off, uint32_t lbo, uint32_t sbo, int layout /*0 none, 2 sw128*/, int base_off = 0)
{
    return (uint64_t)((off >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
static uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

static long long *d_out;
static void run(const char *name, const Probe &p, double macs_per_iter)
{
    probe<<<148, 128, 201 * 1024>>>(p, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double med = (double)h[74] / p.iters;
    printf("%-58s clk/iter min %8.1f med %8.1f max %8.1f | MAC/clk med %7.1f (%4.1f%% of 4096)\n", name, (double)h[0] / p.iters, med,
           (double)h[147] / p.iters, macs_per_iter / med, 100.0 * macs_per_iter / med / 4096.0);
    fflush(stdout);
}

int main()
{
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    cudaMalloc(&d_out, 148 * sizeof(long long));
    const int A0 = 0, B0 = 100 * 1024;   // operand regions inside the 200 KB buffer
    char name[160];
    const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    // ---- 1. un-swizzled K-major, aligned and shifted A starts, SBO 128 / 192
    for (int sbo : {128, 192}) for (int shift : {0, 16, 64}) for (int n : Ns) {
        Probe p; memset(&p, 0, sizeof p); p.nslot = 8; p.iters = 400;
        for (int j = 0; j < 8; j++) {
            // A: 8 planes of 4 KB (LBO = 4096), per slot a different tap offset (multiples of 128 B plus `shift`)
            p.a[j] = desc(A0 + (j % 4) * 1152 * (sbo / 128) + shift + (j / 4) * 2 * 8192, 8192, sbo, 0);
            p.b[j] = desc(B0 + j * 256 * 16 * 2, 256 * 16, 128, 0);
            p.idesc[j] = idesc(128, n);
        }
        snprintf(name, sizeof name, "noswz  A sbo=%d shift=%-2d  M=128 N=%-3d", sbo, shift, n);
        run(name, p, 8.0 * 128 * n * 16);
    }
    // ---- 2. conv2 pattern today: pair (A_hi x N=64, A_lo x N=32), SBO 192, shifts cycling 0..4 px
    for (int aligned : {1, 0}) {
        Probe p; memset(&p, 0, sizeof p); p.nslot = 8; p.iters = 400;
        for (int j = 0; j < 8; j++) {
            const int tap = j / 2, sh = aligned ? 0 : (tap % 5) * 16;
            p.a[j] = desc(A0 + (j & 1) * 4 * 3840 + tap * 192 + sh, 3840, 192, 0);
            p.b[j] = desc(B0 + tap * 4096, 1024, 128, 0);
            p.idesc[j] = idesc(128, (j & 1) ? 32 : 64);
        }
        snprintf(name, sizeof name, "noswz  conv2 pairs (N=64,N=32) sbo=192 %s", aligned ? "aligned" : "tap-shifted");
        run(name, p, 4.0 * 128 * 96 * 16);
    }
    // ---- 3. 128-byte swizzle K-major (rows = 128 B, SBO = 1024), row-shifted A starts with base_offset
    for (int rshift : {0, 1, 3}) for (int n : Ns) {
        Probe p; memset(&p, 0, sizeof p); p.nslot = 8; p.iters = 400;
        for (int j = 0; j < 8; j++) {
            const uint32_t aoff = A0 + (j / 4) * 20480 + rshift * 128 * (1 + j % 4) + (j % 4) * 32;
            p.a[j] = desc(aoff, 16, 1024, 2, (aoff >> 7) & 7);
            p.b[j] = desc(B0 + (j / 4) * 32768 + (j % 4) * 32, 16, 1024, 2);
            p.idesc[j] = idesc(128, n);
        }
        snprintf(name, sizeof name, "sw128  A rowshift=%d (base_offset set)  M=128 N=%-3d", rshift, n);
        run(name, p, 8.0 * 128 * n * 16);
    }
    // ---- 4. sw128 A, row pitch folded into SBO (2-D tile: 8-pixel rows at pitch 12 rows -> SBO = 1536)
    for (int rshift : {0, 1}) for (int n : {64, 128, 256}) {
        Probe p; memset(&p, 0, sizeof p); p.nslot = 8; p.iters = 400;
        for (int j = 0; j < 8; j++) {
            const uint32_t aoff = A0 + rshift * 128 * (1 + j % 4) + (j % 4) * 32 + (j / 4) * 1536;
            p.a[j] = desc(aoff, 16, 1536, 2, (aoff >> 7) & 7);
            p.b[j] = desc(B0 + (j / 4) * 32768 + (j % 4) * 32, 16, 1024, 2);
            p.idesc[j] = idesc(128, n);
        }
        snprintf(name, sizeof name, "sw128  A sbo=1536 rowshift=%d  M=128 N=%-3d", rshift, n);
        run(name, p, 8.0 * 128 * n * 16);
    }
    // ---- 5. M=64
    for (int n : {64, 128, 256}) {
        Probe p; memset(&p, 0, sizeof p); p.nslot = 8; p.iters = 400;
        for (int j = 0; j < 8; j++) {
            p.a[j] = desc(A0 + (j / 4) * 20480 + (j % 4) * 32, 16, 1024, 2);
            p.b[j] = desc(B0 + (j / 4) * 32768 + (j % 4) * 32, 16, 1024, 2);
            p.idesc[j] = idesc(64, n);
        }
        snprintf(name, sizeof name, "sw128  M=64 N=%-3d", n);
        run(name, p, 8.0 * 64 * n * 16);
    }
    return 0;
}
