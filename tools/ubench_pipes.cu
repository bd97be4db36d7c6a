// Instruction-throughput microbenchmark for the pipes the ORB kernels lean on (sm_100a).
// Answers: ops/clk/SM of packed 16x2 integer min/max (VIMNMX/VIMNMX3), half2 min/max (HMNMX2), POPC, LOP3, PRMT,
// SHF, IMAD, and whether pairs of them overlap (separate pipes).  SURVEY.md §8(d) asks for the POPC rate on cc 10.0.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu ; run on one GPU.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

enum Op { OP_IADD3, OP_LOP3, OP_SHF, OP_PRMT, OP_IMAD, OP_VIMNMX32, OP_VIMNMX3_32, OP_VMNMX16X2, OP_VMNMX3_16X2,
          OP_HMNMX2, OP_HFMA2, OP_FMNMX, OP_POPC, OP_VABSDIFF4, OP_VIADD16X2, OP_FFMA,
          MIX_V3_H2, MIX_V3_IMAD, MIX_POPC_LOP3, MIX_POPC_IMAD, MIX_V3_PRMT, MIX_V3_SHF, MIX_H2_PRMT, MIX_LOP3_IMAD, OP_COUNT };
static const char *names[] = {"IADD3", "LOP3", "SHF", "PRMT", "IMAD", "VIMNMX.S32", "VIMNMX3.S32", "VIMNMX.U16x2",
                              "VIMNMX3.U16x2", "HMNMX2", "HFMA2", "FMNMX", "POPC", "VABSDIFF4", "VIADD.16x2", "FFMA",
                              "VIMNMX3.U16x2+HMNMX2", "VIMNMX3.U16x2+IMAD", "POPC+LOP3", "POPC+IMAD", "VIMNMX3.U16x2+PRMT",
                              "VIMNMX3.U16x2+SHF", "HMNMX2+PRMT", "LOP3+IMAD"};

template <int OP>
__device__ __forceinline__ void step(unsigned (&a)[NACC], unsigned b, unsigned c) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
        unsigned x = a[i];
        if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
        if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == OP_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == OP_VIMNMX32) asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(b + i));
        if (OP == OP_VIMNMX3_32) { x = __vimax3_s32(x, b + i, c); }
        if (OP == OP_VMNMX16X2) { x = __vmaxu2(x, b + i); }
        if (OP == OP_VMNMX3_16X2) { x = __vimax3_u16x2(x, b + i, c); }
        if (OP == OP_HMNMX2) { __half2 h = *reinterpret_cast<__half2 *>(&x); unsigned bb = b + i; h = __hmax2(h, *reinterpret_cast<__half2 *>(&bb)); x = *reinterpret_cast<unsigned *>(&h); }
        if (OP == OP_HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == OP_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+r"(x) : "r"(b + i));
        if (OP == OP_POPC) { unsigned t; asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(x)); x = t ^ b; }
        if (OP == OP_VABSDIFF4) { x = __vabsdiffu4(x, b + i); }
        if (OP == OP_VIADD16X2) { x = __vadd2(x, b); }
        if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == MIX_V3_H2) { if (i & 1) { __half2 h = *reinterpret_cast<__half2 *>(&x); unsigned bb = b + i; h = __hmax2(h, *reinterpret_cast<__half2 *>(&bb)); x = *reinterpret_cast<unsigned *>(&h); } else x = __vimax3_u16x2(x, b + i, c); }
        if (OP == MIX_V3_IMAD) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else x = __vimax3_u16x2(x, b + i, c); }
        if (OP == MIX_POPC_LOP3) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); else { unsigned t; asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(x)); x = t + b; } }
        if (OP == MIX_POPC_IMAD) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else { unsigned t; asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(x)); x = t + b; } }
        if (OP == MIX_V3_PRMT) { if (i & 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else x = __vimax3_u16x2(x, b + i, c); }
        if (OP == MIX_V3_SHF) { if (i & 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else x = __vimax3_u16x2(x, b + i, c); }
        if (OP == MIX_H2_PRMT) { if (i & 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else { __half2 h = *reinterpret_cast<__half2 *>(&x); unsigned bb = b + i; h = __hmax2(h, *reinterpret_cast<__half2 *>(&bb)); x = *reinterpret_cast<unsigned *>(&h); } }
        if (OP == MIX_LOP3_IMAD) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); }
        a[i] = x;
    }
}

template <int OP>
__global__ void __launch_bounds__(256) bench(unsigned *out, long long *cycles, unsigned b, unsigned c) {
    unsigned a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) step<OP>(a, b, c);
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory load throughput: 32-bit conflict-free loads
__global__ void __launch_bounds__(256) bench_lds(unsigned *out, long long *cycles) {
    __shared__ unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = i * 7u;
    __syncthreads();
    unsigned acc[4] = {0, 0, 0, 0};
    unsigned idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j & 3] += sm[(idx + j * 256 + it) & 4095];
    }
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, unsigned *d_out, long long *d_cyc) {
    const int blocks_per_sm = 4, nb = sms * blocks_per_sm;
    bench<OP><<<nb, 256>>>(d_out, d_cyc, 0x01010101u, 0x00030201u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<nb, 256>>>(d_out, d_cyc, 0x01010101u, 0x00030201u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = new long long[nb];
    cudaMemcpy(h, d_cyc, sizeof(long long) * nb, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += (double)h[i]; avg /= nb;
    delete[] h;
    double ops_per_sm = (double)blocks_per_sm * 256.0 * NACC * ITERS;
    printf("{\"op\": \"%s\", \"thread_ops_per_clk_per_sm\": %.1f, \"gops_per_s\": %.1f, \"ms\": %.4f}\n", names[OP],
           ops_per_sm / avg, (double)nb * 256.0 * NACC * ITERS / (ms * 1e6), ms);
}

template <int OP> struct Runner { static void go(int sms, unsigned *o, long long *c) { run<OP>(sms, o, c); Runner<OP + 1>::go(sms, o, c); } };
template <> struct Runner<OP_COUNT> { static void go(int, unsigned *, long long *) {} };

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    unsigned *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(unsigned) * sms * 4 * 256); cudaMalloc(&d_cyc, sizeof(long long) * sms * 4);
    Runner<0>::go(sms, d_out, d_cyc);
    {
        int nb = sms * 4;
        bench_lds<<<nb, 256>>>(d_out, d_cyc); cudaDeviceSynchronize();
        bench_lds<<<nb, 256>>>(d_out, d_cyc); cudaDeviceSynchronize();
        long long *h = new long long[nb];
        cudaMemcpy(h, d_cyc, sizeof(long long) * nb, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < nb; i++) avg += (double)h[i]; avg /= nb;
        printf("{\"op\": \"LDS.32\", \"thread_ops_per_clk_per_sm\": %.1f}\n", 4.0 * 256 * 8 * ITERS / avg);
        delete[] h;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
