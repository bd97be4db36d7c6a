// Issue-mix microbenchmark for the pair-plane FAST kernel (orbx_fast2.cu): how many thread-instructions per clock an SM retires
// for the instruction kinds of its scoring loop, alone and mixed, at the warp count the kernel runs with (16 warps / SM).
//   V2   VIMNMX.U16x2  (2 registers)        V3   VIMNMX3.U16x2 (3 registers)
//   HF   HFMA2 with an immediate multiplier (fma.rn.f16x2 x, -1, y)     HR   HFMA2.RELU, same form
//   IM   IMAD (3 registers)                 LD   LDS.32 at an immediate offset (conflict-free)
// and the mixes the kernel can choose between: V2+HF 1:1, V3+HF 1:1, V2+V3+HF+HR+LD in the loop's proportions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mix ubench_mix.cu ; prints one JSON line per mode.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 12

enum Mode { M_V2, M_V3, M_HF, M_HR, M_IM, M_LD, M_V2_HF, M_V3_HF, M_V3_IM, M_V2_V3, M_LOOP, M_LOOP_ALU, M_COUNT };
static const char *names[] = {"VIMNMX.U16x2", "VIMNMX3.U16x2", "HFMA2(imm)", "HFMA2.RELU(imm)", "IMAD", "LDS.32", "VIMNMX+HFMA2 1:1",
                              "VIMNMX3+HFMA2 1:1", "VIMNMX3+IMAD 1:1", "VIMNMX+VIMNMX3 1:1",
                              "loop mix: 17 LDS + 24 V2 + 22 V3 + 26 HFMA2 (8 pairs by HFMA2)", "loop mix: 17 LDS + 40 V2 + 22 V3 + 2 HFMA2 (all pairs by VIMNMX)"};

__device__ __forceinline__ uint32_t v2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint32_t v3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t hf(uint32_t a, uint32_t b) { uint32_t d; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0xBC00BC00u), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hr(uint32_t a, uint32_t b) { uint32_t d; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0xBC00BC00u), "r"(b)); return d; }
__device__ __forceinline__ uint32_t im(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t ld(uint32_t addr, int i) { uint32_t d; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(addr + (uint32_t)i)); return d; }

template <int MODE>
__global__ void __launch_bounds__(128) bench(uint32_t *out, long long *cycles, uint32_t b, uint32_t c, long long *ops_out) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (i * 7u) & 0x00FF00FFu;
    uint32_t a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = ((threadIdx.x * 2654435761u + i * 40503u + blockIdx.x) & 0x00FF00FFu);
    __syncthreads();
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 4u;
    long long ops = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == M_LOOP || MODE == M_LOOP_ALU) {
            // the proportions of one scoring item: 17 loads, then the min/max network
            uint32_t r[17];
            const uint32_t sa = saddr + ((it & 7) << 9);
#define LD17(k) asm volatile("ld.shared.u32 %0, [%1 + %2];" : "=r"(r[k]) : "r"(sa), "n"((k) * 516));
            LD17(0) LD17(1) LD17(2) LD17(3) LD17(4) LD17(5) LD17(6) LD17(7) LD17(8) LD17(9) LD17(10) LD17(11) LD17(12) LD17(13) LD17(14) LD17(15) LD17(16)
#undef LD17
            uint32_t x[16];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == M_LOOP) { const uint32_t t = hr(r[2 * k], r[2 * k + 1]); x[2 * k] = hf(t, r[2 * k]); x[2 * k + 1] = hf(t, r[2 * k + 1]); }
                else { x[2 * k] = v2(r[2 * k], r[2 * k + 1]); x[2 * k + 1] = __vminu2(r[2 * k], r[2 * k + 1]); }
            }
            uint32_t y[16];
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = v2(x[k], x[(k + 2) & 15]);
#pragma unroll
            for (int k = 0; k < 8; k++) { y[k] = v3(y[k], y[k + 8], v2(r[k], r[k + 9])); y[k + 8] = v3(y[k + 8], x[k], __vminu2(r[k + 1], r[k + 8])); }
            uint32_t z = v3(v3(y[0], y[1], y[2]), v3(y[3], y[4], y[5]), v3(y[6], y[7], v3(v3(y[8], y[9], y[10]), v3(y[11], y[12], y[13]), v2(y[14], y[15]))));
            a[0] = v2(a[0], hr(z, r[16]));
            ops += 1;
        } else {
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                uint32_t x = a[i];
                if (MODE == M_V2) x = v2(x, b + i);
                if (MODE == M_V3) x = v3(x, b + i, c);
                if (MODE == M_HF) x = hf(x, b);
                if (MODE == M_HR) x = hr(x, b);
                if (MODE == M_IM) x = im(x, b, c);
                if (MODE == M_LD) x ^= ld(saddr, (i * 512 + (it & 3) * 128) & 16383);
                if (MODE == M_V2_HF) x = (i & 1) ? hf(x, b) : v2(x, b + i);
                if (MODE == M_V3_HF) x = (i & 1) ? hf(x, b) : v3(x, b + i, c);
                if (MODE == M_V3_IM) x = (i & 1) ? im(x, b, c) : v3(x, b + i, c);
                if (MODE == M_V2_V3) x = (i & 1) ? v2(x, b + i) : v3(x, b + i, c);
                a[i] = x;
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)ops;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(int sms, uint32_t *d_out, long long *d_cyc) {
    const int blocks_per_sm = 4, nb = sms * blocks_per_sm;     // 4 x 128 threads = 16 warps / SM
    bench<MODE><<<nb, 128>>>(d_out, d_cyc, 0x00010001u, 0x00030002u, nullptr);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<nb, 128>>>(d_out, d_cyc, 0x00010001u, 0x00030002u, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = new long long[nb];
    cudaMemcpy(h, d_cyc, sizeof(long long) * nb, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += (double)h[i]; avg /= nb;
    delete[] h;
    // instructions per thread per iteration of the timed loop
    const double per_it = (MODE == M_LOOP) ? (17 + 24 + 16 + 22 + 1 + 1 + 1) : (MODE == M_LOOP_ALU) ? (17 + 16 + 16 + 16 + 22 + 1 + 1 + 1) : NACC;
    const double ops_per_sm = (double)blocks_per_sm * 128.0 * per_it * ITERS;
    printf("{\"mode\": \"%s\", \"thread_instr_per_clk_per_sm\": %.1f, \"clk_per_item_per_sm\": %.3f, \"ginstr_per_s\": %.1f, \"ms\": %.4f, \"sm_clocks\": %.0f}\n",
           names[MODE], ops_per_sm / avg, avg / ((double)blocks_per_sm * 128.0 * ITERS), (double)nb * 128.0 * per_it * ITERS / (ms * 1e6), ms, avg);
}

template <int M> struct Runner { static void go(int sms, uint32_t *o, long long *c) { run<M>(sms, o, c); Runner<M + 1>::go(sms, o, c); } };
template <> struct Runner<M_COUNT> { static void go(int, uint32_t *, long long *) {} };

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"warps_per_sm\": 16}\n", p.name, sms);
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(uint32_t) * sms * 4 * 128); cudaMalloc(&d_cyc, sizeof(long long) * sms * 4);
    Runner<0>::go(sms, d_out, d_cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
