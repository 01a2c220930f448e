// Issue-rate microbenchmark for the instruction forms the sweep kernels are made of (sm_100a).
// One CTA of 1024 threads per SM (8 warps per SM sub-partition); every test is an unrolled loop of 8
// independent dependency chains per thread; result = cycles per warp instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

struct K { uint32_t k[8]; };
constexpr int ITER = 2048;

#define CHAINS 8
#define BODY(NAME, NINST, ...)                                                                        \
    __global__ void __launch_bounds__(1024) NAME(uint32_t *out, long long *cyc, const __grid_constant__ K kk) \
    {                                                                                                 \
        uint32_t x[CHAINS], y[CHAINS], z[CHAINS];                                                                \
        for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 7 + c; y[c] = threadIdx.x ^ (c * 77); z[c] = threadIdx.x * 3 + c; } \
        __syncthreads();                                                                              \
        long long t0 = clock64();                                                                     \
        for (int it = 0; it < ITER; ++it) {                                                           \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { __VA_ARGS__ }                               \
        }                                                                                             \
        long long t1 = clock64();                                                                     \
        uint32_t s = 0;                                                                               \
        for (int c = 0; c < CHAINS; ++c) s += x[c] ^ y[c] ^ z[c];                                            \
        out[blockIdx.x * 1024 + threadIdx.x] = s;                                                     \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                              \
    }                                                                                                 \
    static const int NAME##_n = NINST;

// LOP3, three registers
BODY(lop3_rrr, 1, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(y[(c + 1) % CHAINS]));)
// LOP3, two registers + kernel parameter (uniform register / constant bank)
BODY(lop3_rru, 1, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(kk.k[c & 7]));)
// LOP3, two registers + immediate
BODY(lop3_rri, 1, asm volatile("lop3.b32 %0, %0, %1, 0x01010101, 0xf8;" : "+r"(x[c]) : "r"(y[c]));)
// XOR of two registers
BODY(xor_rr, 1, asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c]));)
// IMAD.WIDE by an immediate (Philox multiply), low and high both consumed
BODY(imad_wide, 1, { uint32_t lo, hi; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[c])); x[c] = lo; y[c] = hi; })
// IMAD by a uniform multiplier with accumulate
BODY(imad_u, 1, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(kk.k[c & 7]), "r"(y[c]));)
// one Philox round: 2 IMAD.WIDE + 2 LOP3 (key from the parameter bank)
BODY(philox_round, 2, { uint32_t lo, hi; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[c])); uint32_t n; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n) : "r"(hi), "r"(y[c]), "r"(kk.k[c & 7])); x[c] = n; y[c] = lo; })
// same with the key in a register
BODY(philox_round_r, 2, { uint32_t lo, hi; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[c])); uint32_t n; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n) : "r"(hi), "r"(y[c]), "r"(y[(c + 3) % CHAINS])); x[c] = n; y[c] = lo; })
// carry compare + IMAD.X Horner step
BODY(carry_horner, 2, asm volatile("{ .reg .u32 t; add.cc.u32 t, %0, %1; madc.lo.u32 %0, %0, %2, 0; }" : "+r"(x[c]) : "r"(y[c]), "r"(kk.k[1]));)
// PRMT byte extract
BODY(prmt, 1, asm volatile("prmt.b32 %0, %0, %1, 0x4441;" : "+r"(x[c]) : "r"(y[c]));)
// alternating LOP3 / IMAD (independent pipes)
BODY(lop3_imad_mix, 2, { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(kk.k[c & 7])); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(kk.k[(c + 1) & 7]), "r"(y[(c + 1) % CHAINS])); })

// --- how long does IMAD.WIDE hold its pipe?  (pair it with 1, 2, 3 ALU instructions and with an IMAD)
#define WIDE_STEP { uint32_t lo, hi; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[c])); uint32_t n; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n) : "r"(hi), "r"(y[c]), "r"(kk.k[c & 7])); x[c] = n; y[c] = lo; }
#define EXTRA_LOP3 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[c]) : "r"(z[(c + 1) % CHAINS]), "r"(kk.k[(c + 2) & 7]));
#define EXTRA_IMAD asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[c]) : "r"(kk.k[(c + 1) & 7]), "r"(z[(c + 1) % CHAINS]));
BODY(wide_lop3x2, 3, WIDE_STEP EXTRA_LOP3)
BODY(wide_lop3x3, 4, WIDE_STEP EXTRA_LOP3 EXTRA_LOP3)
BODY(wide_lop3x4, 5, WIDE_STEP EXTRA_LOP3 EXTRA_LOP3 EXTRA_LOP3)
BODY(wide_lop3_imad, 3, WIDE_STEP EXTRA_IMAD)
BODY(wide_lop3_imadx2, 4, WIDE_STEP EXTRA_IMAD EXTRA_IMAD)
// mul.hi alone + lop3
BODY(mulhi_lop3, 2, { uint32_t hi; asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(x[c])); uint32_t n; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n) : "r"(hi), "r"(y[c]), "r"(kk.k[c & 7])); x[c] = n; })
// mul.lo alone + lop3
BODY(mullo_lop3, 2, { uint32_t lo; asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(x[c])); uint32_t n; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n) : "r"(lo), "r"(y[c]), "r"(kk.k[c & 7])); x[c] = n; })
// IMAD.WIDE with the multiplier in a register instead of an immediate
BODY(wide_reg_lop3, 2, { uint32_t lo, hi; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(x[c]), "r"(kk.k[0])); uint32_t n; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(n) : "r"(hi), "r"(y[c]), "r"(kk.k[c & 7])); x[c] = n; y[c] = lo; })

template <typename F>
static void run(const char *name, F kern, int ninst)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * 1024 * sms);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    K kk;
    for (int i = 0; i < 8; ++i) kk.k[i] = 0x9E3779B9u * (i + 1) | 1u;
    kern<<<sms, 1024>>>(out, cyc, kk);
    kern<<<sms, 1024>>>(out, cyc, kk);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double winst_per_smsp = (double)ITER * CHAINS * ninst * 8.0; // 8 warps per sub-partition
    printf("%-16s %7.3f cycles per warp instruction per SMSP  (%s)\n", name, avg / winst_per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
#define RUN(N) run(#N, N, N##_n)
    RUN(lop3_rrr); RUN(lop3_rru); RUN(lop3_rri); RUN(xor_rr); RUN(imad_wide); RUN(imad_u);
    RUN(philox_round); RUN(philox_round_r); RUN(carry_horner); RUN(prmt); RUN(lop3_imad_mix);
    RUN(wide_lop3x2); RUN(wide_lop3x3); RUN(wide_lop3x4); RUN(wide_lop3_imad); RUN(wide_lop3_imadx2); RUN(mulhi_lop3); RUN(mullo_lop3); RUN(wide_reg_lop3);
    return 0;
}
