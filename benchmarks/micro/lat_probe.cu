// Dependent-chain latencies of the warp primitives and fp64 routines the replay kernel leans on (cycles per op,
// one warp).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define FULL 0xFFFFFFFFu
template <typename F>
__device__ long long chain(F f, int n, uint32_t &x)
{
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = f(x);
    return clock64() - t0;
}

__global__ void probe(long long *out, uint32_t seed, double dseed)
{
    const int lane = threadIdx.x;
    uint32_t x = seed + lane * 2654435761u;
    __shared__ uint32_t sm[1024];
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 7 + 1) & 1023;
    __syncwarp();
    const int n = 256;
    long long t[12];
    t[0] = chain([&](uint32_t v) { return __match_any_sync(FULL, v & 0xFFFF) + v; }, n, x);
    t[1] = chain([&](uint32_t v) { return __reduce_or_sync(FULL, v) + lane; }, n, x);
    t[2] = chain([&](uint32_t v) { return __ballot_sync(FULL, v & 1) + v; }, n, x);
    t[3] = chain([&](uint32_t v) { return (uint32_t)((int)(v >> 1) % (int)((v & 0xFFF) + 5)) + v; }, n, x);
    t[4] = chain([&](uint32_t v) { return (v >> 1) % ((v & 0xFFF) + 5) + v; }, n, x);
    t[5] = chain([&](uint32_t v) { return sm[v & 1023]; }, n, x);
    t[6] = chain([&](uint32_t v) { return __shfl_up_sync(FULL, v, 3) + 1; }, n, x);
    double d = dseed + lane;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) d = __ddiv_rn(d, 2147483647.0) + 3.0;
    t[7] = clock64() - t0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) d = exp(-d * 1e-3) + 2.0;
    t[8] = clock64() - t0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) d = __dadd_rn(__dmul_rn(d, 0.999), 1.0);
    t[9] = clock64() - t0;
    float fl = (float)d;
    t0 = clock64();
    for (int i = 0; i < n; ++i) fl = __expf(-fl * 1e-3f) + 2.0f;
    t[10] = clock64() - t0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { __syncwarp(); x += sm[(x + lane) & 1023]; __syncwarp(); sm[(x * 3 + lane) & 1023] = x; }
    t[11] = clock64() - t0;
    if (lane == 0)
        for (int i = 0; i < 12; ++i) out[i] = t[i];
    if (x == 0x12345 && d == 1.5 && fl == 2.5f) out[12] = 1;
}

int main()
{
    long long *d, h[13];
    cudaMalloc(&d, sizeof(h));
    probe<<<1, 32>>>(d, 12345u, 0.75);
    probe<<<1, 32>>>(d, 12345u, 0.75);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[12] = {"match_any", "redux.or", "ballot", "int %", "uint %", "LDS chain", "shfl_up", "ddiv_rn(+add)",
                             "exp f64(+mul,add)", "dmul+dadd", "expf fast(+mul,add)", "sync+LDS+sync+STS"};
    for (int i = 0; i < 12; ++i) printf("%-22s %7.1f cycles\n", names[i], h[i] / 256.0);
    return cudaGetLastError() != cudaSuccess;
}
