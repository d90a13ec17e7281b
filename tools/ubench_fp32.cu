// Microbenchmark (dev tool, GPU box): FP32 issue rates on sm_100a that decide the sphere-loop design.
//   ffma      : scalar fma.rn.ftz.f32, 8 independent chains per thread
//   ffma2     : packed fma.rn.ftz.f32x2, 8 independent chains per thread (16 FMAs per round)
//   fadd2/fmul2 : packed add / mul
//   *_lds     : the same with one broadcast LDS.128 per 12 FP instructions (the sphere loop's ratio)
// Prints lane-FMA/clk/SM for each, to be compared with the 128/clk/SM scalar peak.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define DEV __device__ __forceinline__
DEV float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
DEV unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long r; asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
DEV unsigned long long fadd2(unsigned long long a, unsigned long long b)
{ unsigned long long r; asm volatile("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV unsigned long long fmul2(unsigned long long a, unsigned long long b)
{ unsigned long long r; asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seedf)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seedf, seedf * 0.5f, 1.0f, 0.25f);
    __syncthreads();
    float a[8];
    unsigned long long p[8];
    for (int i = 0; i < 8; i++) { a[i] = seedf + i + threadIdx.x; p[i] = (unsigned long long)__float_as_uint(a[i]) << 32 | __float_as_uint(a[i] * 0.5f); }
    float m = seedf * 0.999f, c = 0.001f;
    unsigned long long pm = (unsigned long long)__float_as_uint(m) << 32 | __float_as_uint(m);
    unsigned long long pc = (unsigned long long)__float_as_uint(c) << 32 | __float_as_uint(c);
    for (int it = 0; it < iters; it++)
    {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 12; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = ffma(a[i], m, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 12; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], pm, pc);
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 12; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = fadd2(p[i], pc);
        } else if (MODE == 3) {
#pragma unroll
            for (int r = 0; r < 12; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = fmul2(p[i], pm);
        } else if (MODE == 4) { // scalar + LDS.128 per 12
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float4 s = sm[(it + r) & 63];
#pragma unroll
                for (int i = 0; i < 12; i++) a[i & 7] = ffma(a[i & 7], s.x, s.w);
            }
        } else if (MODE == 5) { // packed + LDS.128 per 6 packed (=12 FMAs)
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const float4 s = sm[(it + r) & 63];
                const unsigned long long sx = (unsigned long long)__float_as_uint(s.x) << 32 | __float_as_uint(s.y);
                const unsigned long long sw = (unsigned long long)__float_as_uint(s.z) << 32 | __float_as_uint(s.w);
#pragma unroll
                for (int i = 0; i < 6; i++) p[i] = ffma2(p[i], sx, sw);
            }
        } else if (MODE == 6) { // scalar FFMA + scalar FADD mix (2:1)
#pragma unroll
            for (int r = 0; r < 12; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = (i % 3 == 2) ? a[i] + c : ffma(a[i], m, c);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fmasPerIter, int sms, double clkGHz)
{
    float* out; cudaMalloc(&out, sizeof(float) * 256 * sms * 8);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 100, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = fmasPerIter * iters * 256.0 * sms * 8;
    printf("%-12s %8.3f ms  %7.2f Tlane-op/s  = %6.1f lane-ops/clk/SM at %.3f GHz (err=%s)\n", name, ms,
           fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / (clkGHz * 1e9), clkGHz, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk / 1e6;
    printf("%s, %d SMs, %.3f GHz max\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    run<0>("ffma", 96, sms, ghz);
    run<1>("ffma2", 192, sms, ghz);
    run<2>("fadd2", 192, sms, ghz);
    run<3>("fmul2", 192, sms, ghz);
    run<4>("ffma+lds", 96, sms, ghz);
    run<5>("ffma2+lds", 192, sms, ghz);
    run<6>("ffma/fadd", 96, sms, ghz);
    return 0;
}
