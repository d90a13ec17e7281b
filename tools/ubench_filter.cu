// Microbenchmark (dev tool, GPU box): the packed sphere-filter loop in isolation, at 1/2/4/8 warps per
// SM sub-partition, to see how close one warp gets to the FP32 pipe limit (24 pipe-cycles per sphere for
// two rays) and what the FMNMX/SHF tail and the interleave depth cost.
#include "../ataraxia_b200/csrc/atx_device.cuh"
#include <cstdio>
#include <vector>
using namespace atxk;

template <int V>
__device__ __forceinline__ void group8(const float4* q, const RayPair& rp, uint32_t& m0, uint32_t& m1)
{
    if (V == 0)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
            filter_sphere(q[i], rp, m0, m1);
    }
    else if (V == 1)
    {
        // no min(): sign of pre only
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            const float4 sp = q[i];
            const f32x2 ocx = fadd2(rp.ox, pk2(sp.x, sp.x)), ocy = fadd2(rp.oy, pk2(sp.y, sp.y)), ocz = fadd2(rp.oz, pk2(sp.z, sp.z));
            const f32x2 hb = ffma2(ocz, rp.dz, ffma2(ocx, rp.dx, fmul2(ocy, rp.dy)));
            const f32x2 qq = ffma2(ocz, ocz, ffma2(ocx, ocx, fmul2(ocy, ocy)));
            const float nr = fneg(sp.w);
            const f32x2 cc = pk2(ffma(nr, sp.w, lo2(qq)), ffma(nr, sp.w, hi2(qq)));
            const float tiny = 7.888609052210118e-31f;
            const f32x2 pre = ffma2(hb, hb, ffma2(cc, rp.na, pk2(tiny, tiny)));
            m0 = shift_in_sign(m0, lo2(pre));
            m1 = shift_in_sign(m1, hi2(pre));
        }
    }
    else
    {
        // manual 4-way interleave, stage by stage, two halves
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
            float4 sp[4];
            f32x2 ocx[4], ocy[4], ocz[4], hb[4], qq[4], pre[4];
#pragma unroll
            for (int i = 0; i < 4; i++) sp[i] = q[4 * h + i];
#pragma unroll
            for (int i = 0; i < 4; i++) { ocx[i] = fadd2(rp.ox, pk2(sp[i].x, sp[i].x)); ocy[i] = fadd2(rp.oy, pk2(sp[i].y, sp[i].y)); ocz[i] = fadd2(rp.oz, pk2(sp[i].z, sp[i].z)); }
#pragma unroll
            for (int i = 0; i < 4; i++) { hb[i] = fmul2(ocy[i], rp.dy); qq[i] = fmul2(ocy[i], ocy[i]); }
#pragma unroll
            for (int i = 0; i < 4; i++) { hb[i] = ffma2(ocx[i], rp.dx, hb[i]); qq[i] = ffma2(ocx[i], ocx[i], qq[i]); }
#pragma unroll
            for (int i = 0; i < 4; i++) { hb[i] = ffma2(ocz[i], rp.dz, hb[i]); qq[i] = ffma2(ocz[i], ocz[i], qq[i]); }
#pragma unroll
            for (int i = 0; i < 4; i++) { const float nr = fneg(sp[i].w); qq[i] = pk2(ffma(nr, sp[i].w, lo2(qq[i])), ffma(nr, sp[i].w, hi2(qq[i]))); }
            const float tiny = 7.888609052210118e-31f;
#pragma unroll
            for (int i = 0; i < 4; i++) pre[i] = ffma2(qq[i], rp.na, pk2(tiny, tiny));
#pragma unroll
            for (int i = 0; i < 4; i++) pre[i] = ffma2(hb[i], hb[i], pre[i]);
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                if (V == 2) { m0 = shift_in_sign(m0, fmin_(lo2(pre[i]), fneg(lo2(hb[i])))); m1 = shift_in_sign(m1, fmin_(hi2(pre[i]), fneg(hi2(hb[i])))); }
                else { m0 = shift_in_sign(m0, lo2(pre[i])); m1 = shift_in_sign(m1, hi2(pre[i])); }
            }
        }
    }
}

template <int V>
__global__ void k(const float4* sph, int n, int reps, uint32_t* out)
{
    extern __shared__ float4 s[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = sph[i];
    __syncthreads();
    const float t = threadIdx.x * 0.001f + blockIdx.x * 0.01f;
    RayPair rp;
    rp.ox = pk2(t, -t); rp.oy = pk2(1.0f + t, 2.0f - t); rp.oz = pk2(3.0f, 4.0f + t);
    rp.dx = pk2(0.6f, 0.0f); rp.dy = pk2(0.0f, 0.8f); rp.dz = pk2(0.8f, 0.6f);
    rp.na = pk2(-1.0f, -1.0f);
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++)
    {
        uint32_t m0 = 0, m1 = 0;
        const float4* q = s;
#pragma unroll 1
        for (int g = 0; g < n / 8; g++, q += 8)
        {
            group8<V>(q, rp, m0, m1);
            if ((g & 3) == 3) { acc += m0 ^ (m1 * 3u); m0 = m1 = 0; }
        }
        rp.ox = fadd2(rp.ox, pk2(1e-3f, 1e-3f));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
void run(const char* name, const float4* dS, int n, uint32_t* out, int sms, double ghz)
{
    const int cfg[][2] = { {128, 1}, {256, 1}, {512, 1}, {256, 2}, {512, 2}, {256, 3}, {256, 4} };
    printf("%s:", name);
    for (auto& c : cfg)
    {
        const int thr = c[0], ctas = c[1];
        const int reps = 200;
        cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        // shared memory sized so that exactly `ctas` CTAs fit per SM
        const size_t smem = (size_t)(220 * 1024 / ctas) - 2048;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<V><<<sms * ctas, thr, smem>>>(dS, n, 10, out);
        cudaEventRecord(e0);
        k<V><<<sms * ctas, thr, smem>>>(dS, n, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warpsPerSmsp = thr / 32.0 * ctas / 4.0;
        const double spheresPerSmsp = (double)n * reps * warpsPerSmsp;   // warp-level sphere steps per SMSP
        const double cyc = ms * 1e-3 * ghz * 1e9;
        printf("  %gw/smsp: %.1f cyc/sphere", warpsPerSmsp, cyc / spheresPerSmsp);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("(%s)", cudaGetErrorString(e));
    }
    printf("\n");
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk / 1e6;
    const int n = 1024;
    std::vector<float4> h(n);
    for (int i = 0; i < n; i++) h[i] = make_float4(-(float)(i % 37) * 3.0f, -(float)(i % 11), -(float)(i % 53) * 2.0f, 0.5f + (i % 7) * 0.1f);
    float4* dS; cudaMalloc(&dS, n * sizeof(float4)); cudaMemcpy(dS, h.data(), n * sizeof(float4), cudaMemcpyHostToDevice);
    uint32_t* out; cudaMalloc(&out, 4 * 1024 * 1024);
    printf("%s %d SMs %.3f GHz; FP32 pipe limit = 24 cycles per sphere (two rays) per warp-step\n", p.name, p.multiProcessorCount, ghz);
    run<0>("V0 filter_sphere x8 (min+shf)", dS, n, out, p.multiProcessorCount, ghz);
    run<1>("V1 sign(pre) only           ", dS, n, out, p.multiProcessorCount, ghz);
    run<2>("V2 4-way staged + min       ", dS, n, out, p.multiProcessorCount, ghz);
    run<3>("V3 4-way staged, sign(pre)  ", dS, n, out, p.multiProcessorCount, ghz);
    return 0;
}
