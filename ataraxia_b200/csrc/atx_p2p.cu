// atx_p2p.cu — the cross-GPU sum of the float4 accumulation buffers as ONE kernel over NVLink peer memory.
//
// Multi-GPU rendering splits the samples of every pixel across the ranks: each rank ends a step with
// its own float4 partial sums of the whole image, and the image is their sum. Every rank's buffer is mapped into every
// other rank's address space (CUDA IPC, set up once per communicator by atx_capi.cu), and this kernel does the whole
// exchange in place, two-shot:
//   * rank r owns the r-th slice of the image. It reads that slice from EVERY rank's buffer (peer loads over
//     NVLink/NVSwitch), adds the values in rank order 0, 1, ..., N-1 — the same association on every rank, so all
//     ranks end with the same bits, whatever the transport does — and stores the sum into that slice of every rank's
//     buffer (peer stores);
//   * two flag barriers in peer memory bracket it: "my render is complete" before the first peer load (the launch is
//     stream-ordered after the render kernel), "my stores have landed" before the kernel ends. Flags carry the call's
//     epoch, so nothing is ever reset while a peer may still be looking.
// Per rank the wire carries (N-1)/N of the buffer in each direction once; there is no staging copy, no protocol
// chunking and one launch (ncclAllReduce of the same 33 MB took 0.30 ms inside a step at N = 8, mostly latency).
// A peer that never arrives would leave the waiters spinning: every wait gives up after timeoutNs, raises the handle's
// error word (mapped host memory) and lets the kernel finish.
#include "atx_exact.cuh"
#include "atx_kernels.h"

namespace atxk
{

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float4 ld_peer(const float4* p)
{
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag == epoch; false on timeout
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch, unsigned long long timeoutNs)
{
    const unsigned long long t0 = global_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(flag) != epoch)
    {
        if ((++spins & 1023u) == 0u && global_ns() - t0 > timeoutNs)
            return false;
    }
    return true;
}

__global__ void __launch_bounds__(512) p2p_allreduce_kernel(const atx_launch::P2pParams q)
{
    using namespace atx_launch;
    uint32_t* mine = q.flags[q.rank];
    __shared__ int ok;
    if (threadIdx.x == 0)
        ok = 1;
    __syncthreads();
    // ---- "my buffer is complete" to every rank, then wait for everybody's ----
    if (blockIdx.x == 0 && threadIdx.x < q.nRanks)
    {
        __threadfence_system();
        st_release_sys(q.flags[threadIdx.x] + kP2pStart + q.rank, q.epoch);
    }
    if (threadIdx.x < q.nRanks && !wait_flag(mine + kP2pStart + threadIdx.x, q.epoch, q.timeoutNs))
    {
        ok = 0;
        *q.error = 1u;
    }
    __syncthreads();
    if (ok)
    {
        // ---- my slice: sum over the ranks in rank order, result to every rank ----
        const uint32_t chunk = (q.count + q.nRanks - 1u) / q.nRanks;
        const uint32_t lo = min(q.count, q.rank * chunk), hi = min(q.count, lo + chunk);
        for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x)
        {
            float4 v[kP2pMaxRanks];
#pragma unroll
            for (uint32_t r = 0; r < kP2pMaxRanks; r++)
                if (r < q.nRanks)
                    v[r] = ld_peer(q.accum[r] + i);
            float4 s = v[0];
#pragma unroll
            for (uint32_t r = 1; r < kP2pMaxRanks; r++)
                if (r < q.nRanks)
                {
                    s.x = ieee_add(s.x, v[r].x); s.y = ieee_add(s.y, v[r].y);
                    s.z = ieee_add(s.z, v[r].z); s.w = ieee_add(s.w, v[r].w);
                }
#pragma unroll
            for (uint32_t r = 0; r < kP2pMaxRanks; r++)
                if (r < q.nRanks)
                    q.accum[r][i] = s;
        }
    }
    // ---- "my stores have landed" once the last CTA of this rank is through; leave when everybody's have ----
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0)
        last = atomicAdd(mine + kP2pCtaDone, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (last)
    {
        if (threadIdx.x < q.nRanks)
        {
            __threadfence_system();
            st_release_sys(q.flags[threadIdx.x] + kP2pEnd + q.rank, q.epoch);
            if (ok && !wait_flag(mine + kP2pEnd + threadIdx.x, q.epoch, q.timeoutNs))
                *q.error = 2u;
        }
        if (threadIdx.x == 0)
            mine[kP2pCtaDone] = 0u; // for the next call (stream-ordered after this kernel)
    }
}

// fallback of the image-tile split without peer mappings: zero every pixel whose 8x4 tile belongs to another rank, so
// that a plain sum across ranks rebuilds the image
__global__ void keep_own_tiles_kernel(float4* accum, uint32_t width, uint32_t height, uint32_t stride, uint32_t offset)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width * height)
        return;
    const uint32_t x = i % width, y = i / width;
    const uint32_t tile = (y >> 2) * ((width + 7u) >> 3) + (x >> 3);
    if (tile % stride != offset)
        accum[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

} // namespace atxk

namespace atx_launch
{
cudaError_t keep_own_tiles(float4* accum, uint32_t width, uint32_t height, uint32_t stride, uint32_t offset, cudaStream_t s)
{
    const uint32_t n = width * height;
    atxk::keep_own_tiles_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(accum, width, height, stride, offset);
    return cudaGetLastError();
}

cudaError_t p2p_allreduce(const P2pParams& q, int smCount, cudaStream_t s)
{
    if (q.nRanks < 1 || q.nRanks > kP2pMaxRanks)
        return cudaErrorInvalidValue;
    const uint32_t chunk = (q.count + q.nRanks - 1u) / q.nRanks;
    const uint32_t grid = max(1u, min(static_cast<uint32_t>(smCount) * 2u, (chunk + 511u) / 512u));
    atxk::p2p_allreduce_kernel<<<grid, 512, 0, s>>>(q);
    return cudaGetLastError();
}
} // namespace atx_launch
