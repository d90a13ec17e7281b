// atx_wavefront.cu — the wavefront variant of the path-tracing hot path (ATX_VARIANT_WAVEFRONT).
//
// Same per-path code as the megakernels (atx_device.cuh: path_begin / path_hit /
// path_shadow / path_bounce), cut differently: path state lives in HBM, one kernel
// launch advances every live path by ONE bounce, and the paths that continue are
// compacted into the next launch's queue, so a launch never carries dead lanes
// (Renderer::perPixel's loop, Renderer.cu:304-385, turned inside out).
//
//   wf_generate     k frames x P pixels -> path records (primary rays), k = frames per wave
//   wf_bounce       for each queued path: closest-hit trace, then either "miss: the sample
//                   is final" or the whole bounce (hit record, emission, light pick, shadow
//                   trace, Cook-Torrance, roulette, next direction); survivors are pushed to
//                   the output queue with one atomic per warp (ray compaction)
//   wf_accumulate   per pixel: acc += sample[0..k) in frame order (bit-identical to k
//                   sequential reference frames), optional RGBA8 pack
//
// HBM per path: 4 x float4 of state (64 B) + float4 sample (16 B). Per bounce and live path
// the algorithmic traffic is 64 B read + 64 B write + 4 B queue read + 4 B queue write:
// the variant pays HBM bandwidth to buy full warps, which only wins when the megakernels'
// lanes are mostly idle. atx_calibrate() times both on the current scene and
// ATX_VARIANT_AUTO takes the faster.
#include "atx_device.cuh"
#include "atx_kernels.h"

namespace atxk
{

struct WavefrontParams
{
    float4* state;     // [4][capacity]: (o, seed) (d, bounce) (color, -) (throughput, -)
    float4* samples;   // [capacity]
    uint32_t capacity; // paths per wave: framesPerWave * pixels
    uint32_t pixels;
    uint32_t waveFrames;  // frames in this wave
    uint32_t waveFirst;   // index (within the launch) of the wave's first frame
    uint32_t firstWave;   // 1 when this wave starts the launch (zeroFirst applies)
    uint32_t lastWave;    // 1 when RGBA8 is packed after it
};

__device__ __forceinline__ void wf_store(const WavefrontParams& w, uint32_t q, const PathState& s)
{
    w.state[q] = make_float4(s.ox, s.oy, s.oz, __uint_as_float(s.seed));
    w.state[w.capacity + q] = make_float4(s.dx, s.dy, s.dz, __int_as_float(s.bounce));
    w.state[2u * w.capacity + q] = make_float4(s.cr, s.cg, s.cb, 0.0f);
    w.state[3u * w.capacity + q] = make_float4(s.tx, s.ty, s.tz, 0.0f);
}

__device__ __forceinline__ void wf_load(const WavefrontParams& w, uint32_t q, PathState& s)
{
    const float4 a = w.state[q], b = w.state[w.capacity + q], c = w.state[2u * w.capacity + q],
                 d = w.state[3u * w.capacity + q];
    s.ox = a.x; s.oy = a.y; s.oz = a.z; s.seed = __float_as_uint(a.w);
    s.dx = b.x; s.dy = b.y; s.dz = b.z; s.bounce = __float_as_int(b.w);
    s.cr = c.x; s.cg = c.y; s.cb = c.z;
    s.tx = d.x; s.ty = d.y; s.tz = d.z;
}

// path q = slot * pixels + pixel renders frame firstFrame + (waveFirst + slot) * frameStride of `pixel`
__global__ void __launch_bounds__(256) wf_generate(const RenderParams p, const WavefrontParams w)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= w.waveFrames * w.pixels)
        return;
    const uint32_t slot = q / w.pixels, pixel = q - slot * w.pixels;
    const uint32_t x = pixel % p.width, y = pixel / p.width;
    const uint32_t frame = p.firstFrame + (w.waveFirst + slot) * p.frameStride;
    const V3 d0 = primary_direction(p.cam, x, y, p.width, p.height);
    PathState s;
    path_begin(s, p.cam.pos, d0, pixel, frame);
    wf_store(w, q, s);
    w.samples[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // maxBounces < 1: the sample stays (0,0,0)
}

__global__ void __launch_bounds__(256, 3) wf_bounce(const RenderParams p, const WavefrontParams w, const uint32_t* __restrict__ queueIn,
                                                    const uint32_t* __restrict__ countIn, uint32_t* __restrict__ queueOut,
                                                    uint32_t* __restrict__ countOut)
{
    extern __shared__ float4 smem[];
    float4* sphS = smem;
    const uint32_t nIn = queueIn ? *countIn : w.waveFrames * w.pixels;
    if (blockIdx.x * blockDim.x >= nIn)
        return; // launched for the worst case; most late-bounce CTAs leave here
    for (uint32_t i = threadIdx.x; i < p.nSpheres; i += blockDim.x)
        sphS[i] = __ldg(p.spheres + i);
    __syncthreads();

    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < nIn;
    bool survives = false;
    uint32_t q = 0, rays = 0, paths = 0;
    PathState s;
    if (active)
    {
        q = queueIn ? queueIn[i] : i;
        wf_load(w, q, s);
        float tmin = 3.402823466e+38f; // FLT_MAX
        int closest = -1;
        RayConst rk = ray_constants(s.dx, s.dy, s.dz);
        for (uint32_t k = 0; k < p.nSpheres; k++)
            intersect_sphere(sphS[k], static_cast<int>(k), s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, rk, tmin, closest);
        rays++;
        bool ends = true;
        if (closest < 0)
            path_miss(p, s);
        else
        {
            if (path_hit(p, s, sphS[closest], closest, tmin))
            {
                tmin = 3.402823466e+38f;
                closest = -1;
                rk = ray_constants(s.dx, s.dy, s.dz);
                for (uint32_t k = 0; k < p.nSpheres; k++)
                    intersect_sphere(sphS[k], static_cast<int>(k), s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, rk, tmin, closest);
                rays++;
                path_shadow(p, s, closest, tmin);
            }
            ends = path_bounce(p, s);
        }
        if (ends)
        {
            w.samples[q] = make_float4(s.cr, s.cg, s.cb, 1.0f);
            paths = 1;
        }
        else
        {
            wf_store(w, q, s);
            survives = true;
        }
    }
    // ray compaction: one atomic per warp, survivors packed densely into the next queue
    const unsigned mask = __ballot_sync(0xffffffffu, survives);
    if (mask)
    {
        const unsigned lane = threadIdx.x & 31u;
        const int leader = __ffs(mask) - 1;
        uint32_t base = 0;
        if (static_cast<int>(lane) == leader)
            base = atomicAdd(countOut, static_cast<uint32_t>(__popc(mask)));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (survives)
            queueOut[base + __popc(mask & ((1u << lane) - 1u))] = q;
    }
    if (p.counters)
    {
        unsigned long long r = rays, n = paths;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            r += __shfl_xor_sync(0xffffffffu, r, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if ((threadIdx.x & 31u) == 0 && (r | n))
        {
            atomicAdd(p.counters + 0, n);
            atomicAdd(p.counters + 1, r);
            atomicAdd(p.counters + 2, r);
        }
    }
}

// paths that reach the bounce limit inside path_bounce return "ends"; nothing is left in the last
// queue. Samples are added in frame order: accumulation[p] += vec4(color, 1) (Renderer.cu:165, :386).
__global__ void __launch_bounds__(256) wf_accumulate(const RenderParams p, const WavefrontParams w)
{
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= w.pixels)
        return;
    float4 acc = (p.zeroFirst && w.firstWave) ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : p.accum[pixel];
    for (uint32_t slot = 0; slot < w.waveFrames; slot++)
    {
        const float4 c = w.samples[slot * w.pixels + pixel];
        acc.x = fadd(c.x, acc.x); acc.y = fadd(c.y, acc.y); acc.z = fadd(c.z, acc.z);
        acc.w = fadd(acc.w, 1.0f);
    }
    p.accum[pixel] = acc;
    if (p.emitRgba && w.lastWave)
        p.rgba[pixel] = pack_rgba8(acc, u32_to_f32_rn(p.rgbaDivisor));
    if (p.counters && p.maxBounces < 1)
        atomicAdd(p.counters + 0, static_cast<unsigned long long>(w.waveFrames)); // no bounce kernel ran: count the paths here
}

} // namespace atxk

namespace atx_launch
{
using namespace atxk;

size_t wavefront_bytes(uint32_t capacity)
{
    // 4 state rows + samples (float4 each) + two queues + per-bounce counters
    return static_cast<size_t>(capacity) * (5 * sizeof(float4) + 2 * sizeof(uint32_t)) + 256 * sizeof(uint32_t);
}

uint32_t wavefront_frames_per_wave(uint32_t pixels, uint32_t nFrames)
{
    // about 8 M paths in flight: enough to fill the GPU late in the bounce sequence, small next to HBM
    uint32_t k = (8u << 20) / (pixels ? pixels : 1u);
    k = k < 1u ? 1u : (k > 64u ? 64u : k);
    return k < nFrames ? k : nFrames;
}

// One launch worth of frames in waves of `framesPerWave`. `work` is the device scratch of
// wavefront_bytes(framesPerWave * pixels). Returns the number of kernels launched through *launches.
// per device (the attribute belongs to the current device's copy of the function): called from configure()
cudaError_t configure_wavefront()
{
    return cudaFuncSetAttribute(wf_bounce, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
}

cudaError_t render_wavefront(const RenderParams& p, void* work, uint32_t framesPerWave, uint64_t* launches, cudaStream_t s)
{
    const uint32_t P = p.width * p.height;
    const uint32_t capacity = framesPerWave * P;
    if (static_cast<size_t>(p.nSpheres) * sizeof(float4) > static_cast<size_t>(kMaxSmemBytes))
        return cudaErrorInvalidValue;
    WavefrontParams w;
    w.state = static_cast<float4*>(work);
    w.samples = w.state + 4ull * capacity;
    uint32_t* queues = reinterpret_cast<uint32_t*>(w.samples + capacity);
    uint32_t* counts = queues + 2ull * capacity;
    w.capacity = capacity;
    w.pixels = P;
    const size_t smem = sizeof(float4) * (p.nSpheres ? p.nSpheres : 1u);
    for (uint32_t done = 0; done < p.nFrames; done += framesPerWave)
    {
        w.waveFrames = min(framesPerWave, p.nFrames - done);
        w.waveFirst = done;
        w.firstWave = done == 0 ? 1u : 0u;
        w.lastWave = done + w.waveFrames >= p.nFrames ? 1u : 0u;
        const uint32_t n = w.waveFrames * P;
        const uint32_t blocks = (n + 255u) / 256u;
        cudaError_t e = cudaMemsetAsync(counts, 0, 256 * sizeof(uint32_t), s);
        if (e != cudaSuccess)
            return e;
        wf_generate<<<blocks, 256, 0, s>>>(p, w);
        (*launches)++;
        const int bounces = p.maxBounces < 0 ? 0 : p.maxBounces; // <= kWavefrontMaxBounces, checked by the caller
        for (int b = 0; b < bounces; b++)
        {
            const uint32_t* qIn = b == 0 ? nullptr : queues + static_cast<size_t>(b & 1) * capacity;
            uint32_t* qOut = queues + static_cast<size_t>((b + 1) & 1) * capacity;
            wf_bounce<<<blocks, 256, smem, s>>>(p, w, qIn, counts + b, qOut, counts + b + 1);
            (*launches)++;
        }
        wf_accumulate<<<(P + 255u) / 256u, 256, 0, s>>>(p, w);
        (*launches)++;
        e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

} // namespace atx_launch
