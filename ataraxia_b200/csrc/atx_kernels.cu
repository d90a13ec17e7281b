// atx_kernels.cu — hand-written sm_100a kernels of the path-tracing hot path.
//
//   pack_scene_kernel   AoS reference records -> SoA float4 rows (once per upload)
//   megakernel          kernelRender + perPixel (Renderer.cu:150-170, :287-387):
//                       one thread per pixel runs ALL requested frames of that pixel
//                       in one flattened trace loop with path regeneration, sums the
//                       samples in registers in the reference's order and touches the
//                       float4 accumulation buffer once (16 B read + 16 B write)
//   primary_hit_kernel  parity/debug: closest sphere per primary ray
//   ray_dir_kernel      parity/debug: the primary ray table
//   resolve_rgba_kernel display pack of the accumulation buffer
//
// No tensor cores: no stage of this path is a dense contraction. The bound is the
// FP32 FMA pipe (sphere loop) and MUFU (shading); see DESIGN.md §5.
#include "atx_device.cuh"
#include "atx_kernels.h"

namespace atxk
{

// ---------------------------------------------------------------------------
// Scene pack. Runs the per-material subexpressions with the SAME device ops the
// reference executes per bounce, so hoisting them here cannot change a bit.
// ---------------------------------------------------------------------------
__global__ void pack_scene_kernel(const float* __restrict__ sphAoS, uint32_t nSpheres,
                                  const float* __restrict__ matAoS, uint32_t nMaterials,
                                  const float* __restrict__ lightAoS, uint32_t nLights,
                                  float4* __restrict__ spheres, int32_t* __restrict__ sphMat,
                                  float4* __restrict__ mats, float4* __restrict__ lights)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nSpheres)
    {
        const float* s = sphAoS + 5 * i; // Sphere: center[3], radius, id (20 B)
        const float r = s[3];
        spheres[i] = make_float4(s[0], s[1], s[2], r);
        int32_t id = reinterpret_cast<const int32_t*>(s)[4];
        if (static_cast<uint32_t>(id) >= nMaterials) // Renderer.cu:30-37
            id = 0;
        sphMat[i] = id;
    }
    if (i < nMaterials)
    {
        const float* m = matAoS + 13 * i; // Material: albedo[3], roughness, metallic, F0[3], emissionColor[3], emissionIntensity, id
        const float ax = m[0], ay = m[1], az = m[2], rough = m[3], metallic = m[4];
        const float f0x = m[5], f0y = m[6], f0z = m[7];
        const float ecx = m[8], ecy = m[9], ecz = m[10], ei = m[11];
        // baseReflectivity = mix(F0, albedo, metallic) -> fma(F0, 1 - metallic, metallic*albedo)   (Renderer.cu:335)
        const float omm = fsub(1.0f, metallic);
        const float fbx = ffma(f0x, omm, fmul(metallic, ax));
        const float fby = ffma(omm, f0y, fmul(metallic, ay));
        const float fbz = ffma(omm, f0z, fmul(metallic, az));
        const float a = fmul(rough, rough);
        const float a2 = fmul(a, a);
        const float r1 = fadd(rough, 1.0f);
        const float k = fdiv_approx(fmul(r1, r1), 8.0f);
        float4* o = mats + kMatStride * i;
        o[0] = make_float4(ax, ay, az, rough);
        o[1] = make_float4(fbx, fby, fbz, metallic);
        o[2] = make_float4(fsub(1.0f, fbx), fsub(1.0f, fby), fsub(1.0f, fbz), omm);
        o[3] = make_float4(a2, fadd(a2, -1.0f), k, fsub(1.0f, k));
        o[4] = make_float4(fmul(ei, ecx), fmul(ei, ecy), fmul(ei, ecz), ei);
        o[5] = make_float4(ffma(a, a, -1.0f), 0.0f, 0.0f, 0.0f);
    }
    if (i < nLights)
    {
        const float* l = lightAoS + 7 * i; // Light: position[3], color[3], intensity
        lights[kLightStride * i + 0] = make_float4(l[0], l[1], l[2], 0.0f);
        lights[kLightStride * i + 1] = make_float4(fmul(l[6], l[3]), fmul(l[6], l[4]), fmul(l[6], l[5]), 0.0f);
    }
}

// ---------------------------------------------------------------------------
// Shared-memory sphere staging.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void stage_spheres(float4* dst, const float4* __restrict__ src, uint32_t count)
{
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x)
        dst[i] = __ldg(src + i);
}

// trace one ray against spheres [0, count) resident at `sph` (shared memory); indices offset by base
__device__ __forceinline__ void trace_range(const float4* sph, uint32_t count, uint32_t base,
                                            float ox, float oy, float oz, float dx, float dy, float dz,
                                            const RayConst& k, float& tmin, int& closest)
{
#pragma unroll 4
    for (uint32_t i = 0; i < count; i++)
        intersect_sphere(sph[i], static_cast<int>(base + i), ox, oy, oz, dx, dy, dz, k, tmin, closest);
}

// pixel of this thread: a CTA of 256 threads covers a 32x8 tile, each warp an 8x4
// sub-tile (coherent primary rays; 4 x 128 B contiguous float4 segments per warp).
__device__ __forceinline__ void thread_pixel(uint32_t& x, uint32_t& y)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    x = blockIdx.x * 32u + (warp & 3u) * 8u + (lane & 7u);
    y = blockIdx.y * 8u + (warp >> 2) * 4u + (lane >> 3);
}

// ---------------------------------------------------------------------------
// The megakernel.
//
// Per thread: pixel p, frames f = firstFrame + j*frameStride, j < nFrames. The path
// loop of Renderer::perPixel is flattened: every iteration traces ONE ray (closest
// hit or shadow) against all spheres and then runs the matching half of the bounce.
// When a path ends the sample is added to the running sum and the next frame's path
// starts in the same iteration (path regeneration), so lanes stay busy until their
// last frame instead of idling at the slowest path of every frame.
//
// kChunked: the sphere array is larger than one shared-memory chunk; the CTA then
// walks the chunks in lockstep (double-buffered), which needs the outer loop to be
// CTA-uniform (__syncthreads_or on "any thread still has work").
// ---------------------------------------------------------------------------
template <bool kChunked>
__global__ void __launch_bounds__(256, 2) megakernel(const RenderParams p)
{
    extern __shared__ float4 smem[];
    float4* sphS = smem;

    uint32_t x, y;
    thread_pixel(x, y);
    const bool inside = x < p.width && y < p.height;
    const uint32_t pixel = x + y * p.width;

    if (!kChunked)
    {
        stage_spheres(sphS, p.spheres, p.nSpheres);
        __syncthreads();
    }

    // running sum starts from the stored value so the additions happen in the same
    // order as the reference's per-frame "accumulation[p] += color" (Renderer.cu:165)
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (inside && !p.zeroFirst)
        acc = p.accum[pixel];

    V3 d0 = { 0.0f, 0.0f, 0.0f };
    if (inside)
        d0 = primary_direction(p.cam, x, y, p.width, p.height);

    // path state
    uint32_t j = 0;                       // frames done
    uint32_t frame = p.firstFrame;
    bool alive = inside && p.nFrames > 0;
    if (alive && p.maxBounces < 1)
    {
        // perPixel's loop does not run: every sample is (0,0,0,1)   (Renderer.cu:303-304, :386)
        for (uint32_t q = 0; q < p.nFrames; q++)
        {
            acc.x = fadd(0.0f, acc.x); acc.y = fadd(0.0f, acc.y); acc.z = fadd(0.0f, acc.z);
            acc.w = fadd(acc.w, 1.0f);
        }
        j = p.nFrames;
        alive = false;
    }

    float ox = p.cam.pos[0], oy = p.cam.pos[1], oz = p.cam.pos[2];
    float dx = d0.x, dy = d0.y, dz = d0.z;
    float cr = 0.0f, cg = 0.0f, cb = 0.0f;     // color
    float tx = 1.0f, ty = 1.0f, tz = 1.0f;     // throughput
    uint32_t seed = pixel * frame;             // Renderer.cu:300-301 (bounce 0 adds 0)
    int bounce = 0;
    int phase = 0;                             // 0 = closest-hit ray in flight, 1 = shadow ray in flight
    // carried from the closest-hit half to the shadow half of a bounce
    V3 N = { 0.0f, 0.0f, 0.0f }, V = { 0.0f, 0.0f, 0.0f };
    float dist2 = 0.0f;
    int matIndex = 0;
    uint32_t lightIndex = 0;
    uint32_t rays = 0;

    while (kChunked ? __syncthreads_or(alive) : alive)
    {
        // ---- trace the ray in flight against every sphere (Renderer::traceRay) ----
        float tmin = 3.402823466e+38f; // FLT_MAX
        int closest = -1;
        const RayConst rk = ray_constants(dx, dy, dz);
        if (!kChunked)
        {
            trace_range(sphS, p.nSpheres, 0u, ox, oy, oz, dx, dy, dz, rk, tmin, closest);
        }
        else
        {
            // double-buffered chunk walk; all threads of the CTA take part in staging
            const uint32_t C = p.chunkSpheres;
            const uint32_t nChunks = (p.nSpheres + C - 1) / C;
            stage_spheres(sphS, p.spheres, min(C, p.nSpheres));
            for (uint32_t c = 0; c < nChunks; c++)
            {
                __syncthreads(); // chunk c is resident
                float4* cur = sphS + (c & 1u) * C;
                if (c + 1 < nChunks)
                    stage_spheres(sphS + ((c + 1) & 1u) * C, p.spheres + (c + 1) * C, min(C, p.nSpheres - (c + 1) * C));
                if (alive)
                    trace_range(cur, min(C, p.nSpheres - c * C), c * C, ox, oy, oz, dx, dy, dz, rk, tmin, closest);
            }
            __syncthreads(); // nobody still reads the buffers when the next iteration restages
            if (!alive)
                continue;
        }
        rays++;

        bool pathEnds = false;
        bool doBounce = false;
        if (phase == 0)
        {
            if (closest < 0)
            {
                // miss (Renderer.cu:309-318)
                if (p.skyLight)
                {
                    cr = ffma(tx, 0.6f, cr);
                    cg = ffma(ty, 0.7f, cg);
                    cb = ffma(tz, 0.9f, cb);
                }
                pathEnds = true;
            }
            else
            {
                const float4 sp = kChunked ? __ldg(p.spheres + closest) : sphS[closest];
                V3 wp;
                hit_record(sp, ox, oy, oz, dx, dy, dz, tmin, wp, N);
                matIndex = __ldg(p.sphMat + closest);
                const float4 m4 = __ldg(p.mats + kMatStride * matIndex + 4);
                if (m4.w > 0.0f) // emission (Renderer.cu:329-333)
                {
                    cr = ffma(tx, m4.x, cr);
                    cg = ffma(ty, m4.y, cg);
                    cb = ffma(tz, m4.z, cb);
                }
                // next origin == shadow origin: pos + N*1e-4 (Renderer.cu:348, :372), an fma
                const float nox = ffma(N.x, 0.0001f, wp.x);
                const float noy = ffma(N.y, 0.0001f, wp.y);
                const float noz = ffma(N.z, 0.0001f, wp.z);
                if (p.nLights > 0)
                {
                    // light pick reuses the un-advanced seed (Renderer.cu:340)
                    lightIndex = pcg_hash(seed) % p.nLights;
                    const float4 lp = __ldg(p.lights + kLightStride * lightIndex);
                    const float lx = fsub(lp.x, wp.x), ly = fsub(lp.y, wp.y), lz = fsub(lp.z, wp.z);
                    dist2 = fdot3(lx, ly, lz, lx, ly, lz);
                    const float inv = frsqrt_approx(dist2);
                    V = { fsub(0.0f, dx), fsub(0.0f, dy), fsub(0.0f, dz) }; // V = -ray.direction (Renderer.cu:359)
                    dx = fmul(lx, inv); dy = fmul(inv, ly); dz = fmul(inv, lz);
                    phase = 1;
                }
                else
                {
                    doBounce = true;
                }
                ox = nox; oy = noy; oz = noz;
            }
        }
        else
        {
            // shadow result (Renderer.cu:351-368): occluded iff t > 0 && t*t < dist2
            const float ts = closest < 0 ? -1.0f : tmin;
            if (!(ts > 0.0f && fmul(ts, ts) < dist2))
            {
                const float4 m0 = __ldg(p.mats + kMatStride * matIndex + 0);
                const float4 m1 = __ldg(p.mats + kMatStride * matIndex + 1);
                const float4 m2 = __ldg(p.mats + kMatStride * matIndex + 2);
                const float4 m3 = __ldg(p.mats + kMatStride * matIndex + 3);
                const V3 L = { dx, dy, dz };
                const V3 s = cook_torrance(m0, m1, m2, m3, N, V, L);
                const float4 le = __ldg(p.lights + kLightStride * lightIndex + 1);
                // color += emission * specular * throughput / pdf(=1)   (Renderer.cu:362-367): ptxas folds
                // the division by 1 and contracts the last product into the sum (kernelRender SASS 0x32d0-0x3350)
                cr = ffma(fmul(le.x, s.x), tx, cr);
                cg = ffma(fmul(le.y, s.y), ty, cg);
                cb = ffma(fmul(le.z, s.z), tz, cb);
            }
            doBounce = true;
        }

        if (doBounce)
        {
            // throughput, Russian roulette, next direction (Renderer.cu:371-384)
            const float4 m0 = __ldg(p.mats + kMatStride * matIndex + 0);
            const float4 m1 = __ldg(p.mats + kMatStride * matIndex + 1);
            tx = fmul(tx, m0.x); ty = fmul(ty, m0.y); tz = fmul(tz, m0.z);
            const float len = fsqrt_approx(fdot3(tx, ty, tz, tx, ty, tz));
            const float pr = fmax_(fmin_(len, 1.0f), 0.1f);
            if (pcg_float(seed) > pr)
            {
                pathEnds = true;
            }
            else
            {
                tx = fdiv_approx(tx, pr); ty = fdiv_approx(ty, pr); tz = fdiv_approx(tz, pr);
                V3 nd;
                if (m1.w > 0.0f)
                    nd = sample_ggx(N, __ldg(p.mats + kMatStride * matIndex + 5).x, seed);
                else
                    nd = sample_cosine(N, seed);
                dx = nd.x; dy = nd.y; dz = nd.z;
                phase = 0;
                bounce++;
                if (bounce >= p.maxBounces)
                    pathEnds = true;
                else
                    seed += static_cast<uint32_t>(bounce); // Renderer.cu:306
            }
        }

        if (pathEnds)
        {
            // accumulation[p] += vec4(color, 1)   (Renderer.cu:165, :386)
            acc.x = fadd(cr, acc.x); acc.y = fadd(cg, acc.y); acc.z = fadd(cb, acc.z);
            acc.w = fadd(acc.w, 1.0f);
            j++;
            if (j >= p.nFrames)
            {
                alive = false;
            }
            else
            {
                frame += p.frameStride;
                ox = p.cam.pos[0]; oy = p.cam.pos[1]; oz = p.cam.pos[2];
                dx = d0.x; dy = d0.y; dz = d0.z;
                cr = cg = cb = 0.0f;
                tx = ty = tz = 1.0f;
                seed = pixel * frame;
                bounce = 0;
                phase = 0;
            }
        }
    }

    if (inside)
    {
        p.accum[pixel] = acc; // st.global.v4.f32, 512 B contiguous per warp row group
        if (p.emitRgba)
            p.rgba[pixel] = pack_rgba8(acc, u32_to_f32_rn(p.rgbaDivisor));
    }

    // exact counters: one atomic per warp
    if (p.counters)
    {
        unsigned long long r = rays, paths = j; // j == 0 for threads outside the image
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            r += __shfl_xor_sync(0xffffffffu, r, o);
            paths += __shfl_xor_sync(0xffffffffu, paths, o);
        }
        if ((threadIdx.x & 31u) == 0)
        {
            atomicAdd(p.counters + 0, paths);
            atomicAdd(p.counters + 1, r);
        }
    }
}

// ---------------------------------------------------------------------------
// Parity/debug kernels.
// ---------------------------------------------------------------------------
__global__ void primary_hit_kernel(const RenderParams p, int32_t* __restrict__ out)
{
    uint32_t x, y;
    thread_pixel(x, y);
    if (x >= p.width || y >= p.height)
        return;
    const V3 d = primary_direction(p.cam, x, y, p.width, p.height);
    const RayConst rk = ray_constants(d.x, d.y, d.z);
    float tmin = 3.402823466e+38f;
    int closest = -1;
    for (uint32_t i = 0; i < p.nSpheres; i++)
        intersect_sphere(__ldg(p.spheres + i), static_cast<int>(i), p.cam.pos[0], p.cam.pos[1], p.cam.pos[2],
                         d.x, d.y, d.z, rk, tmin, closest);
    out[x + y * p.width] = closest;
}

__global__ void ray_dir_kernel(const RenderParams p, float* __restrict__ out)
{
    uint32_t x, y;
    thread_pixel(x, y);
    if (x >= p.width || y >= p.height)
        return;
    const V3 d = primary_direction(p.cam, x, y, p.width, p.height);
    float* o = out + 3ull * (x + y * p.width);
    o[0] = d.x; o[1] = d.y; o[2] = d.z;
}

__global__ void resolve_rgba_kernel(const float4* __restrict__ accum, uint32_t* __restrict__ rgba, uint32_t n,
                                    uint32_t divisor)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        rgba[i] = pack_rgba8(accum[i], u32_to_f32_rn(divisor));
}

} // namespace atxk

// ---------------------------------------------------------------------------
// Host-callable launchers (the only symbols the C-ABI layer uses).
// ---------------------------------------------------------------------------
namespace atx_launch
{
using namespace atxk;

static dim3 tile_grid(uint32_t w, uint32_t h) { return dim3((w + 31u) / 32u, (h + 7u) / 8u); }

cudaError_t pack_scene(const float* sphAoS, uint32_t nS, const float* matAoS, uint32_t nM, const float* lightAoS,
                       uint32_t nL, float4* spheres, int32_t* sphMat, float4* mats, float4* lights, cudaStream_t s)
{
    const uint32_t n = max(nS, max(nM, nL));
    if (n == 0)
        return cudaSuccess;
    pack_scene_kernel<<<(n + 255) / 256, 256, 0, s>>>(sphAoS, nS, matAoS, nM, lightAoS, nL, spheres, sphMat, mats, lights);
    return cudaGetLastError();
}

size_t megakernel_smem_bytes(const RenderParams& p)
{
    const bool chunked = p.chunkSpheres < p.nSpheres;
    return sizeof(float4) * (chunked ? 2ull * p.chunkSpheres : static_cast<size_t>(p.nSpheres));
}

cudaError_t configure()
{
    cudaError_t e = cudaFuncSetAttribute(megakernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess)
        return e;
    return cudaFuncSetAttribute(megakernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
}

cudaError_t render_mega(const RenderParams& p, cudaStream_t s)
{
    const bool chunked = p.chunkSpheres < p.nSpheres;
    const size_t smem = megakernel_smem_bytes(p);
    const dim3 grid = tile_grid(p.width, p.height);
    if (chunked)
        megakernel<true><<<grid, 256, smem, s>>>(p);
    else
        megakernel<false><<<grid, 256, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t primary_hits(const RenderParams& p, int32_t* out, cudaStream_t s)
{
    primary_hit_kernel<<<tile_grid(p.width, p.height), 256, 0, s>>>(p, out);
    return cudaGetLastError();
}

cudaError_t ray_directions(const RenderParams& p, float* out, cudaStream_t s)
{
    ray_dir_kernel<<<tile_grid(p.width, p.height), 256, 0, s>>>(p, out);
    return cudaGetLastError();
}

cudaError_t resolve_rgba(const float4* accum, uint32_t* rgba, uint32_t n, uint32_t divisor, cudaStream_t s)
{
    resolve_rgba_kernel<<<(n + 255) / 256, 256, 0, s>>>(accum, rgba, n, divisor);
    return cudaGetLastError();
}

} // namespace atx_launch
