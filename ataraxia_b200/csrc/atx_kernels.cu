// atx_kernels.cu — hand-written sm_100a kernels of the path-tracing hot path.
//
//   pack_scene_kernel      AoS reference records -> SoA float4 rows (once per upload), incl. the line filter's records
//   megakernel_*           kernelRender + perPixel (Renderer.cu:150-170, :287-387) as persistent kernels over a pixel
//                          pool: a lane (or a path slot) runs ALL requested frames of a pixel, sums the samples in the
//                          reference's frame order and touches the float4 accumulation buffer once (16 B read + 16 B write)
//       _ww                  while-while form: small scenes, few frames per launch or several lights
//       _wq                  warp-queue form: small scenes, hits queued per warp and bounced 32 at a time
//       _pair                two path slots per thread, packed f32x2 line filter + exact replay, slots free-running
//       _pair_ls             the same trace with all slots of a CTA in lockstep (closest-hit trace, shadow trace)
//   pixel_prologue_kernel  per-pixel launch constants of the warp-queue form, 32 lanes wide
//   primary_hit_kernel     parity/debug: closest sphere per primary ray
//   ray_dir_kernel         parity/debug: the primary ray table
//   resolve_rgba_kernel    display pack of the accumulation buffer
//
// No tensor cores: no stage of this path is a dense contraction. The bound is FP32 issue (sphere loop) and
// instruction issue at large (shading); see DESIGN.md sections 3 and 5.
#include "atx_device.cuh"
#include "atx_kernels.h"
#include <cstdio>

namespace atxk
{

// ---------------------------------------------------------------------------
// Scene pack. Runs the per-material subexpressions with the SAME device ops the
// reference executes per bounce, so hoisting them here cannot change a bit.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float flip_sign(float v) { return __uint_as_float(__float_as_uint(v) ^ 0x80000000u); }

__global__ void pack_scene_kernel(const float* __restrict__ sphAoS, uint32_t nSpheres,
                                  const float* __restrict__ matAoS, uint32_t nMaterials,
                                  const float* __restrict__ lightAoS, uint32_t nLights,
                                  float4* __restrict__ spheres, float4* __restrict__ sphFilter, int32_t* __restrict__ sphMat,
                                  float4* __restrict__ mats, float4* __restrict__ lights)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nSpheres)
    {
        const float* s = sphAoS + 5 * i; // Sphere: center[3], radius, id (20 B)
        const float r = s[3];
        // centre stored negated (pure sign flip): o - c == o + (-c) bit for bit
        spheres[i] = make_float4(flip_sign(s[0]), flip_sign(s[1]), flip_sign(s[2]), r);
        // the line filter's constant (filter_sphere): |c|^2 - r^2 less the one-sided margin, one rounding
        const double c2 = double(s[0]) * double(s[0]) + double(s[1]) * double(s[1]) + double(s[2]) * double(s[2]);
        const double r2 = double(r) * double(r);
        float kk = static_cast<float>(c2 - r2 - double(kFilterMargin) * (c2 + r2));
        if (!(fabsf(kk) <= 3.402823466e+38f))
            kk = -3.402823466e+38f; // overflow or NaN in the inputs: always a candidate, the exact test decides
        sphFilter[i] = make_float4(flip_sign(s[0]), flip_sign(s[1]), flip_sign(s[2]), kk);
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
// Shared-memory sphere staging. The resident array is padded with zero records up to a
// multiple of 8 so the packed loop can run whole groups; padded slots are masked out of
// the candidate set, never tested.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t round_up8(uint32_t n) { return (n + 7u) & ~7u; }

__device__ __forceinline__ uint32_t round_up32(uint32_t n) { return (n + 31u) & ~31u; }

// pad = 8 (scalar forms) or 32 (packed forms: whole blocks of 32 filter steps)
__device__ __forceinline__ void stage_spheres(float4* dst, const float4* __restrict__ src, uint32_t count, uint32_t pad = 8u)
{
    const uint32_t padded = (count + pad - 1u) & ~(pad - 1u);
    for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x)
        dst[i] = i < count ? __ldg(src + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- bulk (TMA) staging of sphere chunks: cp.async.bulk global -> shared, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}

// one thread: expect `bytes` on the barrier, then start the copy (16 B aligned, multiple of 16 B)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "ATX_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@!p bra ATX_WAIT;\n"
                 "}\n" ::"r"(smem_addr(bar)),
                 "r"(parity)
                 : "memory");
}

// scalar: one ray against spheres [0, count) resident at `sph`; indices offset by base
// kCount > 0: the scene has exactly kCount spheres (compile-time: straight-line code, the tests of one ray interleave)
template <bool kFlat = false, int kCount = 0>
__device__ __forceinline__ void trace_range(const float4* sph, uint32_t count, uint32_t base,
                                            float ox, float oy, float oz, float dx, float dy, float dz,
                                            const RayConst& k, float& tmin, int& closest)
{
    if (kCount > 0)
    {
#pragma unroll
        for (int i = 0; i < kCount; i++)
        {
            if (kFlat)
                exact_flat(sph[i], static_cast<int>(base) + i, ox, oy, oz, dx, dy, dz, k, tmin, closest);
            else
                intersect_sphere(sph[i], static_cast<int>(base) + i, ox, oy, oz, dx, dy, dz, k, tmin, closest);
        }
        return;
    }
#pragma unroll 4
    for (uint32_t i = 0; i < count; i++)
    {
        if (kFlat)
            exact_flat(sph[i], static_cast<int>(base + i), ox, oy, oz, dx, dy, dz, k, tmin, closest);
        else
            intersect_sphere(sph[i], static_cast<int>(base + i), ox, oy, oz, dx, dy, dz, k, tmin, closest);
    }
}

// packed: the two rays of a thread against the filter records of spheres [0, count) resident at `sph`
// (shared memory, padded to a multiple of 32); `exact` is the global array of exact records (indexed from
// base: the few candidates of a lane are read through L1). Two stages per super-block of up to kSuperBlock spheres:
//   filter   blocks of 32 spheres, branch-free: two 32-bit candidate words per block go to this
//            thread's private column of `cand` (shared memory, conflict-free), plus one bit per
//            block in a "non-empty" word per slot;
//   resolve  each lane walks its own candidates in ascending sphere index with the exact
//            sequence (Renderer::traceRay keeps the lowest index on ties: strict '<',
//            Renderer.cu:272). The loop runs max-over-lanes(candidates) times per super-block,
//            not once per block, so the few lanes with work share their iterations.
constexpr uint32_t kSuperBlock = 512;                   // spheres resolved together
// candidate words per thread: two slots x blocks of the largest super-block a launch sees
__host__ __device__ __forceinline__ uint32_t cand_words(const RenderParams& p)
{
    const uint32_t resident = p.chunkSpheres < p.nSpheres ? p.chunkSpheres : p.nSpheres;
    const uint32_t span = resident < kSuperBlock ? resident : kSuperBlock;
    return 2u * ((span + 31u) / 32u);
}

__device__ __forceinline__ void trace_range2(const float4* sph, const float4* __restrict__ exact, uint32_t count, uint32_t base, uint32_t* cand,
                                             const RayPair& rp, const PathState& s0, const PathState& s1,
                                             bool live0, bool live1, const RayConst& k0, const RayConst& k1,
                                             float& tmin0, int& closest0, float& tmin1, int& closest1)
{
    uint32_t* mine = cand + threadIdx.x; // word w of this thread: mine[w * blockDim.x]
    const uint32_t T = blockDim.x;
    for (uint32_t sb = 0; sb < count; sb += kSuperBlock)
    {
        const uint32_t sbCount = min(kSuperBlock, count - sb);
        const uint32_t nBlocks = (sbCount + 31u) >> 5;
        uint32_t nz0 = 0u, nz1 = 0u;
        const float4* q = sph + sb;
        uint32_t* w = mine;
        // whole blocks of 32 filter steps, straight-line: the resident array is padded to a multiple of 32, and whatever the
        // padding records answer is masked out of the last block below
#pragma unroll 1
        for (uint32_t blk = 0; blk < nBlocks; blk++, q += 32, w += 2u * T)
        {
            uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
            for (int i = 0; i < 32; i++)
                filter_sphere(q[i], rp, m0, m1);
            // sphere i of the block sits at bit (31 - i); a clear sign bit made it a candidate
            const uint32_t c0 = ~m0, c1 = ~m1;
            w[0] = c0;
            w[T] = c1;
            nz0 = (nz0 << 1) | min(c0, 1u);
            nz1 = (nz1 << 1) | min(c1, 1u);
        }
        if (sbCount & 31u)
        {
            // the last block is partial: drop the padded tail from its two words
            const uint32_t valid = 0xFFFFFFFFu << (32u - (sbCount & 31u));
            uint32_t* wl = mine + 2u * (nBlocks - 1u) * T;
            const uint32_t c0 = wl[0] & valid, c1 = wl[T] & valid;
            wl[0] = c0;
            wl[T] = c1;
            nz0 = (nz0 & ~1u) | min(c0, 1u);
            nz1 = (nz1 & ~1u) | min(c1, 1u);
        }
        // block blk at bit (31 - blk); a retired slot has no candidates
        nz0 = live0 ? nz0 << (32u - nBlocks) : 0u;
        nz1 = live1 ? nz1 << (32u - nBlocks) : 0u;
        uint32_t cur0 = 0u, cur1 = 0u, b0 = 0u, b1 = 0u;
        while (true)
        {
            if (cur0 == 0u && nz0 != 0u)
            {
                b0 = __clz(nz0);
                nz0 &= ~(0x80000000u >> b0);
                cur0 = mine[(2u * b0) * T];
            }
            if (cur1 == 0u && nz1 != 0u)
            {
                b1 = __clz(nz1);
                nz1 &= ~(0x80000000u >> b1);
                cur1 = mine[(2u * b1 + 1u) * T];
            }
            if ((cur0 | cur1) == 0u)
                break;
            if (cur0)
            {
                const uint32_t i = __clz(cur0);
                cur0 &= ~(0x80000000u >> i);
                const uint32_t idx = sb + b0 * 32u + i;
                exact_test(__ldg(exact + base + idx), static_cast<int>(base + idx), s0.ox, s0.oy, s0.oz, s0.dx, s0.dy, s0.dz, k0, tmin0, closest0);
            }
            if (cur1)
            {
                const uint32_t i = __clz(cur1);
                cur1 &= ~(0x80000000u >> i);
                const uint32_t idx = sb + b1 * 32u + i;
                exact_test(__ldg(exact + base + idx), static_cast<int>(base + idx), s1.ox, s1.oy, s1.oz, s1.dx, s1.dy, s1.dz, k1, tmin1, closest1);
            }
        }
    }
}

// pixel of this thread, one pixel per thread: a CTA of 256 threads covers a 32x8 tile, each
// warp an 8x4 sub-tile (coherent primary rays; 4 x 128 B contiguous float4 segments per warp).
__device__ __forceinline__ void thread_pixel(uint32_t& x, uint32_t& y)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    x = blockIdx.x * 32u + (warp & 3u) * 8u + (lane & 7u);
    y = blockIdx.y * 8u + (warp >> 2) * 4u + (lane >> 3);
}

// ---------------------------------------------------------------------------
// Pixel pool. Paths differ in length from pixel to pixel, so a fixed pixel-to-lane map
// leaves lanes idle while the slowest pixel of their warp finishes (measured: 17 of 32
// lanes active on config 2). Both megakernels therefore run as persistent CTAs whose
// lanes CLAIM pixels from one global counter: ids are handed out in 8x4-tile order (32
// consecutive ids = one tile, so a full-warp claim is one coherent tile), a lane keeps a
// pixel for all its frames (the frame order of the float4 sum is what makes the result
// bit-identical to the reference), stores the 16 B result itself and claims the next id.
// Claims are batched per warp (at least claimThreshold idle lanes, or nothing left to do)
// so the per-pixel prologue runs with several lanes active. Which lane renders which pixel
// never changes a result.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool pool_pixel(const RenderParams& p, uint32_t id, uint32_t& x, uint32_t& y)
{
    const uint32_t tilesX = (p.width + 7u) >> 3;
    const uint32_t tile = (id >> 5) * p.tileStride + p.tileOffset, w = id & 31u; // (stride 1, offset 0 unless the image is split across GPUs)
    x = (tile % tilesX) * 8u + (w & 7u);
    y = (tile / tilesX) * 4u + (w >> 3);
    return tile < p.nTiles && x < p.width && y < p.height;
}

// warp-aggregated claim: one atomic per warp; returns this lane's id (valid where need is set)
__device__ __forceinline__ uint32_t pool_claim(uint32_t* counter, bool need)
{
    const unsigned mask = __ballot_sync(0xffffffffu, need);
    if (mask == 0u)
        return 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0u;
    if (static_cast<int>(lane) == leader)
        base = atomicAdd(counter, static_cast<uint32_t>(__popc(mask)));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void count_rays(const RenderParams& p, uint32_t rays, uint32_t traced, uint32_t paths)
{
    // exact counters: one atomic per warp (all 32 lanes of the warp reach this point)
    if (!p.counters)
        return;
    unsigned long long r = rays, t = traced, n = paths;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31u) == 0)
    {
        atomicAdd(p.counters + 0, n);
        atomicAdd(p.counters + 1, r);
        atomicAdd(p.counters + 2, t);
    }
}

// n samples of the same color c added one by one: accumulation[p] += vec4(c, 1)  (Renderer.cu:165, :386).
// When c is exactly (0,0,0) the loop has a closed form that is bit-identical: x + 0 is x after the first
// addition (which also turns a -0 into +0), and the count stays an exactly representable integer while
// it is below 2^24, so w + 1 + ... + 1 == w + n.
__device__ __forceinline__ void accumulate_constant(float4& acc, float cr, float cg, float cb, uint32_t n)
{
    const bool zero = cr == 0.0f && cg == 0.0f && cb == 0.0f;
    const float w = acc.w;
    if (zero && n > 0u && w >= 0.0f && w == floorf(w) && w + static_cast<float>(n) <= 16777216.0f && n <= 16777216u)
    {
        acc.x = fadd(cr, acc.x); acc.y = fadd(cg, acc.y); acc.z = fadd(cb, acc.z);
        acc.w = w + static_cast<float>(n);
        return;
    }
    for (uint32_t q = 0; q < n; q++)
    {
        acc.x = fadd(cr, acc.x); acc.y = fadd(cg, acc.y); acc.z = fadd(cb, acc.z);
        acc.w = fadd(acc.w, 1.0f);
    }
}

// perPixel's loop does not run when maxBounces < 1: every sample is (0,0,0,1)   (Renderer.cu:303-304, :386)
__device__ __forceinline__ void accumulate_black(float4& acc, uint32_t n) { accumulate_constant(acc, 0.0f, 0.0f, 0.0f, n); }

// A pixel whose primary ray misses every sphere: each frame's path is "miss at bounce 0"
// (Renderer.cu:309-318 with throughput 1), so all its samples are the same color.
__device__ __forceinline__ void accumulate_sky(const RenderParams& p, float4& acc, uint32_t n)
{
    PathState s;
    s.cr = s.cg = s.cb = 0.0f;
    s.tx = s.ty = s.tz = 1.0f;
    path_miss(p, s);
    accumulate_constant(acc, s.cr, s.cg, s.cb, n);
}

// ---------------------------------------------------------------------------
// The primary ray of a pixel is the same every frame: Camera::UpdateRayDirection has no
// jitter and no half-pixel offset (Camera.cpp:176-187), and ray.origin is the camera
// position (Renderer.cu:291-293). Its traceRay result (t, sphere) is therefore a per-pixel
// constant of the launch: a lane traces it once when it claims the pixel and starts every
// frame's path from that hit record. The first traceRay call of frames 2..n is not
// executed; `rays` still counts it (it is a traceRay call of the reference), `traced`
// counts the rays whose sphere loop really ran (the roofline uses `traced`).
// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// Megakernel, while-while form (small scenes: shading dominates the sphere loop).
//
// One lane = one claimed pixel, all requested frames f = firstFrame + j*frameStride of it.
// A lane is either "in flight" (its path has a ray to trace) or "parked" (a hit is waiting
// for its bounce). The warp alternates between two phases:
//   A  in-flight lanes trace their ray; a miss ends the path, adds the sample to the
//      running sum and starts the next frame's path at once (path regeneration); a hit is
//      parked;
//   B  parked lanes run the whole bounce in lockstep: hit record, emission, light pick,
//      shadow trace, Cook-Torrance, Russian roulette, next direction.
// Phase B is the expensive part (~5x a 3-sphere trace). It runs when at least
// `parkThreshold` lanes are parked or nothing is in flight, so its lanes are full, and
// lanes that drift apart fall back into step instead of staying out of phase for the
// rest of the launch. A freshly claimed pixel needs no code of its own: its primary ray
// goes through phase A and its primary hit through phase B like any other (`fresh`).
//
// kFixedLight (numLights <= 1): the light pick "PcgHash(seed) % numLights" (Renderer.cu:340)
// is 0 for every seed, so the whole first bounce up to the roulette - hit record,
// emission, shadow ray, Cook-Torrance - is the same for every frame of a pixel. It runs
// once (the fresh pass) and every frame's path starts from that state (color, next
// origin, normal, material) with its own seed; only path_bounce onwards runs per frame.
//
// The samples of a pixel are summed in registers in frame order, so the float4 sums are
// bit-identical to sequential reference frames, and the accumulation buffer is touched
// once: one 16 B read + one 16 B write per pixel per launch.
// ---------------------------------------------------------------------------
#ifndef ATX_WW_FLAT
#define ATX_WW_FLAT 1 // sphere tests of the while-while form: 1 = branch-free (exact_flat), 0 = line filter + hit branch
#endif
#ifndef ATX_SMALL_STATIC
#define ATX_SMALL_STATIC 4 // scenes of up to this many spheres get small-scene kernels with the count compiled in (0: none)
#endif
// kN > 0: the scene has exactly kN spheres (straight-line sphere tests); 0: any count up to kWhileWhileMaxSpheres
template <bool kFixedLight, int kN>
__global__ void __launch_bounds__(256, 3) megakernel_ww(const RenderParams p)
{
    extern __shared__ float4 smem[];
    float4* sphS = smem;
    constexpr unsigned kFull = 0xffffffffu;

    stage_spheres(sphS, p.spheres, p.nSpheres);
    __syncthreads();

    // lane state
    bool alive = false;     // owns a pixel with frames left
    bool parked = false;    // a hit waits for its bounce (tminP, closestP)
    bool fresh = false;     // the ray in flight / hit parked is the pixel's primary
    bool exhausted = false; // the pool has no more pixels
    uint32_t pixel = 0, j = 0, frame = 0;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    V3 d0 = { 0.0f, 0.0f, 0.0f };
    PathState s;
    path_begin(s, p.cam.pos, d0, 0u, 0u);
    s.N = s.V = d0;
    s.dist2 = 0.0f;
    s.matIndex = 0;
    s.lightIndex = 0u;
    float tminP = 0.0f, tPrimary = 0.0f;
    int closestP = -1, cPrimary = -1;
    uint32_t raysPerStart = 1; // reference traceRay calls a cached start stands for
    // kFixedLight: what every frame of this pixel shares - the path state after the first bounce's shading
    // (color, next origin, normal, material) and the frame-independent half of the first path_bounce:
    // throughput = 1 * albedo, so the roulette probability pr0, the throughput after it (tq) and the
    // tangent frame of N are per-pixel constants; each frame only draws its three random numbers
    float c0r = 0.0f, c0g = 0.0f, c0b = 0.0f, o0x = 0.0f, o0y = 0.0f, o0z = 0.0f;
    V3 N0 = { 0.0f, 0.0f, 0.0f }, T0 = { 0.0f, 0.0f, 0.0f }, B0 = { 0.0f, 0.0f, 0.0f };
    float pr0 = 0.0f, tq0x = 0.0f, tq0y = 0.0f, tq0z = 0.0f, ggxT0 = 0.0f;
    int mat0 = 0;
    bool ggx0 = false;
    uint32_t rays = 0, traced = 0, paths = 0;

    auto trace = [&](float& tmin, int& closest) {
        tmin = 3.402823466e+38f; // FLT_MAX
        closest = -1;
        const RayConst rk = ray_constants(s.dx, s.dy, s.dz);
        trace_range<ATX_WW_FLAT != 0, kN>(sphS, p.nSpheres, 0u, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, rk, tmin, closest);
        traced++;
    };
    // the pixel is complete: st.global.v4.f32 of the running sum (+ the display pack)
    auto retire = [&]() {
        store_pixel(p, pixel, acc);
        if (p.emitRgba)
            p.rgba[pixel] = pack_rgba8(acc, u32_to_f32_rn(p.rgbaDivisor));
        paths += j;
        alive = false;
        parked = false;
    };
    // Start frame `frame`'s path and run it to its first per-frame trace: on return the lane is
    // parked, in flight, or has retired its pixel. (kFixedLight loops only when a path ends at
    // its first roulette / bounce limit.)
    auto start = [&]() {
        while (true)
        {
            if (j >= p.nFrames)
            {
                retire();
                return;
            }
            rays += raysPerStart;
            if (!kFixedLight)
            {
                path_begin(s, p.cam.pos, d0, pixel, frame);
                parked = true;
                tminP = tPrimary;
                closestP = cPrimary;
                return;
            }
            // path_bounce (Renderer.cu:371-384) at bounce 0 with the per-pixel constants folded in
            s.cr = c0r; s.cg = c0g; s.cb = c0b;
            s.seed = pixel * frame;
            parked = false;
            if (!(pcg_float(s.seed) > pr0))
            {
                float x, y, z;
                sample_local(ggx0, ggxT0, s.seed, x, y, z);
                const V3 nd = frame_combine(N0, T0, B0, x, y, z);
                if (1 < p.maxBounces)
                {
                    s.ox = o0x; s.oy = o0y; s.oz = o0z;
                    s.dx = nd.x; s.dy = nd.y; s.dz = nd.z;
                    s.tx = tq0x; s.ty = tq0y; s.tz = tq0z;
                    s.N = N0;
                    s.matIndex = mat0;
                    s.bounce = 1;
                    s.seed += 1u; // Renderer.cu:306
                    return;
                }
            }
            // the path ended at its first roulette or at the bounce limit: the sample is the cached color
            accumulate_sample(acc, s);
            j++;
            frame += p.frameStride;
        }
    };
    auto finish = [&]() {
        accumulate_sample(acc, s);
        j++;
        frame += p.frameStride;
        start();
    };

    while (true)
    {
        // ---- claim pixels for idle lanes ----
        const unsigned aliveMask = __ballot_sync(kFull, alive);
        if (aliveMask != kFull)
        {
            const bool need = !alive && !exhausted;
            const unsigned needMask = __ballot_sync(kFull, need);
            if (needMask != 0u && (static_cast<uint32_t>(__popc(needMask)) >= p.claimThreshold || aliveMask == 0u))
            {
                const uint32_t id = pool_claim(p.pool, need);
                uint32_t x, y;
                if (need)
                {
                    if (id >= p.poolSize)
                        exhausted = true;
                    else if (pool_pixel(p, id, x, y))
                    {
                        pixel = x + y * p.width;
                        // running sum starts from the stored value so the additions happen in the same
                        // order as the reference's per-frame "accumulation[p] += color" (Renderer.cu:165)
                        acc = p.zeroFirst ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : p.accum[pixel];
                        j = 0;
                        frame = p.firstFrame;
                        if (p.maxBounces < 1)
                        {
                            accumulate_black(acc, p.nFrames);
                            j = p.nFrames;
                            retire();
                        }
                        else
                        {
                            d0 = primary_direction(p.cam, x, y, p.width, p.height);
                            path_begin(s, p.cam.pos, d0, pixel, frame);
                            alive = true;
                            fresh = true;
                            parked = false;
                        }
                    }
                }
            }
            if (!__any_sync(kFull, alive))
            {
                if (__all_sync(kFull, exhausted))
                    break;
                continue;
            }
        }

        const unsigned parkedMask = __ballot_sync(kFull, parked);
        const unsigned nParked = __popc(parkedMask);
        const bool anyFlying = (__ballot_sync(kFull, alive) & ~parkedMask) != 0u;
        if (nParked >= p.parkThreshold || !anyFlying)
        {
            // ---- phase B: one bounce for every parked hit ----
            if (parked)
            {
                parked = false;
                if (fresh && !kFixedLight)
                {
                    tPrimary = tminP;
                    cPrimary = closestP;
                    rays++; // the first frame's primary traceRay call
                    fresh = false;
                }
                if (path_hit(p, s, sphS[closestP], closestP, tminP))
                {
                    float tmin;
                    int closest;
                    trace(tmin, closest);
                    if (!fresh)
                        rays++;
                    path_shadow(p, s, closest, tmin);
                    if (fresh)
                        raysPerStart = 2;
                }
                else if (fresh)
                    raysPerStart = 1;
                if (fresh)
                {
                    // kFixedLight: this was the frame-independent first bounce; keep its result
                    c0r = s.cr; c0g = s.cg; c0b = s.cb;
                    o0x = s.ox; o0y = s.oy; o0z = s.oz;
                    N0 = s.N;
                    mat0 = s.matIndex;
                    {
                        // the same instructions path_bounce runs, on the values every frame would feed it
                        const float4 m0 = __ldg(p.mats + kMatStride * mat0 + 0);
                        const float4 m1 = __ldg(p.mats + kMatStride * mat0 + 1);
                        const float ax = fmul(1.0f, m0.x), ay = fmul(1.0f, m0.y), az = fmul(1.0f, m0.z);
                        const float len = fsqrt_approx(fdot3(ax, ay, az, ax, ay, az));
                        pr0 = fmax_(fmin_(len, 1.0f), 0.1f);
                        tq0x = fdiv_approx(ax, pr0); tq0y = fdiv_approx(ay, pr0); tq0z = fdiv_approx(az, pr0);
                        ggx0 = m1.w > 0.0f;
                        ggxT0 = ggx0 ? __ldg(p.mats + kMatStride * mat0 + 5).x : 0.0f;
                        tangent_frame(N0, T0, B0);
                    }
                    fresh = false;
                    start();
                }
                else if (path_bounce(p, s))
                    finish();
            }
        }
        else if (alive && !parked)
        {
            // ---- phase A: trace the ray in flight ----
            float tmin;
            int closest;
            trace(tmin, closest);
            if (!fresh)
                rays++;
            if (closest >= 0)
            {
                parked = true;
                tminP = tmin;
                closestP = closest;
            }
            else if (fresh)
            {
                // the primary ray misses: every frame of this pixel is a bounce-0 miss
                accumulate_sky(p, acc, p.nFrames);
                rays += p.nFrames;
                j = p.nFrames;
                fresh = false;
                retire();
            }
            else
            {
                path_miss(p, s);
                finish();
            }
        }
    }
    count_rays(p, rays, traced, paths);
}

// ---------------------------------------------------------------------------
// Megakernel, warp-queue form (small scenes, many frames per launch).
//
// The while-while form above leaves lanes idle twice over: a lane whose path has a hit
// waits for the bounce phase (which then runs with the 8-16 lanes that are parked), and the
// frames of a pixel are strictly serial on their lane (measured on config 2: 17 of 32 lanes
// active). Here every lane owns a pixel (claimed from the global pool, a few lanes at a time)
// and the warp keeps two kinds of work apart, each close to full width:
//   G  "generate": lane i starts the next frame of ITS pixel (the per-pixel constants of the
//      launch - primary hit, and with one light the whole first bounce - stay in its
//      registers) and traces that ray; a miss ends the path, a hit is pushed onto the warp's
//      hit queue in shared memory (17 words: ray, color, throughput, seed, bounce, t, sphere,
//      and a tag = owning lane + ring slot);
//   B  "bounce": every lane pops ONE queued hit (any lane's pixel), runs the whole bounce -
//      hit record, emission, shadow trace, Cook-Torrance, roulette, next direction - traces
//      the continuing ray and pushes it back if it hits again.
// A lane is never parked: after a hit it goes on with the next frame of its pixel. The
// samples of a pixel therefore complete out of order, while the reference adds them in
// frame order (accumulation[p] += color once per frame, Renderer.cu:165) and float sums do
// not commute bit for bit. Each lane has a ring of kRingFrames sample slots in shared
// memory; a finished path writes its color to its slot and sets the slot's bit, and the
// owning lane adds completed slots to its running sum strictly in frame order (a sample
// that completes while nothing older is pending goes straight to the sum). A lane whose
// ring is full skips G until its oldest path has come back; B runs at once with 32 queued
// hits and earlier when stalled lanes have cost as much as a narrow B would waste.
// Measured on config 2 (1024 frames): G runs 28.9 lanes wide, B 28.4; 22.6 ms against the
// while-while form's 31.1 ms. Shared memory (10.4 KB per warp) allows 20 warps per SM; a
// 32-slot ring at 12 warps per SM is slower (32.5 ms), an 8-slot ring at 24 warps equal.
// Same path_* code between traces as every other form, same frame order of the sums: the
// results are bit-identical.
//
// Round 2, 22.6 -> 16.5 ms on config 2 (all bit-identical to the reference's 1024 frames):
//  * sphere tests as the reference's sequence without its branches and without its second root
//    (flat_tail, atx_device.cuh), sphere count compiled in for scenes of up to 4 spheres (kN): the
//    tests of a ray are straight-line code and interleave (-> 20.6 ms);
//  * with one light every frame's bounce ray leaves the same cached hit, so the origin-only part
//    of each test (oc = o0 - c, cc = |oc|^2 - r^2) is kept per pixel (kHoist, -> 19.4, 18.7 ms);
//  * G generates frames in PAIRS, side by side in one basic block: the hashes, samplers and sphere
//    tests of frame j and j + 1 are independent chains the scheduler interleaves, the pass overhead
//    is paid once per four frames, and a lane's first hit of a pair costs one push (-> 17.3 ms);
//  * less per-lane state (ptxas asks for 109 registers, five CTAs per SM allow 96): the reference's ray
//    count is derived from the traced count, the sample count is formed at retire, the per-pixel
//    counters live in shared memory, no common push after one-light G passes (-> 16.5 ms).
// Measured and not kept: a 32-frame window with two-bit codes for samples that end without a
// queued hit (fewer ring stalls - 29.2 lanes generating instead of 28.3 - but every sample then
// goes through the summing loop: 19.3 ms); the bounce pass tracing its shadow ray and its
// continuing ray together (+0.9 ms); per-pixel constants shuffled from the owning lane instead
// of queued (+0.2 ms); three pairs per pass, other bookkeeping periods and claim sizes (+-0.5 %).
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Per-pixel prologue of the warp-queue form with at most one light, as its own kernel.
//
// What every frame of a pixel shares there - primary ray, its hit, and the whole first bounce up to the roulette
// (hit record, emission, shadow ray, Cook-Torrance, the roulette probability and tangent frame of the first
// path_bounce: see megakernel_ww) - was computed by the lane that claimed the pixel. Claims come a few lanes at a
// time (lanes finish their pixels at different moments), so those ~700 instructions ran ~3 lanes wide: 2 % of a
// 1024-frame launch of config 2, but 16 % of a 128-frame one (one rank's share of an 8-GPU split: 3.37 ms where
// 2.83 would be an eighth of the full render). Here one thread per pixel of the launch's pool computes them 32 lanes
// wide and stores 96 B per pixel; the claiming lane loads them. Same functions on the same inputs: same bits.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_prologue_kernel(const RenderParams p)
{
    extern __shared__ float4 smem[];
    float4* sphS = smem;
    stage_spheres(sphS, p.spheres, p.nSpheres);
    __syncthreads();
    uint32_t traced = 0;
    for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < p.poolSize; id += gridDim.x * blockDim.x)
    {
        uint32_t x, y;
        if (!pool_pixel(p, id, x, y))
            continue;
        const uint32_t pixel = x + y * p.width;
        const V3 d0 = primary_direction(p.cam, x, y, p.width, p.height);
        PathState s;
        path_begin(s, p.cam.pos, d0, pixel, p.firstFrame);
        float tPrimary = 3.402823466e+38f;
        int cPrimary = -1;
        {
            const RayConst rk = ray_constants(s.dx, s.dy, s.dz);
            trace_range(sphS, p.nSpheres, 0u, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, rk, tPrimary, cPrimary);
            traced++;
        }
        float4* c = p.pixelCache + static_cast<size_t>(pixel) * kPrologueStride;
        if (cPrimary < 0)
        {
            c[0] = make_float4(tPrimary, __int_as_float(cPrimary), 0.0f, 0.0f);
            continue;
        }
        uint32_t raysPerStart = 1u;
        if (path_hit(p, s, sphS[cPrimary], cPrimary, tPrimary))
        {
            float tmin = 3.402823466e+38f;
            int closest = -1;
            const RayConst rk = ray_constants(s.dx, s.dy, s.dz);
            trace_range(sphS, p.nSpheres, 0u, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, rk, tmin, closest);
            traced++;
            path_shadow(p, s, closest, tmin);
            raysPerStart = 2u;
        }
        // the same instructions path_bounce runs, on the values every frame would feed it
        const float4 m0 = __ldg(p.mats + kMatStride * s.matIndex + 0);
        const float4 m1 = __ldg(p.mats + kMatStride * s.matIndex + 1);
        const float ax = fmul(1.0f, m0.x), ay = fmul(1.0f, m0.y), az = fmul(1.0f, m0.z);
        const float len = fsqrt_approx(fdot3(ax, ay, az, ax, ay, az));
        const float pr0 = fmax_(fmin_(len, 1.0f), 0.1f);
        const bool ggx0 = m1.w > 0.0f;
        const float ggxT0 = ggx0 ? __ldg(p.mats + kMatStride * s.matIndex + 5).x : 0.0f;
        V3 T0, B0;
        tangent_frame(s.N, T0, B0);
        c[0] = make_float4(tPrimary, __int_as_float(cPrimary), pr0, ggxT0);
        c[1] = make_float4(s.cr, s.cg, s.cb, __uint_as_float(raysPerStart | (ggx0 ? 0x100u : 0u)));
        c[2] = make_float4(s.ox, s.oy, s.oz, fdiv_approx(ax, pr0));
        c[3] = make_float4(s.N.x, s.N.y, s.N.z, fdiv_approx(ay, pr0));
        c[4] = make_float4(T0.x, T0.y, T0.z, fdiv_approx(az, pr0));
        c[5] = make_float4(B0.x, B0.y, B0.z, 0.0f);
    }
    // rays whose sphere loop ran (the roofline's count); the reference-equivalent count is kept by the render kernel
    if (p.counters)
    {
        unsigned long long t = traced;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31u) == 0)
            atomicAdd(p.counters + 2, t);
    }
}

#ifndef ATX_RING
#define ATX_RING 16
#endif
#ifndef ATX_WQ_CTAS
#define ATX_WQ_CTAS 5
#endif
#ifndef ATX_WQ_BOOKKEEP
#define ATX_WQ_BOOKKEEP 7u // retire/claim check every 8th pass of the warp loop (power of two minus one; measured: every pass 25.5 ms, 4th 24.7, 8th 24.5)
#endif
#ifndef ATX_WQ_CONSUME
#define ATX_WQ_CONSUME 3u // completed ring slots are added to the sums every (mask + 1)th pass (measured at one frame per pass: every pass 24.4 ms, 2nd 24.1, 4th 23.9, 8th 23.8)
#endif
#ifndef ATX_WQ_FLAT
#define ATX_WQ_FLAT 1 // sphere tests of the warp-queue form: 1 = the reference's sequence branch-free (exact_flat), 0 = line filter + hit branch
#endif
#ifndef ATX_WQ_HOIST
#define ATX_WQ_HOIST 1 // one light, compile-time sphere count: origin-only part of the G-phase sphere tests kept per pixel
#endif
#ifndef ATX_WQ_GPAIRS
#define ATX_WQ_GPAIRS 2u // pairs per pass of the warp loop (config 2: one pair 17.62 ms, two 17.29)
#endif
#ifndef ATX_WQ_BFULL
#define ATX_WQ_BFULL 32u // queued hits that trigger a bounce pass with no stall debt
#endif
constexpr uint32_t kRingFrames = ATX_RING;  // sample slots per pixel (power of two, <= 32: one done-bit each)
constexpr uint32_t kQueueCap = 64;    // hit-queue entries per warp (G runs below 32, B pops 32 and pushes <= 32)
constexpr uint32_t kQueueFields = 17;
constexpr uint32_t kWqWarps = 4;      // warps per CTA
constexpr uint32_t kWqWordsPerWarp = 3u * kRingFrames * 32u + 32u + kQueueFields * kQueueCap;

#ifdef ATX_WQ_STATS
// development build only: where the lanes of the warp-queue form go (printed by render_mega after each launch)
__device__ unsigned long long g_wqStats[8];
#define WQ_STAT(i, v) do { wqStat[i] += static_cast<unsigned long long>(v); } while (0)
#else
#define WQ_STAT(i, v) do { } while (0)
#endif

template <bool kFixedLight, int kN>
__global__ void __launch_bounds__(kWqWarps * 32, ATX_WQ_CTAS) megakernel_wq(const RenderParams p)
{
    // kN > 0: the scene has exactly kN spheres. The sphere tests are then straight-line code, and with one light the
    // part of a test that depends on the ray ORIGIN only is a per-pixel constant of the launch as well: every frame's
    // bounce ray leaves the cached first hit o0, so oc = o0 - c and cc = |oc|^2 - r^2 are formed once per claim
    // (same instructions on the same inputs as Renderer.cu:263-266) and a test in G is 3 + 15 instructions
    constexpr bool kHoist = kFixedLight && kN > 0 && ATX_WQ_HOIST != 0;
    constexpr int kH = kHoist ? kN : 1;
    extern __shared__ float4 smem[];
    float4* sphS = smem;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr uint32_t K = kRingFrames;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* warpWords = reinterpret_cast<uint32_t*>(smem + round_up8(p.nSpheres)) + (threadIdx.x >> 5) * kWqWordsPerWarp;
    float* ring = reinterpret_cast<float*>(warpWords);   // [3][K][32]: channel, slot, pixel lane
    uint32_t* done = warpWords + 3u * K * 32u;           // [32]: completed-slot bits per pixel lane
    uint32_t* q = done + 32u;                            // [kQueueFields][kQueueCap]
    float* qf = reinterpret_cast<float*>(q);

    unsigned long long* ctaCount = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint32_t*>(smem + round_up8(p.nSpheres)) + kWqWarps * kWqWordsPerWarp); // [0] paths, [1] raysFixed
    if (threadIdx.x < 2u)
        ctaCount[threadIdx.x] = 0ull;
    stage_spheres(sphS, p.spheres, p.nSpheres);
    __syncthreads();

    // `traced` counts sphere loops that ran. Every one of them in this kernel stands for one traceRay call of the reference;
    // the calls it does NOT run (the per-frame primary ray, the cached first shadow ray) are a per-pixel constant times
    // the frames of the pixel and are added once per pixel (raysFixed): rays = traced + raysFixed
    // (raysFixed and the path count change once per pixel: two CTA-wide words in shared memory instead of two registers
    // per lane - this kernel has no register to spare)
    uint32_t traced = 0;
#ifdef ATX_WQ_STATS
    unsigned long long wqStat[8] = {};
#endif

    auto trace_ray = [&](float ox, float oy, float oz, float dx, float dy, float dz, float& tmin, int& closest) {
        tmin = 3.402823466e+38f; // FLT_MAX
        closest = -1;
        const RayConst rk = ray_constants(dx, dy, dz);
        trace_range<ATX_WQ_FLAT != 0, kN>(sphS, p.nSpheres, 0u, ox, oy, oz, dx, dy, dz, rk, tmin, closest);
        traced++;
    };
    auto trace = [&](const PathState& s, float& tmin, int& closest) { trace_ray(s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, tmin, closest); };
    // a finished path hands its sample to the pixel's ring
    auto complete3 = [&](uint32_t tag, float cr, float cg, float cb) {
        const uint32_t pl = tag & 31u, slot = tag >> 8;
        ring[(0u * K + slot) * 32u + pl] = cr;
        ring[(1u * K + slot) * 32u + pl] = cg;
        ring[(2u * K + slot) * 32u + pl] = cb;
        atomicOr(done + pl, 1u << slot);
    };
    auto complete = [&](uint32_t tag, const PathState& s) { complete3(tag, s.cr, s.cg, s.cb); };

    // per-pixel state (lane = the pixel it owns until every frame of it is in the sum)
    bool live = false;      // owns a pixel with frames to render through G/B
    bool exhausted = false; // the pool has no more pixels
    uint32_t pixel = 0, j = 0, head = 0; // frames started / frames added to the sum
    // running sums of the pixel. The sample count (.w) is not carried: it grows by exactly 1.0 per frame, so at retire it
    // is the stored count + nFrames - formed in one addition where every intermediate count is an integer below 2^24 (the
    // additions are then exact, one by one or at once), one by one otherwise
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    V3 d0 = { 0.0f, 0.0f, 0.0f };
    float tPrimary = 0.0f;
    int cPrimary = -1;
    float c0r = 0.0f, c0g = 0.0f, c0b = 0.0f, o0x = 0.0f, o0y = 0.0f, o0z = 0.0f;
    V3 N0 = { 0.0f, 0.0f, 0.0f }, T0 = { 0.0f, 0.0f, 0.0f }, B0 = { 0.0f, 0.0f, 0.0f };
    float pr0 = 0.0f, tq0x = 0.0f, tq0y = 0.0f, tq0z = 0.0f, ggxT0 = 0.0f;
    bool ggx0 = false;
    float hocx[kH] = {}, hocy[kH] = {}, hocz[kH] = {}, hcc[kH] = {};
    uint32_t qHead = 0u, qCount = 0u, stallDebt = 0u, pass = 0u;
    done[lane] = 0u;
    __syncwarp();

    // a lane that owns no pixel takes the next one of the pool: primary ray, its hit, and with one light
    // the frame-independent first bounce. Claims are batched (claimThreshold idle lanes, or nothing else to
    // do) so this prologue runs with several lanes; the first claim of a warp is a whole 8x4 tile.
    auto claim = [&](uint32_t id) {
        uint32_t x, y;
        if (id >= p.poolSize)
        {
            exhausted = true;
            return;
        }
        if (!pool_pixel(p, id, x, y))
            return;
        pixel = x + y * p.width;
        acc = p.zeroFirst ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : p.accum[pixel];
        bool store = false;
        if (p.maxBounces < 1)
        {
            accumulate_black(acc, p.nFrames);
            store = true;
        }
        else if (kFixedLight)
        {
            // one light (or none): primary hit and the frame-independent first bounce were computed for every pixel of the
            // launch by pixel_prologue_kernel, 32 lanes wide; a claim (a few lanes at a time) only loads them
            const float4* c = p.pixelCache + static_cast<size_t>(pixel) * kPrologueStride;
            const float4 q0 = c[0];
            cPrimary = __float_as_int(q0.y);
            if (cPrimary < 0)
            {
                accumulate_sky(p, acc, p.nFrames);
                atomicAdd(ctaCount + 1, static_cast<unsigned long long>(p.nFrames));
                store = true;
            }
            else
            {
                const float4 q1 = c[1], q2 = c[2], q3 = c[3], q4 = c[4], q5 = c[5];
                tPrimary = q0.x; pr0 = q0.z; ggxT0 = q0.w;
                c0r = q1.x; c0g = q1.y; c0b = q1.z;
                const uint32_t flags = __float_as_uint(q1.w);
                ggx0 = (flags >> 8) != 0u;
                o0x = q2.x; o0y = q2.y; o0z = q2.z; tq0x = q2.w;
                N0 = { q3.x, q3.y, q3.z }; tq0y = q3.w;
                T0 = { q4.x, q4.y, q4.z }; tq0z = q4.w;
                B0 = { q5.x, q5.y, q5.z };
                if (kHoist)
                {
#pragma unroll
                    for (int i = 0; i < kH; i++)
                    {
                        const float4 sp = sphS[i];
                        hocx[i] = fadd(o0x, sp.x);
                        hocy[i] = fadd(o0y, sp.y);
                        hocz[i] = fadd(o0z, sp.z);
                        hcc[i] = ffma(fneg(sp.w), sp.w, fdot3(hocx[i], hocy[i], hocz[i], hocx[i], hocy[i], hocz[i]));
                    }
                }
                live = true;
                j = 0u;
                head = 0u;
            }
        }
        else
        {
            // the pixel's primary ray and its hit: the same for every frame (Camera.cpp:176-187)
            d0 = primary_direction(p.cam, x, y, p.width, p.height);
            PathState s;
            path_begin(s, p.cam.pos, d0, pixel, p.firstFrame);
            trace(s, tPrimary, cPrimary);
            atomicAdd(ctaCount + 1, ~0ull); // (this sphere loop is the first of the nFrames primary calls counted for the pixel)
            if (cPrimary < 0)
            {
                accumulate_sky(p, acc, p.nFrames);
                atomicAdd(ctaCount + 1, static_cast<unsigned long long>(p.nFrames));
                store = true;
            }
            else
            {
                live = true;
                j = 0u;
                head = 0u;
            }
        }
        if (store)
        {
            store_pixel(p, pixel, acc);
            if (p.emitRgba)
                p.rgba[pixel] = pack_rgba8(acc, u32_to_f32_rn(p.rgbaDivisor));
            atomicAdd(ctaCount, static_cast<unsigned long long>(p.nFrames));
        }
    };

    {
        while (true)
        {
            // ---- add completed samples in frame order ----
            if (live && head < j && (pass & ATX_WQ_CONSUME) == 0u)
            {
                uint32_t d = done[lane];
                while (head < j && ((d >> (head & (K - 1u))) & 1u))
                {
                    const uint32_t slot = head & (K - 1u);
                    acc.x = fadd(ring[(0u * K + slot) * 32u + lane], acc.x);
                    acc.y = fadd(ring[(1u * K + slot) * 32u + lane], acc.y);
                    acc.z = fadd(ring[(2u * K + slot) * 32u + lane], acc.z);
                    d &= ~(1u << slot);
                    head++;
                }
                done[lane] = d;
            }
            // retiring and claiming pixels is looked at every (ATX_WQ_BOOKKEEP + 1)th pass: a finished lane idles
            // for at most that many of its pixel's nFrames passes
            if ((pass++ & ATX_WQ_BOOKKEEP) == 0u)
            {
            if (live && head >= p.nFrames)
            {
                // every frame of the pixel is in the sum: st.global.v4.f32 + the display pack
                {
                    float w = p.zeroFirst ? 0.0f : p.accum[pixel].w;
                    const float n = u32_to_f32_rn(p.nFrames);
                    if (w >= 0.0f && w == floorf(w) && p.nFrames <= 16777216u && w + n <= 16777216.0f)
                        w = fadd(w, n);
                    else
                        for (uint32_t q = 0; q < p.nFrames; q++)
                            w = fadd(w, 1.0f);
                    acc.w = w;
                }
                store_pixel(p, pixel, acc);
                if (p.emitRgba)
                    p.rgba[pixel] = pack_rgba8(acc, u32_to_f32_rn(p.rgbaDivisor));
                atomicAdd(ctaCount, static_cast<unsigned long long>(p.nFrames));
                // reference calls per frame that the cached start stands for: the primary ray, and with one light the first shadow ray
                atomicAdd(ctaCount + 1, static_cast<unsigned long long>(p.nFrames) * (kFixedLight ? (__float_as_uint(p.pixelCache[static_cast<size_t>(pixel) * kPrologueStride + 1].w) & 0xffu) : 1u));
                live = false;
            }
            {
                const bool need = !live && !exhausted;
                const unsigned needMask = __ballot_sync(kFull, need);
                if (needMask != 0u &&
                    (static_cast<uint32_t>(__popc(needMask)) >= p.claimThreshold || (!__any_sync(kFull, live) && qCount == 0u)))
                {
                    const uint32_t id = pool_claim(p.pool, need);
                    if (need)
                        claim(id);
                }
            }
            }
            __syncwarp(); // the cleared bits are visible before another lane's bounce sets new ones
            const bool wants = live && j < p.nFrames;
            const bool gen = wants && (j - head) < K;
            const unsigned genMask = __ballot_sync(kFull, gen);
            const uint32_t nStalled = __popc(__ballot_sync(kFull, wants && !gen)); // ring full: waiting for an old path

            PathState s;
            bool wantPush = false;
            bool bouncePass = false; // (with one light G queues its hits itself: only a bounce pass reaches the common push below)
            float hitT = 0.0f;
            int hitC = -1;
            uint32_t tag = 0u;

            // When to bounce: at once with 32 queued hits (a full warp). With fewer, B wastes (32 - n)/32 of
            // its ~560 instructions, while every G iteration that passes with stalled lanes wastes 1/32 of
            // its ~290 per stalled lane: run B when the stall debt has reached what a narrow B would waste
            // (in lane-iterations: debt + 2n >= 64), the ski-rental rule, never worse than twice the optimum.
            if (qCount != 0u && (genMask == 0u || stallDebt + 2u * qCount >= 2u * ATX_WQ_BFULL))
            {
                stallDebt = 0u;
                bouncePass = true;
                WQ_STAT(4, 1); WQ_STAT(5, qCount < 32u ? qCount : 32u);
                // ---- B: one bounce for up to 32 queued hits ----
                const uint32_t n = qCount < 32u ? qCount : 32u;
                if (lane < n)
                {
                    const uint32_t e = (qHead + lane) & (kQueueCap - 1u);
                    s.ox = qf[0u * kQueueCap + e]; s.oy = qf[1u * kQueueCap + e]; s.oz = qf[2u * kQueueCap + e];
                    s.dx = qf[3u * kQueueCap + e]; s.dy = qf[4u * kQueueCap + e]; s.dz = qf[5u * kQueueCap + e];
                    s.cr = qf[6u * kQueueCap + e]; s.cg = qf[7u * kQueueCap + e]; s.cb = qf[8u * kQueueCap + e];
                    s.tx = qf[9u * kQueueCap + e]; s.ty = qf[10u * kQueueCap + e]; s.tz = qf[11u * kQueueCap + e];
                    s.seed = q[12u * kQueueCap + e];
                    s.bounce = static_cast<int>(q[13u * kQueueCap + e]);
                    const float t = qf[14u * kQueueCap + e];
                    const int c = static_cast<int>(q[15u * kQueueCap + e]);
                    tag = q[16u * kQueueCap + e];
                    if (path_hit(p, s, sphS[c], c, t))
                    {
                        float tmin;
                        int closest;
                        trace(s, tmin, closest);
                        path_shadow(p, s, closest, tmin);
                    }
                    bool ended = path_bounce(p, s);
                    if (!ended)
                    {
                        trace(s, hitT, hitC);
                        if (hitC >= 0)
                            wantPush = true;
                        else
                        {
                            path_miss(p, s);
                            ended = true;
                        }
                    }
                    if (ended)
                        complete(tag, s);
                }
                qHead = (qHead + n) & (kQueueCap - 1u);
                qCount -= n;
            }
            else if (genMask != 0u)
            {
                // ---- G: the next frame of this lane's pixel ----
                stallDebt += nStalled * (kFixedLight ? 2u * ATX_WQ_GPAIRS : 1u);
                WQ_STAT(0, 1); WQ_STAT(1, __popc(genMask)); WQ_STAT(2, nStalled);
                WQ_STAT(3, __popc(__ballot_sync(kFull, live && j >= p.nFrames)));
                if (!kFixedLight)
                {
                    if (gen)
                    {
                        // the path starts at the cached primary hit: straight to the bounce queue
                        tag = lane | ((j & (K - 1u)) << 8);
                        path_begin(s, p.cam.pos, d0, pixel, p.firstFrame + j * p.frameStride);
                        wantPush = true;
                        hitT = tPrimary;
                        hitC = cPrimary;
                        j++;
                    }
                }
                else
                {
                    // path_bounce (Renderer.cu:371-384) at bounce 0 with the per-pixel constants folded in. The path
                    // state of this branch IS the per-pixel constants plus a seed and a direction, so it is kept in
                    // its own registers and queued from them (no copy into a PathState)
                    // Two frames side by side in straight-line code: the three hashes, the sampler and the sphere tests
                    // of frame j and frame j + 1 are independent chains, so they fill each other's latencies (five warps
                    // per scheduler do not hide a serial PCG chain on their own). A lane whose path dies at the roulette
                    // computes the rest anyway - it would idle behind the others' branch otherwise.
#pragma unroll
                    for (uint32_t rep = 0u; rep < ATX_WQ_GPAIRS; rep++)
                    {
                        const uint32_t inFlight = j - head;
                        const bool onA = rep == 0u ? gen : (live && j < p.nFrames && inFlight < K);
                        bool onB = onA && j + 1u < p.nFrames && inFlight + 1u < K;
                        const uint32_t frameA = p.firstFrame + j * p.frameStride;
                        uint32_t seedA = pixel * frameA, seedB = pixel * (frameA + p.frameStride);
                        const float uA = pcg_float(seedA), uB = pcg_float(seedB);
                        float xA, yA, zA, xB, yB, zB;
                        sample_local(ggx0, ggxT0, seedA, xA, yA, zA);
                        sample_local(ggx0, ggxT0, seedB, xB, yB, zB);
                        const V3 ndA = frame_combine(N0, T0, B0, xA, yA, zA);
                        const V3 ndB = frame_combine(N0, T0, B0, xB, yB, zB);
                        seedA += 1u; // Renderer.cu:306
                        seedB += 1u;
                        const bool deep = 1 < p.maxBounces;
                        const bool flyA = onA && deep && !(uA > pr0);
                        bool flyB = onB && deep && !(uB > pr0);
                        float tA = 3.402823466e+38f, tB = 3.402823466e+38f; // FLT_MAX
                        int cA = -1, cB = -1;
                        if (kHoist)
                        {
                            const RayConst rkA = ray_constants(ndA.x, ndA.y, ndA.z), rkB = ray_constants(ndB.x, ndB.y, ndB.z);
#pragma unroll
                            for (int i = 0; i < kH; i++)
                            {
                                flat_tail(fdot3(hocx[i], hocy[i], hocz[i], ndA.x, ndA.y, ndA.z), hcc[i], i, rkA, tA, cA);
                                flat_tail(fdot3(hocx[i], hocy[i], hocz[i], ndB.x, ndB.y, ndB.z), hcc[i], i, rkB, tB, cB);
                            }
                        }
                        else
                        {
                            if (flyA)
                                trace_ray(o0x, o0y, o0z, ndA.x, ndA.y, ndA.z, tA, cA);
                            if (flyB)
                                trace_ray(o0x, o0y, o0z, ndB.x, ndB.y, ndB.z, tB, cB);
                        }
                        const bool hitA = flyA && cA >= 0;
                        bool hitB = flyB && cB >= 0;
                        // Queue room. A lane's first hit always fits (at most 32 queued when a pair starts, at most 32 first
                        // hits). A second hit that would not fit is taken back together with its frame: nothing of frame j + 1
                        // has been recorded yet, and the next pass generates it again from the same seed.
                        const unsigned mA = __ballot_sync(kFull, hitA), mB = __ballot_sync(kFull, hitB);
                        const unsigned m1 = mA | mB;
                        const uint32_t below = (1u << lane) - 1u;
                        const uint32_t n1 = __popc(m1);
                        if (hitA && hitB && static_cast<uint32_t>(__popc(mA & mB & below)) >= kQueueCap - qCount - n1)
                        {
                            if (!kHoist)
                                traced--; // (trace_ray has counted the sphere loop of the frame that is taken back)
                            onB = false;
                            flyB = false;
                            hitB = false;
                        }
                        const unsigned m2 = __ballot_sync(kFull, hitA && hitB);
                        const uint32_t tagA = lane | ((j & (K - 1u)) << 8), tagB = lane | (((j + 1u) & (K - 1u)) << 8);
                        if (kHoist)
                            traced += (flyA ? 1u : 0u) + (flyB ? 1u : 0u);
                        // samples that are complete, in frame order: the cached color (+ the sky seen by the bounce ray,
                        // path_miss); straight to the sum when nothing older is pending
                        if (onA && !hitA)
                        {
                            float cr = c0r, cg = c0g, cb = c0b;
                            if (flyA && p.skyLight)
                            {
                                cr = ffma(tq0x, 0.6f, cr);
                                cg = ffma(tq0y, 0.7f, cg);
                                cb = ffma(tq0z, 0.9f, cb);
                            }
                            if (head == j)
                            {
                                acc.x = fadd(cr, acc.x); acc.y = fadd(cg, acc.y); acc.z = fadd(cb, acc.z);
                                head++;
                            }
                            else
                                complete3(tagA, cr, cg, cb);
                        }
                        if (onB && !hitB)
                        {
                            float cr = c0r, cg = c0g, cb = c0b;
                            if (flyB && p.skyLight)
                            {
                                cr = ffma(tq0x, 0.6f, cr);
                                cg = ffma(tq0y, 0.7f, cg);
                                cb = ffma(tq0z, 0.9f, cb);
                            }
                            if (head == j + 1u)
                            {
                                acc.x = fadd(cr, acc.x); acc.y = fadd(cg, acc.y); acc.z = fadd(cb, acc.z);
                                head++;
                            }
                            else
                                complete3(tagB, cr, cg, cb);
                        }
                        j += (onA ? 1u : 0u) + (onB ? 1u : 0u);
                        if (hitA || hitB)
                        {
                            const uint32_t e = (qHead + qCount + __popc(m1 & below)) & (kQueueCap - 1u);
                            qf[0u * kQueueCap + e] = o0x; qf[1u * kQueueCap + e] = o0y; qf[2u * kQueueCap + e] = o0z;
                            qf[3u * kQueueCap + e] = hitA ? ndA.x : ndB.x; qf[4u * kQueueCap + e] = hitA ? ndA.y : ndB.y; qf[5u * kQueueCap + e] = hitA ? ndA.z : ndB.z;
                            qf[6u * kQueueCap + e] = c0r; qf[7u * kQueueCap + e] = c0g; qf[8u * kQueueCap + e] = c0b;
                            qf[9u * kQueueCap + e] = tq0x; qf[10u * kQueueCap + e] = tq0y; qf[11u * kQueueCap + e] = tq0z;
                            q[12u * kQueueCap + e] = hitA ? seedA : seedB;
                            q[13u * kQueueCap + e] = 1u; // bounce
                            qf[14u * kQueueCap + e] = hitA ? tA : tB;
                            q[15u * kQueueCap + e] = static_cast<uint32_t>(hitA ? cA : cB);
                            q[16u * kQueueCap + e] = hitA ? tagA : tagB;
                        }
                        qCount += n1;
                        if (m2 != 0u)
                        {
                            if (hitA && hitB)
                            {
                                const uint32_t e = (qHead + qCount + __popc(m2 & below)) & (kQueueCap - 1u);
                                qf[0u * kQueueCap + e] = o0x; qf[1u * kQueueCap + e] = o0y; qf[2u * kQueueCap + e] = o0z;
                                qf[3u * kQueueCap + e] = ndB.x; qf[4u * kQueueCap + e] = ndB.y; qf[5u * kQueueCap + e] = ndB.z;
                                qf[6u * kQueueCap + e] = c0r; qf[7u * kQueueCap + e] = c0g; qf[8u * kQueueCap + e] = c0b;
                                qf[9u * kQueueCap + e] = tq0x; qf[10u * kQueueCap + e] = tq0y; qf[11u * kQueueCap + e] = tq0z;
                                q[12u * kQueueCap + e] = seedB;
                                q[13u * kQueueCap + e] = 1u;
                                qf[14u * kQueueCap + e] = tB;
                                q[15u * kQueueCap + e] = static_cast<uint32_t>(cB);
                                q[16u * kQueueCap + e] = tagB;
                            }
                            qCount += __popc(m2);
                        }
                        if (qCount > 32u)
                            break;
                    }
                }
            }
            else if (__all_sync(kFull, exhausted && !live))
                break; // no pixel left anywhere in the warp
            else
                continue; // pixels were just retired or claimed: look again

            // ---- push the rays that hit (converged: every lane takes part in the ballot) ----
            if (!kFixedLight || bouncePass)
            {
            __syncwarp();
            const unsigned pushMask = __ballot_sync(kFull, wantPush);
            if (wantPush)
            {
                const uint32_t e = (qHead + qCount + __popc(pushMask & ((1u << lane) - 1u))) & (kQueueCap - 1u);
                qf[0u * kQueueCap + e] = s.ox; qf[1u * kQueueCap + e] = s.oy; qf[2u * kQueueCap + e] = s.oz;
                qf[3u * kQueueCap + e] = s.dx; qf[4u * kQueueCap + e] = s.dy; qf[5u * kQueueCap + e] = s.dz;
                qf[6u * kQueueCap + e] = s.cr; qf[7u * kQueueCap + e] = s.cg; qf[8u * kQueueCap + e] = s.cb;
                qf[9u * kQueueCap + e] = s.tx; qf[10u * kQueueCap + e] = s.ty; qf[11u * kQueueCap + e] = s.tz;
                q[12u * kQueueCap + e] = s.seed;
                q[13u * kQueueCap + e] = static_cast<uint32_t>(s.bounce);
                qf[14u * kQueueCap + e] = hitT;
                q[15u * kQueueCap + e] = static_cast<uint32_t>(hitC);
                q[16u * kQueueCap + e] = tag;
            }
            qCount += __popc(pushMask);
            }
            __syncwarp();
        }

    }
#ifdef ATX_WQ_STATS
    if (lane == 0u)
        for (int i = 0; i < 8; i++)
            atomicAdd(&g_wqStats[i], wqStat[i]);
#endif
    count_rays(p, traced, traced, 0u);
    __syncthreads();
    if (threadIdx.x == 0u && p.counters)
    {
        atomicAdd(p.counters + 0, ctaCount[0]);
        atomicAdd(p.counters + 1, ctaCount[1]); // (sums modulo 2^64: the -1 of a primary loop and the +nFrames of its pixel always meet in the same CTA)
    }
}

// ---------------------------------------------------------------------------
// Megakernel, two-slot packed form (large scenes: the sphere loop dominates).
//
// One thread = two path slots, each rendering a claimed pixel. The path loop of
// Renderer::perPixel is flattened: every iteration traces ONE ray per slot (primary,
// shadow or bounce ray, whichever that slot's path needs next) against all spheres with
// the packed f32x2 loop (trace_range2), then runs the matching half of the bounce per
// slot. A slot whose path ends adds its sample and starts the next frame's path from the
// cached primary hit in the same iteration; a slot whose pixel is complete stores it and
// claims another, so both lanes of every packed instruction carry a live ray until the
// pool is empty.
//
// kChunked: the sphere array does not fit the shared-memory budget; the CTA walks it in
// double-buffered chunks in lockstep, which needs the outer loop to be CTA-uniform
// (__syncthreads_or on "any thread still has work").
// ---------------------------------------------------------------------------
struct Slot
{
    PathState s;
    float4 acc;
    V3 d0;
    float tPrimary;
    int cPrimary;
    uint32_t pixel;
    uint32_t j;      // frames done
    uint32_t frame;
    bool alive;
    bool shadow;     // free-running form: the ray in flight is a shadow ray
    bool fresh;      // the ray in flight is the pixel's primary ray
    bool ray;        // lockstep form: the slot has a ray for the coming trace
};

__device__ __forceinline__ void slot_retire(const RenderParams& p, Slot& t, uint32_t& paths)
{
    store_pixel(p, t.pixel, t.acc);
    if (p.emitRgba)
        p.rgba[t.pixel] = pack_rgba8(t.acc, u32_to_f32_rn(p.rgbaDivisor));
    paths += t.j;
    t.alive = false;
}

// the slot takes pool pixel `id` (or learns that the pool is empty)
__device__ __forceinline__ void slot_claim(const RenderParams& p, Slot& t, uint32_t id, bool& exhausted, uint32_t& paths)
{
    uint32_t x, y;
    if (id >= p.poolSize)
    {
        exhausted = true;
        return;
    }
    if (!pool_pixel(p, id, x, y))
        return;
    t.pixel = x + y * p.width;
    t.acc = p.zeroFirst ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : p.accum[t.pixel];
    t.j = 0;
    t.frame = p.firstFrame;
    if (p.maxBounces < 1)
    {
        accumulate_black(t.acc, p.nFrames);
        t.j = p.nFrames;
        slot_retire(p, t, paths);
        return;
    }
    t.d0 = primary_direction(p.cam, x, y, p.width, p.height);
    path_begin(t.s, p.cam.pos, t.d0, t.pixel, t.frame);
    t.alive = true;
    t.fresh = true;
    t.shadow = false;
    t.ray = true;
}

// Start frame t.frame's path from the cached primary hit and run it up to its next trace:
// on return the slot has a shadow or bounce ray in flight, or has retired its pixel. (Loops
// only when paths end before any trace: no lights and roulette / bounce limit at bounce 0.)
template <bool kChunked>
__device__ __forceinline__ void slot_resume(const RenderParams& p, Slot& t, const float4* sphS, uint32_t& rays, uint32_t& paths)
{
    const float4 sp = __ldg(p.spheres + t.cPrimary); // shared memory holds the filter records; the exact record comes through L1
    while (true)
    {
        if (t.j >= p.nFrames)
        {
            slot_retire(p, t, paths);
            return;
        }
        path_begin(t.s, p.cam.pos, t.d0, t.pixel, t.frame);
        rays++; // the primary traceRay call this path starts with
        if (path_hit(p, t.s, sp, t.cPrimary, t.tPrimary))
        {
            t.shadow = true;
            return;
        }
        t.shadow = false;
        if (!path_bounce(p, t.s))
            return;
        accumulate_sample(t.acc, t.s);
        t.j++;
        t.frame += p.frameStride;
    }
}

template <bool kChunked>
__device__ __forceinline__ void slot_finish(const RenderParams& p, Slot& t, const float4* sphS, uint32_t& rays, uint32_t& paths)
{
    accumulate_sample(t.acc, t.s);
    t.j++;
    t.frame += p.frameStride;
    slot_resume<kChunked>(p, t, sphS, rays, paths);
}

// the half-bounce that follows the trace of this slot's ray
template <bool kChunked>
__device__ __forceinline__ void slot_advance(const RenderParams& p, Slot& t, const float4* sphS, int closest, float tmin,
                                             uint32_t& rays, uint32_t& paths)
{
    if (t.fresh)
    {
        // the primary ray of a newly claimed pixel: keep its hit for every frame
        t.fresh = false;
        t.tPrimary = tmin;
        t.cPrimary = closest;
        if (closest < 0)
        {
            accumulate_sky(p, t.acc, p.nFrames);
            rays += p.nFrames;
            t.j = p.nFrames;
            slot_retire(p, t, paths);
        }
        else
            slot_resume<kChunked>(p, t, sphS, rays, paths);
        return;
    }
    rays++;
    bool bounce = false;
    if (!t.shadow)
    {
        if (closest < 0)
        {
            path_miss(p, t.s);
            slot_finish<kChunked>(p, t, sphS, rays, paths);
            return;
        }
        const float4 sp = __ldg(p.spheres + closest);
        if (path_hit(p, t.s, sp, closest, tmin))
            t.shadow = true;
        else
            bounce = true;
    }
    else
    {
        path_shadow(p, t.s, closest, tmin);
        t.shadow = false;
        bounce = true;
    }
    if (bounce && path_bounce(p, t.s))
        slot_finish<kChunked>(p, t, sphS, rays, paths);
}

// one ray per slot against the whole scene
template <bool kChunked>
__device__ __forceinline__ void trace_pair(const RenderParams& p, const float4* sphS, uint32_t* candS, uint64_t* mbar, uint32_t& phase,
                                           const Slot& a, const Slot& b, bool live0, bool live1, float& tmin0, int& closest0, float& tmin1,
                                           int& closest1)
{
    tmin0 = tmin1 = 3.402823466e+38f; // FLT_MAX
    closest0 = closest1 = -1;
    const RayConst k0 = ray_constants(a.s.dx, a.s.dy, a.s.dz);
    const RayConst k1 = ray_constants(b.s.dx, b.s.dy, b.s.dz);
    RayPair rp;
    float od0, od1, g0, g1;
    ray_pair_lane(a.s.ox, a.s.oy, a.s.oz, a.s.dx, a.s.dy, a.s.dz, k0.a, od0, g0);
    ray_pair_lane(b.s.ox, b.s.oy, b.s.oz, b.s.dx, b.s.dy, b.s.dz, k1.a, od1, g1);
    rp.dx = pk2(a.s.dx, b.s.dx); rp.dy = pk2(a.s.dy, b.s.dy); rp.dz = pk2(a.s.dz, b.s.dz);
    rp.o2x = pk2(fadd(a.s.ox, a.s.ox), fadd(b.s.ox, b.s.ox));
    rp.o2y = pk2(fadd(a.s.oy, a.s.oy), fadd(b.s.oy, b.s.oy));
    rp.o2z = pk2(fadd(a.s.oz, a.s.oz), fadd(b.s.oz, b.s.oz));
    rp.od = pk2(od0, od1);
    rp.na = pk2(fneg(k0.a), fneg(k1.a));
    rp.g = pk2(g0, g1);
    if (!kChunked)
    {
        trace_range2(sphS, p.spheres, p.nSpheres, 0u, candS, rp, a.s, b.s, live0, live1, k0, k1, tmin0, closest0, tmin1, closest1);
    }
    else
    {
        // double-buffered chunk walk: one thread starts the bulk copy (TMA) of chunk c+1 while the CTA
        // traces chunk c; the copy signals an mbarrier, the CTA barrier at the end of an iteration says
        // "everyone is done with this buffer", which is what lets the next copy overwrite it
        const uint32_t C = p.chunkSpheres, stride = round_up32(C);
        const uint32_t nChunks = (p.nSpheres + C - 1) / C;
        float4* buf = const_cast<float4*>(sphS);
        if (threadIdx.x == 0)
            bulk_load(buf, p.sphFilter, min(C, p.nSpheres) * 16u, &mbar[0]);
        for (uint32_t c = 0; c < nChunks; c++)
        {
            const uint32_t cur = c & 1u;
            if (c + 1 < nChunks && threadIdx.x == 0)
                bulk_load(buf + (cur ^ 1u) * stride, p.sphFilter + (c + 1) * C, min(C, p.nSpheres - (c + 1) * C) * 16u, &mbar[cur ^ 1u]);
            mbar_wait(&mbar[cur], (phase >> cur) & 1u); // chunk c has landed
            phase ^= 1u << cur;
            if (live0 || live1)
                trace_range2(buf + cur * stride, p.spheres, min(C, p.nSpheres - c * C), c * C, candS, rp, a.s, b.s, live0, live1, k0, k1,
                             tmin0, closest0, tmin1, closest1);
            __syncthreads();
        }
    }
}

template <bool kChunked>
__global__ void __launch_bounds__(256, 2) megakernel_pair(const RenderParams p)
{
    extern __shared__ float4 smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);        // kChunked: one mbarrier per staging buffer
    uint32_t* candS = reinterpret_cast<uint32_t*>(smem + 1);   // candWords x blockDim.x candidate words
    float4* sphS = smem + 1 + cand_words(p) * 256u / 4u;
    constexpr unsigned kFull = 0xffffffffu;
    uint32_t phase = 0u; // parity of the next completion of each mbarrier

    if (!kChunked)
        stage_spheres(sphS, p.sphFilter, p.nSpheres, 32u);
    else
    {
        // the padded tail of a buffer is masked out of the candidate set, never tested; zero it once so
        // no uninitialised shared memory is ever read
        const uint32_t words = 2u * round_up32(p.chunkSpheres);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            sphS[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (threadIdx.x == 0)
        {
            mbar_init(&mbar[0], 1u);
            mbar_init(&mbar[1], 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    Slot a, b;
    a.alive = b.alive = false;
    a.fresh = b.fresh = false;
    a.shadow = b.shadow = false;
    a.j = b.j = 0;
    a.pixel = b.pixel = 0;
    a.acc = b.acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    a.d0 = b.d0 = { 0.0f, 0.0f, 0.0f };
    path_begin(a.s, p.cam.pos, a.d0, 0u, 0u);
    path_begin(b.s, p.cam.pos, b.d0, 0u, 0u);
    bool exhausted = false;
    uint32_t rays = 0, traced = 0, paths = 0;

    while (true)
    {
        // ---- claim pixels for idle slots ----
        const bool needA = !a.alive && !exhausted, needB = !b.alive && !exhausted;
        const uint32_t nNeed = __popc(__ballot_sync(kFull, needA)) + __popc(__ballot_sync(kFull, needB));
        const bool anyAlive = __any_sync(kFull, a.alive || b.alive);
        if (nNeed != 0u && (nNeed >= 2u * p.claimThreshold || !anyAlive))
        {
            const uint32_t idA = pool_claim(p.pool, needA);
            if (needA)
                slot_claim(p, a, idA, exhausted, paths);
            const uint32_t idB = pool_claim(p.pool, needB && !exhausted);
            if (needB && !exhausted)
                slot_claim(p, b, idB, exhausted, paths);
        }
        if (kChunked)
        {
            if (!__syncthreads_or(a.alive || b.alive || !exhausted))
                break;
        }
        else if (!__any_sync(kFull, a.alive || b.alive))
        {
            if (__all_sync(kFull, exhausted))
                break;
            continue;
        }

        // ---- trace the ray in flight of each slot against every sphere (Renderer::traceRay) ----
        float tmin0, tmin1;
        int closest0, closest1;
        trace_pair<kChunked>(p, sphS, candS, mbar, phase, a, b, a.alive, b.alive, tmin0, closest0, tmin1, closest1);
        if (a.alive)
        {
            traced++;
            slot_advance<kChunked>(p, a, sphS, closest0, tmin0, rays, paths);
        }
        if (b.alive)
        {
            traced++;
            slot_advance<kChunked>(p, b, sphS, closest1, tmin1, rays, paths);
        }
    }
    count_rays(p, rays, traced, paths);
}

// ---------------------------------------------------------------------------
// Megakernel, two-slot packed form in LOCKSTEP (large scenes, several frames per launch).
//
// In the free-running form above a slot traces whatever ray its path needs next, so after a while half the
// slots of a warp come back from a trace with a shadow-ray result (Cook-Torrance, roulette, next direction) and
// half with a closest-hit result (hit record, light pick): the code between two traces ran 6-14 lanes wide and
// was 20 % of all executed instructions on config 3 (profiles/r02a_ncu_full_megakernel_c3.txt). Here every
// iteration of the CTA makes the SAME two traces in the same order:
//   C  closest-hit rays of all slots (a newly claimed pixel's primary ray is one of them); then, for every slot
//      at once: a miss ends the path, adds the sample and restarts from the cached primary hit - and that
//      restart and a hit both continue with path_hit (hit record, emission, light pick), so the whole warp runs it
//      together;
//   S  shadow rays of all slots; then path_shadow + path_bounce for every slot at once.
// A path that ends in S (roulette, bounce limit: a few per cent of the paths) has no ray for the next C trace:
// its slot sits that one trace out and restarts with everybody else's path_hit. A newly claimed pixel waits at
// most one S trace for its primary ray. Everything else (pixel pool, packed filter + exact replay, sample order)
// is the free-running form's, and so are the results, bit for bit. Launches of fewer than kLockstepMinFrames
// frames keep the free-running form: there the one idle trace per pixel is not small against the pixel's work.
// ---------------------------------------------------------------------------
// closest-hit half: what follows the trace of this slot's closest-hit (or primary) ray
__device__ __forceinline__ void ls_closest(const RenderParams& p, Slot& t, int closest, float tmin, uint32_t& rays, uint32_t& traced,
                                           uint32_t& paths)
{
    if (!t.alive)
        return;
    bool restart = !t.ray; // the path ended in the shade half: start the next frame's path here
    if (t.ray)
    {
        traced++;
        if (t.fresh)
        {
            // the primary ray of a newly claimed pixel: its hit serves every frame (see slot_advance)
            t.fresh = false;
            t.tPrimary = tmin;
            t.cPrimary = closest;
            if (closest < 0)
            {
                accumulate_sky(p, t.acc, p.nFrames);
                rays += p.nFrames;
                t.j = p.nFrames;
                slot_retire(p, t, paths);
                t.ray = false;
                return;
            }
            restart = true;
        }
        else
        {
            rays++;
            if (closest < 0)
            {
                path_miss(p, t.s);
                accumulate_sample(t.acc, t.s);
                t.j++;
                restart = true;
            }
        }
    }
    if (restart)
    {
        if (t.j >= p.nFrames)
        {
            slot_retire(p, t, paths);
            t.ray = false;
            return;
        }
        path_begin(t.s, p.cam.pos, t.d0, t.pixel, p.firstFrame + t.j * p.frameStride);
        rays++; // the primary traceRay call this path starts with
        closest = t.cPrimary;
        tmin = t.tPrimary;
    }
    path_hit(p, t.s, __ldg(p.spheres + closest), closest, tmin);
    t.ray = true; // a shadow ray is in flight (without lights: the bounce is pending)
}

// shade half: what follows the trace of this slot's shadow ray
__device__ __forceinline__ void ls_shade(const RenderParams& p, Slot& t, int closest, float tmin, uint32_t& rays, uint32_t& traced,
                                         uint32_t& paths)
{
    if (!t.alive || !t.ray)
        return;
    if (p.nLights > 0)
    {
        traced++;
        rays++;
        path_shadow(p, t.s, closest, tmin);
    }
    if (path_bounce(p, t.s))
    {
        accumulate_sample(t.acc, t.s);
        t.j++;
        t.ray = false; // restarts with the next closest-hit half
        if (t.j >= p.nFrames)
            slot_retire(p, t, paths); // the pixel is complete: the slot claims a new one before the next trace
    }
}

#ifndef ATX_LS_THREADS
#define ATX_LS_THREADS 256
#endif
#ifndef ATX_LS_CTAS
#define ATX_LS_CTAS 2
#endif
constexpr uint32_t kLsThreads = ATX_LS_THREADS; // CTA size of the lockstep form
constexpr uint32_t kLsCtas = ATX_LS_CTAS;       // CTAs per SM its register budget allows

template <bool kChunked>
__global__ void __launch_bounds__(kLsThreads, kLsCtas) megakernel_pair_ls(const RenderParams p)
{
    extern __shared__ float4 smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);        // kChunked: one mbarrier per staging buffer
    uint32_t* candS = reinterpret_cast<uint32_t*>(smem + 1);   // candWords x blockDim.x candidate words
    float4* sphS = smem + 1 + cand_words(p) * kLsThreads / 4u;
    constexpr unsigned kFull = 0xffffffffu;
    uint32_t phase = 0u;

    if (!kChunked)
        stage_spheres(sphS, p.sphFilter, p.nSpheres, 32u);
    else
    {
        const uint32_t words = 2u * round_up32(p.chunkSpheres);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            sphS[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (threadIdx.x == 0)
        {
            mbar_init(&mbar[0], 1u);
            mbar_init(&mbar[1], 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    Slot a, b;
    a.alive = b.alive = false;
    a.fresh = b.fresh = false;
    a.shadow = b.shadow = false;
    a.ray = b.ray = false;
    a.j = b.j = 0;
    a.frame = b.frame = 0;
    a.pixel = b.pixel = 0;
    a.tPrimary = b.tPrimary = 0.0f;
    a.cPrimary = b.cPrimary = -1;
    a.acc = b.acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    a.d0 = b.d0 = { 0.0f, 0.0f, 0.0f };
    path_begin(a.s, p.cam.pos, a.d0, 0u, 0u);
    path_begin(b.s, p.cam.pos, b.d0, 0u, 0u);
    bool exhausted = false;
    uint32_t rays = 0, traced = 0, paths = 0;

    while (true)
    {
        // ---- claim pixels for idle slots: their primary rays join the closest-hit trace below ----
        const bool needA = !a.alive && !exhausted, needB = !b.alive && !exhausted;
        const uint32_t nNeed = __popc(__ballot_sync(kFull, needA)) + __popc(__ballot_sync(kFull, needB));
        const bool anyAlive = __any_sync(kFull, a.alive || b.alive);
        if (nNeed != 0u && (nNeed >= 2u * p.claimThreshold || !anyAlive))
        {
            const uint32_t idA = pool_claim(p.pool, needA);
            if (needA)
                slot_claim(p, a, idA, exhausted, paths);
            const uint32_t idB = pool_claim(p.pool, needB && !exhausted);
            if (needB && !exhausted)
                slot_claim(p, b, idB, exhausted, paths);
        }
        if (kChunked)
        {
            if (!__syncthreads_or(a.alive || b.alive || !exhausted))
                break;
        }
        else if (!__any_sync(kFull, a.alive || b.alive))
        {
            if (__all_sync(kFull, exhausted))
                break;
            continue;
        }

        float tmin0, tmin1;
        int closest0, closest1;
        // ---- C: closest-hit rays of every slot (Renderer::traceRay), then path_hit for all of them ----
        trace_pair<kChunked>(p, sphS, candS, mbar, phase, a, b, a.alive && a.ray, b.alive && b.ray, tmin0, closest0, tmin1, closest1);
        ls_closest(p, a, closest0, tmin0, rays, traced, paths);
        ls_closest(p, b, closest1, tmin1, rays, traced, paths);
        // ---- S: shadow rays of every slot, then Cook-Torrance, roulette and the next direction for all of them ----
        if (p.nLights > 0)
            trace_pair<kChunked>(p, sphS, candS, mbar, phase, a, b, a.alive && a.ray, b.alive && b.ray, tmin0, closest0, tmin1, closest1);
        ls_shade(p, a, closest0, tmin0, rays, traced, paths);
        ls_shade(p, b, closest1, tmin1, rays, traced, paths);
    }
    count_rays(p, rays, traced, paths);
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
                       uint32_t nL, float4* spheres, float4* sphFilter, int32_t* sphMat, float4* mats, float4* lights, cudaStream_t s)
{
    const uint32_t n = max(nS, max(nM, nL));
    if (n == 0)
        return cudaSuccess;
    pack_scene_kernel<<<(n + 255) / 256, 256, 0, s>>>(sphAoS, nS, matAoS, nM, lightAoS, nL, spheres, sphFilter, sphMat, mats, lights);
    return cudaGetLastError();
}

// which megakernel form a launch uses: the while-while form needs the whole scene resident
int mega_kind(const RenderParams& p, int requested)
{
    const bool chunked = p.chunkSpheres < p.nSpheres;
    // the packed form runs in lockstep (megakernel_pair_ls) when a launch has enough frames per pixel
    // and the scene is small enough for the shading between traces to matter: a path that ends in the shade half costs
    // the lockstep form one idle trace, which grows with the sphere count while the shading does not (measured, 4K:
    // 256 spheres / 16 lights +9.4 %, 4096 spheres / 4 lights -1.0 %)
    const int pairKind = (p.nFrames >= kLockstepMinFrames && p.nSpheres <= kLockstepMaxSpheres) ? kMegaPairLockstep : kMegaPair;
    if (chunked)
        return (requested == kMegaPair || requested == kMegaPairLockstep) ? requested : pairKind;
    if (requested == kMegaWhileWhile || requested == kMegaPair || requested == kMegaWarpQueue || requested == kMegaPairLockstep)
        return requested;
    if (p.nSpheres > kWhileWhileMaxSpheres)
        return pairKind;
    // The warp-queue form pays off where the first bounce is cached per pixel (at most one light) and a launch
    // has enough frames for the hit queue to fill. With several lights every frame starts with a full bounce, all
    // lanes need it at once, and the while-while form runs it without the trip through the queue (measured on 12
    // spheres / 3 lights, 1080p, 256 spp: while-while 14.2 ms, warp-queue 15.3 ms, two-slot packed 33.5 ms).
    return (p.nLights <= 1u && p.nFrames >= kWarpQueueMinFrames) ? kMegaWarpQueue : kMegaWhileWhile;
}

static size_t warp_queue_smem_bytes(const RenderParams& p)
{
    // spheres + ring, done bits and hit queue per warp + two CTA-wide counters
    return sizeof(float4) * ((static_cast<size_t>(p.nSpheres) + 7u) & ~size_t(7)) + sizeof(uint32_t) * (kWqWordsPerWarp * kWqWarps + 4u);
}

size_t megakernel_smem_bytes(const RenderParams& p)
{
    const bool chunked = p.chunkSpheres < p.nSpheres;
    const size_t pad = (static_cast<size_t>(chunked ? p.chunkSpheres : p.nSpheres) + 31u) & ~size_t(31); // whole blocks of 32 filter steps
    // the two-slot forms also keep cand_words() candidate words per thread (sized for the largest CTA of the forms)
    const uint32_t threads = kLsThreads > 256u ? kLsThreads : 256u;
    return sizeof(float4) * (1 + (chunked ? 2 * pad : pad)) + sizeof(uint32_t) * cand_words(p) * threads; // +16 B: two mbarriers
}

// the warp-queue kernel of a scene: one light or several, sphere count compiled in for the smallest scenes
typedef void (*WqKernel)(const RenderParams);
template <bool kFixedLight>
static WqKernel wq_kernel_by_count(uint32_t nSpheres)
{
    switch (nSpheres <= ATX_SMALL_STATIC ? nSpheres : 0u)
    {
#if ATX_SMALL_STATIC >= 1
    case 1: return megakernel_wq<kFixedLight, 1>;
#endif
#if ATX_SMALL_STATIC >= 2
    case 2: return megakernel_wq<kFixedLight, 2>;
#endif
#if ATX_SMALL_STATIC >= 3
    case 3: return megakernel_wq<kFixedLight, 3>;
#endif
#if ATX_SMALL_STATIC >= 4
    case 4: return megakernel_wq<kFixedLight, 4>;
#endif
    default: return megakernel_wq<kFixedLight, 0>;
    }
}
static WqKernel wq_kernel(const RenderParams& p)
{
    return p.nLights <= 1u ? wq_kernel_by_count<true>(p.nSpheres) : wq_kernel_by_count<false>(p.nSpheres);
}
template <bool kFixedLight>
static WqKernel ww_kernel_by_count(uint32_t nSpheres)
{
    switch (nSpheres <= ATX_SMALL_STATIC ? nSpheres : 0u)
    {
#if ATX_SMALL_STATIC >= 1
    case 1: return megakernel_ww<kFixedLight, 1>;
#endif
#if ATX_SMALL_STATIC >= 2
    case 2: return megakernel_ww<kFixedLight, 2>;
#endif
#if ATX_SMALL_STATIC >= 3
    case 3: return megakernel_ww<kFixedLight, 3>;
#endif
#if ATX_SMALL_STATIC >= 4
    case 4: return megakernel_ww<kFixedLight, 4>;
#endif
    default: return megakernel_ww<kFixedLight, 0>;
    }
}
static WqKernel ww_kernel(const RenderParams& p)
{
    return p.nLights <= 1u ? ww_kernel_by_count<true>(p.nSpheres) : ww_kernel_by_count<false>(p.nSpheres);
}
static cudaError_t configure_wq()
{
    for (uint32_t n = 0; n <= ATX_SMALL_STATIC; n++)
    {
        cudaError_t e = cudaFuncSetAttribute(wq_kernel_by_count<true>(n), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
        if (e != cudaSuccess)
            return e;
        e = cudaFuncSetAttribute(wq_kernel_by_count<false>(n), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
        if (e != cudaSuccess)
            return e;
        e = cudaFuncSetAttribute(ww_kernel_by_count<true>(n), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
        if (e != cudaSuccess)
            return e;
        e = cudaFuncSetAttribute(ww_kernel_by_count<false>(n), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

cudaError_t configure()
{
    cudaError_t e = cudaFuncSetAttribute(megakernel_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess)
        return e;
    e = cudaFuncSetAttribute(megakernel_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess)
        return e;
    e = cudaFuncSetAttribute(megakernel_pair_ls<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess)
        return e;
    e = cudaFuncSetAttribute(megakernel_pair_ls<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess)
        return e;
    e = configure_wq();
    if (e != cudaSuccess)
        return e;
    e = cudaFuncSetAttribute(pixel_prologue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess)
        return e;
    return configure_wavefront();
}

cudaError_t render_mega(const RenderParams& p, int kind, int smCount, cudaStream_t s)
{
    const bool chunked = p.chunkSpheres < p.nSpheres;
    const size_t smem = megakernel_smem_bytes(p);
    // persistent CTAs: as many as stay resident (3 or 2 per SM), fewer for tiny images
    const uint32_t byWork = (p.poolSize + 255u) / 256u;
    if (mega_kind(p, kind) == kMegaWarpQueue)
    {
        const size_t wq = warp_queue_smem_bytes(p);
        if (wq > static_cast<size_t>(kMaxSmemBytes))
            return cudaErrorInvalidValue;
        int perSm = 0;
        const WqKernel kernel = wq_kernel(p);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kWqWarps * 32, wq);
        if (e != cudaSuccess)
            return e;
        // no more CTAs than the image has 8x4 tiles per warp
        const uint32_t tiles = p.poolSize / 32u;
        const uint32_t grid = max(1u, min(static_cast<uint32_t>(smCount * max(perSm, 1)), (tiles + kWqWarps - 1u) / kWqWarps));
        if (p.nLights <= 1u)
        {
            if (!p.pixelCache)
                return cudaErrorInvalidValue;
            if (p.maxBounces >= 1)
            {
                // per-pixel constants of the launch, one thread per pixel of the pool (4 resident CTAs of 256 per SM is plenty)
                const size_t psm = sizeof(float4) * ((static_cast<size_t>(p.nSpheres) + 7u) & ~size_t(7));
                const uint32_t pgrid = max(1u, min(static_cast<uint32_t>(smCount) * 4u, (p.poolSize + 255u) / 256u));
                pixel_prologue_kernel<<<pgrid, 256, psm, s>>>(p);
            }
        }
        kernel<<<grid, kWqWarps * 32, wq, s>>>(p);
#ifdef ATX_WQ_STATS
        {
            unsigned long long st[8], zero[8] = {};
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(st, g_wqStats, sizeof(st));
            cudaMemcpyToSymbol(g_wqStats, zero, sizeof(zero));
            fprintf(stderr, "wq stats: G iters %llu, lanes generating %.2f, stalled %.2f, finished %.2f | B iters %llu, lanes %.2f\n", st[0],
                    double(st[1]) / double(st[0] ? st[0] : 1), double(st[2]) / double(st[0] ? st[0] : 1),
                    double(st[3]) / double(st[0] ? st[0] : 1), st[4], double(st[5]) / double(st[4] ? st[4] : 1));
        }
#endif
        return cudaGetLastError();
    }
    if (mega_kind(p, kind) == kMegaWhileWhile)
    {
        const uint32_t grid = min(static_cast<uint32_t>(smCount) * 3u, byWork);
        ww_kernel(p)<<<grid, 256, smem, s>>>(p);
    }
    else
    {
        const uint32_t grid = min(static_cast<uint32_t>(smCount) * 2u, (byWork + 1u) / 2u);
        const bool lockstep = mega_kind(p, kind) == kMegaPairLockstep;
        if (lockstep)
        {
            // persistent CTAs: as many as stay resident (registers: kLsCtas per SM; shared memory may allow fewer)
            int perSm = 0;
            cudaError_t e = chunked ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, megakernel_pair_ls<true>, kLsThreads, smem)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, megakernel_pair_ls<false>, kLsThreads, smem);
            if (e != cudaSuccess)
                return e;
            const uint32_t lsGrid = max(1u, min(static_cast<uint32_t>(smCount * max(perSm, 1)), (p.poolSize + 2u * kLsThreads - 1u) / (2u * kLsThreads)));
            if (chunked)
                megakernel_pair_ls<true><<<lsGrid, kLsThreads, smem, s>>>(p);
            else
                megakernel_pair_ls<false><<<lsGrid, kLsThreads, smem, s>>>(p);
        }
        else if (chunked)
            megakernel_pair<true><<<grid, 256, smem, s>>>(p);
        else
            megakernel_pair<false><<<grid, 256, smem, s>>>(p);
    }
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
