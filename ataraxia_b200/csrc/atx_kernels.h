// atx_kernels.h — host-side declarations of the kernel launchers (atx_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace atxk
{
struct RenderParams;
}

namespace atx_launch
{
// dynamic shared memory the megakernel may opt in to (227 KB is the sm_100 per-CTA limit;
// the default budget keeps two CTAs resident per SM)
constexpr int kMaxSmemBytes = 227 * 1024;
constexpr int kSmemBudgetTwoCtas = 72 * 1024; // sphere records; + 32 KB candidate words + 1 KB reserve, twice, fits 227 KB

// megakernel forms (ATX_TUNE_MEGA_KIND); 0 = by sphere count
constexpr int kMegaAuto = 0;
constexpr int kMegaWhileWhile = 1; // one pixel per thread, hits gathered before the shading phase
constexpr int kMegaPair = 2;       // two pixels per thread, packed f32x2 sphere loop
constexpr int kMegaWarpQueue = 3;  // one tile per warp, hits queued in shared memory and bounced 32 at a time
constexpr int kMegaPairLockstep = 4; // two-slot packed form, every slot of the CTA alternates closest-hit and shadow traces together
constexpr uint32_t kLockstepMinFrames = 4; // frames per launch from which the packed form runs in lockstep
constexpr uint32_t kLockstepMaxSpheres = 1536; // ... up to this many spheres (beyond, the free-running form wins)
constexpr uint32_t kWhileWhileMaxSpheres = 16;
constexpr uint32_t kWarpQueueMinFrames = 4;  // frames per launch from which the warp-queue form replaces the while-while form
                                             // (config 2 geometry, ms per launch, while-while / warp-queue: 1 frame 0.128 / 0.156,
                                             // 2: 0.169 / 0.179, 4: 0.257 / 0.237, 8: 0.428 / 0.324, 32: 1.36 / 1.10, 1024: 31.1 / 22.6)

cudaError_t configure();           // per atx_create, after cudaSetDevice: shared-memory opt-in of every kernel on that device
cudaError_t configure_wavefront(); // (atx_wavefront.cu's share of it)
int mega_kind(const atxk::RenderParams& p, int requested);
cudaError_t pack_scene(const float* sphAoS, uint32_t nS, const float* matAoS, uint32_t nM, const float* lightAoS,
                       uint32_t nL, float4* spheres, float4* sphFilter, int32_t* sphMat, float4* mats, float4* lights, cudaStream_t s);
size_t megakernel_smem_bytes(const atxk::RenderParams& p);
cudaError_t render_mega(const atxk::RenderParams& p, int kind, int smCount, cudaStream_t s);
// wavefront variant (atx_wavefront.cu)
constexpr int kWavefrontMaxBounces = 254; // one queue counter per bounce
size_t wavefront_bytes(uint32_t capacity);
uint32_t wavefront_frames_per_wave(uint32_t pixels, uint32_t nFrames);
cudaError_t render_wavefront(const atxk::RenderParams& p, void* work, uint32_t framesPerWave, uint64_t* launches, cudaStream_t s);
// cross-GPU sum of the accumulation buffers over peer memory (atx_p2p.cu)
constexpr uint32_t kP2pMaxRanks = 8;   // one NVSwitch domain
constexpr uint32_t kP2pStart = 0;      // flag block of a rank: [0, 8) "render complete" per peer, [8, 16) "stores landed" per peer,
constexpr uint32_t kP2pEnd = 8;        // [16] CTAs of the local kernel that are through
constexpr uint32_t kP2pCtaDone = 16;
constexpr uint32_t kP2pFlagWords = 32;
struct P2pParams
{
    float4* accum[kP2pMaxRanks];    // every rank's accumulation buffer as mapped into this process (own entry: the local pointer)
    uint32_t* flags[kP2pMaxRanks];  // every rank's flag block, likewise
    uint32_t* error;                // mapped host word: 1 = a peer never announced its buffer, 2 = a peer never finished
    uint32_t nRanks, rank, epoch;
    uint32_t count;                 // float4 elements
    unsigned long long timeoutNs;
};
cudaError_t p2p_allreduce(const P2pParams& q, int smCount, cudaStream_t s); // count == 0: the two flag barriers only
cudaError_t keep_own_tiles(float4* accum, uint32_t width, uint32_t height, uint32_t stride, uint32_t offset, cudaStream_t s);
cudaError_t primary_hits(const atxk::RenderParams& p, int32_t* out, cudaStream_t s);
cudaError_t ray_directions(const atxk::RenderParams& p, float* out, cudaStream_t s);
cudaError_t resolve_rgba(const float4* accum, uint32_t* rgba, uint32_t n, uint32_t divisor, cudaStream_t s);
}
