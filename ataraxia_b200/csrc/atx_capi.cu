// atx_capi.cu — the C-ABI (include/ataraxia_b200.h) over the sm_100a kernels.
//
// This is the headless host layer that replaces the reference's
// Renderer::Render choreography (Renderer.cu:173-249): no per-frame cudaMalloc,
// no per-frame W*H*12 B ray-table upload (Camera.cpp:197-210), no device-wide
// syncs; one non-default stream per handle; status codes instead of exit().
// There is no CPU fallback anywhere in this file: every render entry point
// launches CUDA kernels or fails.
#include "../../include/ataraxia_b200.h"
#include "../../include/ataraxia/Math.h"
#include "atx_device.cuh"
#include "atx_kernels.h"

#include <dlfcn.h>
#include <nccl.h> // types and enums only; the library is resolved at run time with dlopen

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace
{
thread_local std::string g_lastError;

atx_status fail(atx_status code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}

#define ATX_CUDA(call)                                                                                         \
    do                                                                                                         \
    {                                                                                                          \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return fail(ATX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- NCCL, resolved lazily so the library loads on hosts without it ----------
struct NcclApi
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string loadError;
};

void load_nccl(NcclApi& api)
{
    // a process that already carries NCCL (e.g. torch's bundled copy) gets that same
    // instance back from dlopen by soname; otherwise the system library is loaded
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names)
    {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib)
            break;
    }
    if (!api.lib)
    {
        const char* why = dlerror();
        api.loadError = why ? why : "libnccl.so.2 not found";
        return;
    }
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GetErrorString;
    if (!api.ok)
        api.loadError = "symbols missing";
}

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once; // handles on several devices may be driven from several threads
    std::call_once(once, load_nccl, std::ref(api));
    return api;
}

struct CameraState
{
    bool set = false;
    bool fromParams = false;
    float pos[3] = { 0, 0, 0 }, dir[3] = { 0, 0, -1 };
    float fov = 45.0f, nearClip = 0.1f, farClip = 100.0f;
    atx::mat4 invProj{ 1.0f }, invView{ 1.0f };
};

void camera_matrices(const float pos[3], const float dir[3], float fov, float nearClip, float farClip, uint32_t w,
                     uint32_t h, atx::mat4& proj, atx::mat4& view, atx::mat4& invProj, atx::mat4& invView)
{
    // Camera::UpdateProjectionMatrix (Camera.cpp:134-149) and UpdateViewMatrix (:151-159)
    const float aspect = static_cast<float>(w) / static_cast<float>(h);
    proj = atx::perspective(atx::radians(fov), aspect, nearClip, farClip);
    invProj = atx::inverse(proj);
    const atx::vec3 p(pos[0], pos[1], pos[2]), d(dir[0], dir[1], dir[2]);
    view = atx::lookAt(p, p + d, atx::vec3(0.0f, 1.0f, 0.0f));
    invView = atx::inverse(view);
}
} // namespace

struct atx_renderer
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    cudaEvent_t evUser[8] = {};
    bool timed = false;

    uint32_t width = 0, height = 0;
    float4* dAccum = nullptr;
    uint32_t* dRgba = nullptr;
    float4* dPreview = nullptr; // progressive multi-GPU preview: sum over ranks of the accumulation buffers, out of place
    int32_t* dHit = nullptr;   // lazily allocated debug buffers
    float* dRays = nullptr;
    unsigned long long* dCounters = nullptr;
    uint32_t* dPool = nullptr;  // pixel-pool counter of the persistent megakernels
    float4* dPixelCache = nullptr; // per-pixel launch constants of the warp-queue form (lazily allocated, 96 B per pixel)
    void* dWave = nullptr;      // wavefront variant: path records, samples, queues (lazily allocated)
    size_t waveBytes = 0;
    int autoVariant = ATX_VARIANT_MEGAKERNEL; // what ATX_VARIANT_AUTO resolves to (atx_calibrate)
    int smCount = 0;

    // scene
    float *dSphAoS = nullptr, *dMatAoS = nullptr, *dLightAoS = nullptr;
    float4 *dSpheres = nullptr, *dSphFilter = nullptr, *dMats = nullptr, *dLights = nullptr;
    int32_t* dSphMat = nullptr;
    size_t capS = 0, capM = 0, capL = 0;
    uint32_t nS = 0, nM = 0, nL = 0;
    std::vector<uint8_t> hostScene; // the records as uploaded (spheres, materials, lights): what a checkpoint's scene hash covers

    CameraState cam;
    bool accumulation = true, skyLight = false;
    int maxBounces = 15; // Settings default, Scene.h:53
    uint32_t frameIndex = 1;
    uint32_t lastFrame = 1;     // frameIndex of the last rendered frame (display divisor)
    uint32_t chunkOverride = 0; // tuning: force the shared-memory chunk size (spheres)
    int megaKind = 0;           // tuning: 0 = by sphere count and frames per launch, 1 = while-while, 2 = two-slot packed, 3 = warp-queue
    int lastKind = 0;           // form the last megakernel launch used
    uint32_t parkThreshold = 8;  // tuning: parked hits per warp that trigger the bounce phase (while-while form)
    uint32_t claimThreshold = 0; // tuning: idle lanes per warp that trigger a batched pixel claim (0 = per form)
    uint64_t launches = 0;
    bool calibrating = false;   // atx_calibrate's scratch launches: no counters

    ncclComm_t comm = nullptr;
    int nRanks = 1, rank = 0;

    // peer-memory sum of the accumulation buffers (atx_p2p.cu): every rank's buffer and flag block mapped here by CUDA IPC
    int reduceMode = 0;           // tuning: 0 = peer memory when it can be set up, 1 = ncclAllReduce
    bool p2pReady = false;        // the tables below describe the current dAccum of every rank
    bool p2pFailed = false;       // set-up was tried for this communicator and is not possible: NCCL from now on
    bool p2pExported = false;     // peers may hold a mapping of dAccum: it must outlive them (see atx_resize)
    float4* peerAccum[atx_launch::kP2pMaxRanks] = {};
    uint32_t* peerFlags[atx_launch::kP2pMaxRanks] = {};
    uint32_t* dP2pFlags = nullptr;
    uint8_t* dP2pStage = nullptr; // IPC handles on their way through ncclAllGather, and the agreement word
    uint32_t* hP2pError = nullptr; // mapped host memory
    uint32_t p2pEpoch = 0;
    int lastReduce = 0;           // 0 none yet, 1 peer memory, 2 NCCL
    std::vector<void*> retired;   // accumulation buffers replaced by atx_resize while peers could still have them mapped
};

namespace
{
void p2p_teardown(atx_handle h, bool freeRetired); // peer-memory reduce: defined with the multi-GPU entry points
atx_status p2p_check(atx_handle h);

atx_status make_params(atx_handle h, atxk::RenderParams& p)
{
    if (h->width == 0 || h->height == 0)
        return fail(ATX_ERR_INVALID, "atx_resize has not been called");
    if (!h->dAccum || !h->dRgba)
        return fail(ATX_ERR_INVALID, "no image buffers (a failed atx_resize leaves the handle without an image)");
    if (!h->cam.set)
        return fail(ATX_ERR_INVALID, "no camera: call atx_set_camera or atx_set_camera_matrices");
    std::memset(&p, 0, sizeof(p));
    p.width = h->width;
    p.height = h->height;
    p.maxBounces = h->maxBounces;
    p.skyLight = h->skyLight ? 1 : 0;
    p.nSpheres = h->nS;
    p.nMaterials = h->nM;
    p.nLights = h->nL;
    p.spheres = h->dSpheres;
    p.sphFilter = h->dSphFilter;
    p.sphMat = h->dSphMat;
    p.mats = h->dMats;
    p.lights = h->dLights;
    p.accum = h->dAccum;
    p.rgba = h->dRgba;
    p.counters = h->calibrating ? nullptr : h->dCounters; // calibration launches are not the caller's paths
    // shared-memory plan: whole scene if it fits the two-CTAs-per-SM budget, else chunks
    uint32_t chunk = h->nS;
    const uint32_t fit = atx_launch::kSmemBudgetTwoCtas / sizeof(float4);
    if (h->chunkOverride)
        chunk = h->chunkOverride;
    else if (h->nS > fit)
        chunk = fit / 2; // double-buffered
    p.chunkSpheres = chunk ? chunk : 1;
    p.parkThreshold = h->parkThreshold;
    p.nTiles = ((h->width + 7u) / 8u) * ((h->height + 3u) / 4u);
    p.tileStride = 1u;
    p.tileOffset = 0u;
    p.nPush = 0u;
    p.poolSize = p.nTiles * 32u;
    p.pool = h->dPool;
    const atx::mat4& ip = h->cam.invProj;
    const atx::mat4& iv = h->cam.invView;
    for (int i = 0; i < 4; i++)
    {
        p.cam.ip0[i] = ip[0][i];
        p.cam.ip1[i] = ip[1][i];
        // (m2*1 + m3*1): the constant half of glm's mat4*vec4 (type_mat4x4.inl:568-571)
        p.cam.ipA1[i] = ip[2][i] * 1.0f + ip[3][i] * 1.0f;
    }
    for (int i = 0; i < 3; i++)
    {
        p.cam.iv0[i] = iv[0][i];
        p.cam.iv1[i] = iv[1][i];
        p.cam.iv2[i] = iv[2][i];
        p.cam.iv3z[i] = iv[3][i] * 0.0f;
        p.cam.pos[i] = h->cam.pos[i];
    }
    return ATX_OK;
}

atx_status ensure_device(atx_handle h)
{
    if (!h)
        return fail(ATX_ERR_INVALID, "null handle");
    ATX_CUDA(cudaSetDevice(h->device));
    return ATX_OK;
}

template <typename T>
atx_status grow(T*& ptr, size_t& cap, size_t need, size_t elemsPer)
{
    if (need <= cap && ptr)
        return ATX_OK;
    if (ptr)
        ATX_CUDA(cudaFree(ptr));
    ptr = nullptr;
    const size_t n = std::max<size_t>(need, 1);
    ATX_CUDA(cudaMalloc(&ptr, n * elemsPer * sizeof(T)));
    cap = n;
    return ATX_OK;
}
} // namespace

extern "C" {

const char* atx_last_error(void) { return g_lastError.c_str(); }

const char* atx_version(void) { return "ataraxia_b200 0.1 (sm_100a; C-ABI 1)"; }

atx_status atx_create(int device_ordinal, atx_handle* out)
{
    if (!out)
        return fail(ATX_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ATX_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_ordinal < 0 || device_ordinal >= count)
        return fail(ATX_ERR_INVALID, "device ordinal %d out of range [0,%d)", device_ordinal, count);
    ATX_CUDA(cudaSetDevice(device_ordinal));
    atx_renderer* h = new (std::nothrow) atx_renderer();
    if (!h)
        return fail(ATX_ERR_ALLOC, "out of host memory");
    h->device = device_ordinal;
    ATX_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    ATX_CUDA(cudaEventCreate(&h->evStart));
    ATX_CUDA(cudaEventCreate(&h->evStop));
    ATX_CUDA(cudaMalloc(&h->dCounters, 4 * sizeof(unsigned long long)));
    ATX_CUDA(cudaMemsetAsync(h->dCounters, 0, 4 * sizeof(unsigned long long), h->stream));
    ATX_CUDA(cudaMalloc(&h->dPool, sizeof(uint32_t)));
    ATX_CUDA(cudaDeviceGetAttribute(&h->smCount, cudaDevAttrMultiProcessorCount, device_ordinal));
    ATX_CUDA(atx_launch::configure());
    *out = h;
    return ATX_OK;
}

atx_status atx_destroy(atx_handle h)
{
    if (!h)
        return ATX_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    p2p_teardown(h, true);
    if (h->comm && nccl().ok)
        nccl().CommDestroy(h->comm);
    cudaFree(h->dP2pFlags); cudaFree(h->dP2pStage);
    if (h->hP2pError)
        cudaFreeHost(h->hP2pError);
    cudaFree(h->dAccum); cudaFree(h->dRgba); cudaFree(h->dPreview); cudaFree(h->dHit); cudaFree(h->dRays); cudaFree(h->dCounters); cudaFree(h->dPool); cudaFree(h->dWave); cudaFree(h->dPixelCache);
    cudaFree(h->dSphAoS); cudaFree(h->dMatAoS); cudaFree(h->dLightAoS);
    cudaFree(h->dSpheres); cudaFree(h->dSphFilter); cudaFree(h->dMats); cudaFree(h->dLights); cudaFree(h->dSphMat);
    cudaEventDestroy(h->evStart); cudaEventDestroy(h->evStop);
    for (cudaEvent_t e : h->evUser)
        if (e)
            cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
    delete h;
    return ATX_OK;
}

atx_status atx_resize(atx_handle h, uint32_t width, uint32_t height)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (width == 0 || height == 0)
        return fail(ATX_ERR_INVALID, "width or height cannot be zero");
    if (h->dAccum && h->width == width && h->height == height)
        return ATX_OK; // Renderer.cu:100-101
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    // the size-dependent scratch goes first (it is re-created on demand), then the new image is allocated BEFORE the
    // old one is given up: a failed allocation leaves the handle exactly as it was (old size, old buffers)
    cudaFree(h->dPreview); cudaFree(h->dHit); cudaFree(h->dRays); cudaFree(h->dWave); cudaFree(h->dPixelCache);
    h->dPreview = nullptr; h->dHit = nullptr; h->dRays = nullptr; h->dWave = nullptr; h->dPixelCache = nullptr;
    h->waveBytes = 0;
    const size_t P = static_cast<size_t>(width) * height;
    float4* newAccum = nullptr;
    uint32_t* newRgba = nullptr;
    cudaError_t e = cudaMalloc(&newAccum, P * sizeof(float4));
    if (e == cudaSuccess)
        e = cudaMalloc(&newRgba, P * sizeof(uint32_t));
    if (e != cudaSuccess)
    {
        cudaFree(newAccum);
        cudaGetLastError(); // the failed allocation is reported here, not by the next launch
        return fail(ATX_ERR_ALLOC, "atx_resize(%u, %u): %s; the %ux%u image is kept", width, height, cudaGetErrorString(e),
                    h->width, h->height);
    }
    if (h->p2pExported && h->dAccum)
        h->retired.push_back(h->dAccum); // a peer may still have it mapped: freed with the communicator
    else
        cudaFree(h->dAccum);
    h->p2pReady = false; // the peers' tables describe the old buffer: set up again at the next reduce
    cudaFree(h->dRgba);
    h->dAccum = newAccum;
    h->dRgba = newRgba;
    ATX_CUDA(cudaMemsetAsync(h->dAccum, 0, P * sizeof(float4), h->stream));
    ATX_CUDA(cudaMemsetAsync(h->dRgba, 0, P * sizeof(uint32_t), h->stream));
    h->width = width;
    h->height = height;
    h->frameIndex = 1; // Renderer.cu:145
    h->lastFrame = 1;
    if (h->cam.set && h->cam.fromParams)
    {
        atx::mat4 proj, view;
        camera_matrices(h->cam.pos, h->cam.dir, h->cam.fov, h->cam.nearClip, h->cam.farClip, width, height, proj, view,
                        h->cam.invProj, h->cam.invView);
    }
    return ATX_OK;
}

atx_status atx_upload_scene(atx_handle h, const atx_sphere* spheres, size_t n_spheres, const atx_material* materials,
                            size_t n_materials, const atx_light* lights, size_t n_lights)
{
    if (atx_status s = ensure_device(h))
        return s;
    if ((n_spheres && !spheres) || (n_materials && !materials) || (n_lights && !lights))
        return fail(ATX_ERR_INVALID, "null array with non-zero count");
    if (n_spheres > 0x7fffffffu || n_materials > 0x7fffffffu || n_lights > 0x7fffffffu)
        return fail(ATX_ERR_INVALID, "scene too large");
    if (n_spheres && !n_materials)
        // the reference clamps every id to 0 and then reads materials[0] of an empty array (Renderer.cu:30-37, :326):
        // undefined there, refused here
        return fail(ATX_ERR_INVALID, "%zu spheres but no materials: every sphere needs a material to shade with", n_spheres);
    static_assert(sizeof(atx_sphere) == 20 && sizeof(atx_material) == 52 && sizeof(atx_light) == 28, "POD layout");
    size_t c;
    c = h->capS; if (atx_status s = grow(h->dSphAoS, c, n_spheres, 5)) return s;
    c = h->capS; if (atx_status s = grow(h->dSphMat, c, n_spheres, 1)) return s;
    c = h->capS; if (atx_status s = grow(h->dSphFilter, c, n_spheres, 1)) return s;
    if (atx_status s = grow(h->dSpheres, h->capS, n_spheres, 1)) return s;
    c = h->capM; if (atx_status s = grow(h->dMatAoS, c, n_materials, 13)) return s;
    if (atx_status s = grow(h->dMats, h->capM, n_materials, atxk::kMatStride)) return s;
    c = h->capL; if (atx_status s = grow(h->dLightAoS, c, n_lights, 7)) return s;
    if (atx_status s = grow(h->dLights, h->capL, n_lights, atxk::kLightStride)) return s;
    if (n_spheres)
        ATX_CUDA(cudaMemcpyAsync(h->dSphAoS, spheres, n_spheres * sizeof(atx_sphere), cudaMemcpyHostToDevice, h->stream));
    if (n_materials)
        ATX_CUDA(cudaMemcpyAsync(h->dMatAoS, materials, n_materials * sizeof(atx_material), cudaMemcpyHostToDevice, h->stream));
    if (n_lights)
        ATX_CUDA(cudaMemcpyAsync(h->dLightAoS, lights, n_lights * sizeof(atx_light), cudaMemcpyHostToDevice, h->stream));
    h->nS = static_cast<uint32_t>(n_spheres);
    h->nM = static_cast<uint32_t>(n_materials);
    h->nL = static_cast<uint32_t>(n_lights);
    try
    {
        h->hostScene.clear();
        const auto append = [&](const void* ptr, size_t bytes) {
            const uint8_t* b = static_cast<const uint8_t*>(ptr);
            h->hostScene.insert(h->hostScene.end(), b, b + bytes);
        };
        append(spheres, n_spheres * sizeof(atx_sphere));
        append(materials, n_materials * sizeof(atx_material));
        append(lights, n_lights * sizeof(atx_light));
    }
    catch (const std::bad_alloc&)
    {
        return fail(ATX_ERR_ALLOC, "out of host memory");
    }
    ATX_CUDA(atx_launch::pack_scene(h->dSphAoS, h->nS, h->dMatAoS, h->nM, h->dLightAoS, h->nL, h->dSpheres, h->dSphFilter, h->dSphMat,
                                    h->dMats, h->dLights, h->stream));
    h->launches++;
    // the caller's arrays may be pageable: the copies above must have consumed them before we return
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return ATX_OK;
}

atx_status atx_set_camera(atx_handle h, const float position[3], const float direction[3], float fov_degrees,
                          float near_clip, float far_clip)
{
    if (!h || !position || !direction)
        return fail(ATX_ERR_INVALID, "null argument");
    if (h->width == 0 || h->height == 0)
        return fail(ATX_ERR_INVALID, "atx_resize must precede atx_set_camera (the projection needs the aspect ratio)");
    std::memcpy(h->cam.pos, position, 12);
    std::memcpy(h->cam.dir, direction, 12);
    h->cam.fov = fov_degrees;
    h->cam.nearClip = near_clip;
    h->cam.farClip = far_clip;
    atx::mat4 proj, view;
    camera_matrices(h->cam.pos, h->cam.dir, fov_degrees, near_clip, far_clip, h->width, h->height, proj, view,
                    h->cam.invProj, h->cam.invView);
    h->cam.set = true;
    h->cam.fromParams = true;
    return ATX_OK;
}

atx_status atx_set_camera_matrices(atx_handle h, const float position[3], const float inv_projection[16],
                                   const float inv_view[16])
{
    if (!h || !position || !inv_projection || !inv_view)
        return fail(ATX_ERR_INVALID, "null argument");
    std::memcpy(h->cam.pos, position, 12);
    std::memcpy(&h->cam.invProj, inv_projection, 64);
    std::memcpy(&h->cam.invView, inv_view, 64);
    h->cam.set = true;
    h->cam.fromParams = false;
    return ATX_OK;
}

atx_status atx_set_settings(atx_handle h, int accumulation, int sky_light, int max_bounces)
{
    if (!h)
        return fail(ATX_ERR_INVALID, "null handle");
    h->accumulation = accumulation != 0;
    h->skyLight = sky_light != 0;
    h->maxBounces = max_bounces;
    return ATX_OK;
}

atx_status atx_set_tuning(atx_handle h, int key, int64_t value)
{
    if (!h)
        return fail(ATX_ERR_INVALID, "null handle");
    switch (key)
    {
    case ATX_TUNE_CHUNK_SPHERES:
        if (value < 0 || value > 7000)
            return fail(ATX_ERR_INVALID, "chunk_spheres must be in [0, 7000] (0 = automatic)");
        h->chunkOverride = static_cast<uint32_t>(value);
        return ATX_OK;
    case ATX_TUNE_MEGA_KIND:
        if (value < 0 || value > 4)
            return fail(ATX_ERR_INVALID, "mega_kind must be 0 (auto), 1 (while-while), 2 (two-slot packed), 3 (warp-queue) or 4 (two-slot packed, lockstep)");
        h->megaKind = static_cast<int>(value);
        return ATX_OK;
    case ATX_TUNE_PARK_THRESHOLD:
        if (value < 1 || value > 32)
            return fail(ATX_ERR_INVALID, "park_threshold must be in [1, 32]");
        h->parkThreshold = static_cast<uint32_t>(value);
        return ATX_OK;
    case ATX_TUNE_CLAIM_THRESHOLD:
        if (value < 0 || value > 32)
            return fail(ATX_ERR_INVALID, "claim_threshold must be in [0, 32] (0 = automatic)");
        h->claimThreshold = static_cast<uint32_t>(value);
        return ATX_OK;
    case ATX_TUNE_REDUCE:
        if (value < 0 || value > 1)
            return fail(ATX_ERR_INVALID, "reduce must be 0 (peer memory when possible) or 1 (ncclAllReduce)");
        h->reduceMode = static_cast<int>(value);
        return ATX_OK;
    default:
        return fail(ATX_ERR_INVALID, "unknown tuning key %d", key);
    }
}

atx_status atx_reset(atx_handle h)
{
    if (!h)
        return fail(ATX_ERR_INVALID, "null handle");
    h->frameIndex = 1;
    return ATX_OK;
}

atx_status atx_frame_index(atx_handle h, uint32_t* out)
{
    if (!h || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    *out = h->frameIndex;
    return ATX_OK;
}

static atx_status launch_frames(atx_handle h, uint32_t first, uint32_t n, uint32_t stride, bool zeroFirst, int variant,
                                bool emitRgba, uint32_t rgbaDivisor, uint32_t tileStride = 1, uint32_t tileOffset = 0, bool push = false)
{
    if (variant != ATX_VARIANT_AUTO && variant != ATX_VARIANT_MEGAKERNEL && variant != ATX_VARIANT_WAVEFRONT)
        return fail(ATX_ERR_INVALID, "unknown variant %d", variant);
    if (variant == ATX_VARIANT_AUTO)
        variant = h->autoVariant;
    atxk::RenderParams p;
    if (atx_status s = make_params(h, p))
        return s;
    p.firstFrame = first;
    p.nFrames = n;
    p.frameStride = stride;
    p.zeroFirst = zeroFirst ? 1 : 0;
    p.emitRgba = emitRgba ? 1 : 0;
    p.rgbaDivisor = rgbaDivisor;
    if (tileStride > 1u)
    {
        if (variant == ATX_VARIANT_WAVEFRONT)
            return fail(ATX_ERR_INVALID, "the image-tile split needs the megakernel's pixel pool");
        p.tileStride = tileStride;
        p.tileOffset = tileOffset;
        p.poolSize = ((p.nTiles + p.tileStride - 1u) / p.tileStride) * 32u;
        if (push && h->p2pReady)
            for (int r = 0; r < h->nRanks; r++)
                if (r != h->rank)
                    p.push[p.nPush++] = h->peerAccum[r];
    }
    if (variant == ATX_VARIANT_WAVEFRONT)
    {
        h->lastKind = 0;
        if (h->maxBounces > atx_launch::kWavefrontMaxBounces)
            return fail(ATX_ERR_INVALID, "the wavefront variant supports maxBounces <= %d", atx_launch::kWavefrontMaxBounces);
        if (static_cast<size_t>(h->nS) * sizeof(float4) > static_cast<size_t>(atx_launch::kMaxSmemBytes))
            return fail(ATX_ERR_INVALID, "the wavefront variant keeps the whole scene in shared memory: too many spheres");
        const uint32_t P = h->width * h->height;
        const uint32_t perWave = atx_launch::wavefront_frames_per_wave(P, n);
        const size_t need = atx_launch::wavefront_bytes(perWave * P);
        if (need > h->waveBytes)
        {
            ATX_CUDA(cudaStreamSynchronize(h->stream));
            if (h->dWave)
                ATX_CUDA(cudaFree(h->dWave));
            h->dWave = nullptr;
            h->waveBytes = 0;
            ATX_CUDA(cudaMalloc(&h->dWave, need));
            h->waveBytes = need;
        }
        ATX_CUDA(atx_launch::render_wavefront(p, h->dWave, perWave, &h->launches, h->stream));
        return ATX_OK;
    }
    if (atx_launch::megakernel_smem_bytes(p) > static_cast<size_t>(atx_launch::kMaxSmemBytes))
        return fail(ATX_ERR_INVALID, "shared-memory plan exceeds the device limit");
    // pixel claims: whole 8x4 tiles for the while-while form (its lockstep lives on coherent warps), small
    // batches for the packed form (the sphere loop does not care which pixels share a warp) and for the
    // warp-queue form (measured on config 2: 3-4 idle lanes per claim are best, 1 or 8 cost 1-2 %)
    const int formKind = atx_launch::mega_kind(p, h->megaKind);
    p.claimThreshold = h->claimThreshold ? h->claimThreshold
                                         : (formKind == atx_launch::kMegaWhileWhile ? 32u : formKind == atx_launch::kMegaWarpQueue ? 3u : 2u);
    h->lastKind = formKind;
    if (formKind == atx_launch::kMegaWarpQueue && p.nLights <= 1u)
    {
        if (!h->dPixelCache)
            ATX_CUDA(cudaMalloc(&h->dPixelCache, static_cast<size_t>(h->width) * h->height * atxk::kPrologueStride * sizeof(float4)));
        p.pixelCache = h->dPixelCache;
        if (p.maxBounces >= 1)
            h->launches++; // pixel_prologue_kernel
    }
    ATX_CUDA(cudaMemsetAsync(h->dPool, 0, sizeof(uint32_t), h->stream));
    ATX_CUDA(atx_launch::render_mega(p, h->megaKind, h->smCount, h->stream));
    h->launches++;
    return ATX_OK;
}

atx_status atx_render(atx_handle h, uint32_t n_frames, int variant)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (n_frames == 0)
        return ATX_OK;
    h->timed = false; // until this call has recorded both events
    ATX_CUDA(cudaEventRecord(h->evStart, h->stream));
    if (h->accumulation)
    {
        const uint32_t first = h->frameIndex;
        const uint32_t last = first + n_frames - 1;
        if (atx_status s = launch_frames(h, first, n_frames, 1, first == 1, variant, true, last))
            return s;
        h->lastFrame = last;
        h->frameIndex = last + 1; // Renderer.cu:245-246
    }
    else
    {
        // accumulation off: every frame restarts from a cleared buffer at frameIndex 1 (Renderer.cu:181-182, :247-248)
        for (uint32_t k = 0; k < n_frames; k++)
            if (atx_status s = launch_frames(h, 1, 1, 1, true, variant, true, 1))
                return s;
        h->lastFrame = 1;
        h->frameIndex = 1;
    }
    ATX_CUDA(cudaEventRecord(h->evStop, h->stream));
    h->timed = true;
    return ATX_OK;
}

atx_status atx_render_frames(atx_handle h, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride,
                             int zero_first, int variant)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (frame_stride == 0)
        return fail(ATX_ERR_INVALID, "frame_stride must be >= 1");
    if (!h->dAccum)
        return fail(ATX_ERR_INVALID, "atx_resize has not been called");
    h->timed = false;
    ATX_CUDA(cudaEventRecord(h->evStart, h->stream));
    if (n_frames == 0)
    {
        if (zero_first)
            ATX_CUDA(cudaMemsetAsync(h->dAccum, 0, static_cast<size_t>(h->width) * h->height * sizeof(float4), h->stream));
    }
    else
    {
        if (atx_status s = launch_frames(h, first_frame, n_frames, frame_stride, zero_first != 0, variant, false, 1))
            return s;
        h->lastFrame = first_frame + (n_frames - 1) * frame_stride;
    }
    ATX_CUDA(cudaEventRecord(h->evStop, h->stream));
    h->timed = true;
    return ATX_OK;
}

atx_status atx_calibrate(atx_handle h, uint32_t n_frames, float* megakernel_ms, float* wavefront_ms)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (n_frames == 0)
        return fail(ATX_ERR_INVALID, "n_frames must be >= 1");
    if (!h->dAccum)
        return fail(ATX_ERR_INVALID, "atx_resize has not been called");
    // render into a scratch buffer so the caller's accumulation is untouched
    float4* keep = h->dAccum;
    float4* scratch = nullptr;
    const size_t bytes = static_cast<size_t>(h->width) * h->height * sizeof(float4);
    ATX_CUDA(cudaMalloc(&scratch, bytes));
    h->dAccum = scratch;
    // the caller's counters, launch count and "last form" describe the caller's renders, not these
    h->calibrating = true;
    const uint64_t keepLaunches = h->launches;
    const int keepKind = h->lastKind;
    float ms[2] = { 0.0f, 0.0f };
    const int variants[2] = { ATX_VARIANT_MEGAKERNEL, ATX_VARIANT_WAVEFRONT };
    atx_status st = ATX_OK;
    for (int v = 0; v < 2 && st == ATX_OK; v++)
    {
        for (int rep = 0; rep < 2 && st == ATX_OK; rep++) // first pass warms caches and allocations
        {
            cudaEventRecord(h->evStart, h->stream);
            st = launch_frames(h, 1, n_frames, 1, true, variants[v], false, 1);
            cudaEventRecord(h->evStop, h->stream);
            if (st == ATX_OK && cudaEventSynchronize(h->evStop) != cudaSuccess)
                st = fail(ATX_ERR_CUDA, "calibration launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (st == ATX_OK)
                cudaEventElapsedTime(&ms[v], h->evStart, h->evStop);
        }
        if (st != ATX_OK && v == 1)
        {
            // the wavefront variant cannot run this configuration: the megakernel stays
            ms[1] = -1.0f;
            st = ATX_OK;
        }
    }
    h->dAccum = keep;
    h->calibrating = false;
    h->launches = keepLaunches;
    h->lastKind = keepKind;
    cudaStreamSynchronize(h->stream);
    cudaFree(scratch);
    h->timed = false;
    if (st != ATX_OK)
        return st;
    h->autoVariant = (ms[1] > 0.0f && ms[1] < ms[0]) ? ATX_VARIANT_WAVEFRONT : ATX_VARIANT_MEGAKERNEL;
    if (h->autoVariant == ATX_VARIANT_MEGAKERNEL && h->dWave)
    {
        // the wavefront scratch (5 float4 + 2 u32 per path) is only worth keeping for the variant that uses it
        cudaFree(h->dWave);
        h->dWave = nullptr;
        h->waveBytes = 0;
    }
    if (megakernel_ms) *megakernel_ms = ms[0];
    if (wavefront_ms) *wavefront_ms = ms[1];
    return ATX_OK;
}

atx_status atx_sync(atx_handle h)
{
    if (atx_status s = ensure_device(h))
        return s;
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return p2p_check(h);
}

atx_status atx_last_render_ms(atx_handle h, float* out_ms)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!out_ms)
        return fail(ATX_ERR_INVALID, "out_ms is null");
    if (!h->timed)
        return fail(ATX_ERR_INVALID, "nothing rendered yet");
    ATX_CUDA(cudaEventSynchronize(h->evStop));
    ATX_CUDA(cudaEventElapsedTime(out_ms, h->evStart, h->evStop));
    return ATX_OK;
}

atx_status atx_last_mega_kind(atx_handle h, int* out)
{
    if (!h || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    *out = h->lastKind;
    return ATX_OK;
}

atx_status atx_event_record(atx_handle h, int slot)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (slot < 0 || slot >= 8)
        return fail(ATX_ERR_INVALID, "event slot out of range");
    if (!h->evUser[slot])
        ATX_CUDA(cudaEventCreate(&h->evUser[slot]));
    ATX_CUDA(cudaEventRecord(h->evUser[slot], h->stream));
    return ATX_OK;
}

atx_status atx_event_elapsed_ms(atx_handle h, int slot_begin, int slot_end, float* out_ms)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!out_ms || slot_begin < 0 || slot_begin >= 8 || slot_end < 0 || slot_end >= 8 || !h->evUser[slot_begin] ||
        !h->evUser[slot_end])
        return fail(ATX_ERR_INVALID, "bad or unrecorded event slot");
    ATX_CUDA(cudaEventSynchronize(h->evUser[slot_end]));
    ATX_CUDA(cudaEventElapsedTime(out_ms, h->evUser[slot_begin], h->evUser[slot_end]));
    return ATX_OK;
}

atx_status atx_read_accum(atx_handle h, float* dst)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!dst || !h->dAccum)
        return fail(ATX_ERR_INVALID, "no destination or no image");
    const size_t bytes = static_cast<size_t>(h->width) * h->height * sizeof(float4);
    ATX_CUDA(cudaMemcpyAsync(dst, h->dAccum, bytes, cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return p2p_check(h);
}

atx_status atx_write_accum(atx_handle h, const float* src, uint32_t next_frame_index)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!src || !h->dAccum || next_frame_index == 0)
        return fail(ATX_ERR_INVALID, "no source, no image, or frame index 0");
    const size_t bytes = static_cast<size_t>(h->width) * h->height * sizeof(float4);
    ATX_CUDA(cudaMemcpyAsync(h->dAccum, src, bytes, cudaMemcpyHostToDevice, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    h->frameIndex = next_frame_index;
    h->lastFrame = next_frame_index > 1 ? next_frame_index - 1 : 1;
    return ATX_OK;
}

atx_status atx_read_rgba8(atx_handle h, uint32_t* dst, uint32_t divisor)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!dst || !h->dRgba)
        return fail(ATX_ERR_INVALID, "no destination or no image");
    const uint32_t n = h->width * h->height;
    if (divisor != 0)
    {
        // explicit divisor (e.g. total spp after a multi-GPU reduce): re-resolve on the device
        ATX_CUDA(atx_launch::resolve_rgba(h->dAccum, h->dRgba, n, divisor, h->stream));
        h->launches++;
    }
    ATX_CUDA(cudaMemcpyAsync(dst, h->dRgba, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return ATX_OK;
}

atx_status atx_host_alloc(size_t bytes, void** out)
{
    // page-locked host memory: a read-back into it is one DMA instead of a staged copy (the reference reads into
    // pageable new[] memory, Renderer.cu:124-129, :240)
    if (!out || bytes == 0)
        return fail(ATX_ERR_INVALID, "bad arguments");
    *out = nullptr;
    ATX_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return ATX_OK;
}

atx_status atx_host_free(void* p)
{
    if (p)
        ATX_CUDA(cudaFreeHost(p));
    return ATX_OK;
}

atx_status atx_read_hit_ids(atx_handle h, int32_t* dst)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!dst)
        return fail(ATX_ERR_INVALID, "dst is null");
    atxk::RenderParams p;
    if (atx_status s = make_params(h, p))
        return s;
    const size_t P = static_cast<size_t>(h->width) * h->height;
    if (!h->dHit)
        ATX_CUDA(cudaMalloc(&h->dHit, P * sizeof(int32_t)));
    ATX_CUDA(atx_launch::primary_hits(p, h->dHit, h->stream));
    h->launches++;
    ATX_CUDA(cudaMemcpyAsync(dst, h->dHit, P * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return ATX_OK;
}

atx_status atx_read_ray_directions(atx_handle h, float* dst)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!dst)
        return fail(ATX_ERR_INVALID, "dst is null");
    atxk::RenderParams p;
    if (atx_status s = make_params(h, p))
        return s;
    const size_t P = static_cast<size_t>(h->width) * h->height;
    if (!h->dRays)
        ATX_CUDA(cudaMalloc(&h->dRays, P * 3 * sizeof(float)));
    ATX_CUDA(atx_launch::ray_directions(p, h->dRays, h->stream));
    h->launches++;
    ATX_CUDA(cudaMemcpyAsync(dst, h->dRays, P * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return ATX_OK;
}

atx_status atx_get_counters(atx_handle h, atx_counters* out)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!out)
        return fail(ATX_ERR_INVALID, "out is null");
    unsigned long long c[4];
    ATX_CUDA(cudaMemcpyAsync(c, h->dCounters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    out->paths = c[0];
    out->rays = c[1];
    out->sphere_tests = c[1] * h->nS; // every traceRay tests every sphere (Renderer.cu:256)
    out->launches = h->launches;
    out->rays_traced = c[2];
    out->sphere_tests_executed = c[2] * h->nS;
    return ATX_OK;
}

atx_status atx_reset_counters(atx_handle h)
{
    if (atx_status s = ensure_device(h))
        return s;
    ATX_CUDA(cudaMemsetAsync(h->dCounters, 0, 4 * sizeof(unsigned long long), h->stream));
    h->launches = 0;
    return ATX_OK;
}

atx_status atx_accum_device_ptr(atx_handle h, void** out)
{
    if (!h || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    *out = h->dAccum;
    return ATX_OK;
}

atx_status atx_stream(atx_handle h, void** out_cuda_stream)
{
    if (!h || !out_cuda_stream)
        return fail(ATX_ERR_INVALID, "null argument");
    *out_cuda_stream = h->stream;
    return ATX_OK;
}

} // extern "C"

// ---- resumable renders on disk ---------------------------------
// The accumulation buffer plus the next frame index is the complete state of a progressive render: the RNG is a
// pure function of (pixel, frameIndex) (Renderer.cu:300-306) and the image is accumulation / frameIndex
// (Renderer.cu:165-168, :181-182, :245-248). A checkpoint is that state, bound to the scene and camera it was
// rendered with by a SHA-256 over the uploaded records and the camera matrices.
namespace
{
struct Sha256
{
    uint32_t h[8] = { 0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u };
    uint8_t block[64];
    size_t fill = 0;
    uint64_t total = 0;

    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

    void compress(const uint8_t* b)
    {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
            0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
            0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
            0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
            0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
            0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u };
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = (uint32_t(b[4 * i]) << 24) | (uint32_t(b[4 * i + 1]) << 16) | (uint32_t(b[4 * i + 2]) << 8) | uint32_t(b[4 * i + 3]);
        for (int i = 16; i < 64; i++)
        {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t v[8];
        std::memcpy(v, h, sizeof(v));
        for (int i = 0; i < 64; i++)
        {
            const uint32_t S1 = rotr(v[4], 6) ^ rotr(v[4], 11) ^ rotr(v[4], 25);
            const uint32_t ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
            const uint32_t t1 = v[7] + S1 + ch + K[i] + w[i];
            const uint32_t S0 = rotr(v[0], 2) ^ rotr(v[0], 13) ^ rotr(v[0], 22);
            const uint32_t maj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
            const uint32_t t2 = S0 + maj;
            v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = v[3] + t1;
            v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = t1 + t2;
        }
        for (int i = 0; i < 8; i++)
            h[i] += v[i];
    }

    void update(const void* data, size_t n)
    {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        total += n;
        while (n)
        {
            const size_t take = std::min(n, sizeof(block) - fill);
            std::memcpy(block + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == sizeof(block))
            {
                compress(block);
                fill = 0;
            }
        }
    }

    void finish(uint8_t out[32])
    {
        const uint64_t bits = total * 8u;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56)
            update(&zero, 1);
        uint8_t len[8];
        for (int i = 0; i < 8; i++)
            len[i] = static_cast<uint8_t>(bits >> (56 - 8 * i));
        update(len, 8);
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 4; j++)
                out[4 * i + j] = static_cast<uint8_t>(h[i] >> (24 - 8 * j));
    }
};

struct CheckpointHeader // 160 bytes, little-endian, no padding
{
    char magic[8];          // "ATXCKPT1"
    uint32_t version;       // 1
    uint32_t headerBytes;   // sizeof(CheckpointHeader)
    uint32_t width, height;
    uint32_t nextFrameIndex; // the frame index the render continues with
    uint32_t frameStride;    // 1, or the number of ranks of an spp-split render (this file is one rank's share)
    int32_t accumulation, skyLight, maxBounces;
    uint32_t nSpheres, nMaterials, nLights;
    int32_t rank, nRanks;    // communicator coordinates at save time (0, 1 without a communicator)
    uint64_t payloadBytes;   // width * height * 16
    uint8_t sceneSha256[32]; // uploaded scene records + camera position + inverse projection + inverse view
    uint8_t payloadSha256[32];
    uint8_t reserved[24];
};
static_assert(sizeof(CheckpointHeader) == 160, "checkpoint header layout");

void scene_digest(atx_handle h, uint8_t out[32])
{
    Sha256 sha;
    const uint32_t counts[3] = { h->nS, h->nM, h->nL };
    sha.update(counts, sizeof(counts));
    if (!h->hostScene.empty())
        sha.update(h->hostScene.data(), h->hostScene.size());
    sha.update(h->cam.pos, sizeof(h->cam.pos));
    sha.update(&h->cam.invProj, 64);
    sha.update(&h->cam.invView, 64);
    sha.finish(out);
}
} // namespace

extern "C" {

atx_status atx_host_sha256(const void* data, size_t bytes, uint8_t out[32])
{
    if ((!data && bytes) || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    Sha256 sha;
    sha.update(data, bytes);
    sha.finish(out);
    return ATX_OK;
}

atx_status atx_scene_sha256(atx_handle h, uint8_t out[32])
{
    if (!h || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    if (!h->cam.set)
        return fail(ATX_ERR_INVALID, "no camera: the scene hash covers the camera matrices");
    scene_digest(h, out);
    return ATX_OK;
}

atx_status atx_save_checkpoint(atx_handle h, const char* path, uint32_t next_frame_index, uint32_t frame_stride)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!path)
        return fail(ATX_ERR_INVALID, "path is null");
    if (!h->dAccum || !h->cam.set)
        return fail(ATX_ERR_INVALID, "nothing to save: atx_resize and a camera come first");
    const size_t P = static_cast<size_t>(h->width) * h->height;
    std::vector<float> acc;
    try
    {
        acc.resize(P * 4);
    }
    catch (const std::bad_alloc&)
    {
        return fail(ATX_ERR_ALLOC, "out of host memory");
    }
    ATX_CUDA(cudaMemcpyAsync(acc.data(), h->dAccum, P * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    CheckpointHeader hd;
    std::memset(&hd, 0, sizeof(hd));
    std::memcpy(hd.magic, "ATXCKPT1", 8);
    hd.version = 1;
    hd.headerBytes = sizeof(hd);
    hd.width = h->width;
    hd.height = h->height;
    hd.nextFrameIndex = next_frame_index ? next_frame_index : h->frameIndex;
    hd.frameStride = frame_stride ? frame_stride : 1u;
    hd.accumulation = h->accumulation;
    hd.skyLight = h->skyLight;
    hd.maxBounces = h->maxBounces;
    hd.nSpheres = h->nS; hd.nMaterials = h->nM; hd.nLights = h->nL;
    hd.rank = h->rank; hd.nRanks = h->nRanks;
    hd.payloadBytes = P * sizeof(float4);
    scene_digest(h, hd.sceneSha256);
    Sha256 sha;
    sha.update(acc.data(), hd.payloadBytes);
    sha.finish(hd.payloadSha256);
    // write to a temporary name and rename: a crash mid-write never leaves a half checkpoint under `path`
    const std::string tmp = std::string(path) + ".part";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f)
        return fail(ATX_ERR_INVALID, "cannot open %s for writing", tmp.c_str());
    const bool ok = std::fwrite(&hd, sizeof(hd), 1, f) == 1 && std::fwrite(acc.data(), 1, hd.payloadBytes, f) == hd.payloadBytes;
    const bool closed = std::fclose(f) == 0;
    if (!ok || !closed || std::rename(tmp.c_str(), path) != 0)
    {
        std::remove(tmp.c_str());
        return fail(ATX_ERR_INVALID, "writing %s failed", path);
    }
    return ATX_OK;
}

atx_status atx_load_checkpoint(atx_handle h, const char* path, uint32_t* next_frame_index, uint32_t* frame_stride)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!path)
        return fail(ATX_ERR_INVALID, "path is null");
    if (!h->dAccum || !h->cam.set)
        return fail(ATX_ERR_INVALID, "atx_resize, the scene and the camera come before atx_load_checkpoint (the file is checked against them)");
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return fail(ATX_ERR_INVALID, "cannot open %s", path);
    CheckpointHeader hd;
    atx_status st = ATX_OK;
    std::vector<float> acc;
    if (std::fread(&hd, sizeof(hd), 1, f) != 1 || std::memcmp(hd.magic, "ATXCKPT1", 8) != 0)
        st = fail(ATX_ERR_INVALID, "%s is not a checkpoint of this library", path);
    else if (hd.version != 1 || hd.headerBytes != sizeof(hd))
        st = fail(ATX_ERR_INVALID, "%s: unsupported checkpoint version %u", path, hd.version);
    else if (hd.width != h->width || hd.height != h->height)
        st = fail(ATX_ERR_INVALID, "%s holds a %ux%u image, the renderer is %ux%u", path, hd.width, hd.height, h->width, h->height);
    else if (hd.payloadBytes != static_cast<uint64_t>(h->width) * h->height * sizeof(float4) || hd.nextFrameIndex == 0 || hd.frameStride == 0)
        st = fail(ATX_ERR_INVALID, "%s: inconsistent header", path);
    else if (hd.maxBounces != h->maxBounces || (hd.skyLight != 0) != h->skyLight)
        st = fail(ATX_ERR_INVALID, "%s was rendered with maxBounces %d / skyLight %d, the renderer is set to %d / %d", path, hd.maxBounces,
                  hd.skyLight, h->maxBounces, int(h->skyLight));
    else
    {
        uint8_t now[32];
        scene_digest(h, now);
        if (std::memcmp(now, hd.sceneSha256, 32) != 0)
            st = fail(ATX_ERR_INVALID, "%s was rendered with a different scene or camera (scene hash mismatch): continuing it would mix two images", path);
    }
    if (st == ATX_OK)
    {
        try
        {
            acc.resize(hd.payloadBytes / sizeof(float));
        }
        catch (const std::bad_alloc&)
        {
            st = fail(ATX_ERR_ALLOC, "out of host memory");
        }
    }
    if (st == ATX_OK)
    {
        uint8_t extra;
        if (std::fread(acc.data(), 1, hd.payloadBytes, f) != hd.payloadBytes || std::fread(&extra, 1, 1, f) != 0)
            st = fail(ATX_ERR_INVALID, "%s is truncated or has trailing bytes", path);
        else
        {
            uint8_t digest[32];
            Sha256 sha;
            sha.update(acc.data(), hd.payloadBytes);
            sha.finish(digest);
            if (std::memcmp(digest, hd.payloadSha256, 32) != 0)
                st = fail(ATX_ERR_INVALID, "%s: payload checksum mismatch (corrupt file)", path);
        }
    }
    std::fclose(f);
    if (st != ATX_OK)
        return st;
    ATX_CUDA(cudaMemcpyAsync(h->dAccum, acc.data(), hd.payloadBytes, cudaMemcpyHostToDevice, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    h->frameIndex = hd.nextFrameIndex;
    h->lastFrame = hd.nextFrameIndex > 1 ? hd.nextFrameIndex - 1 : 1;
    if (next_frame_index) *next_frame_index = hd.nextFrameIndex;
    if (frame_stride) *frame_stride = hd.frameStride;
    return ATX_OK;
}

} // extern "C"

extern "C" {

// ---- multi-GPU ---------------------------------------------------------------

} // extern "C"

namespace
{
void p2p_teardown(atx_handle h, bool freeRetired)
{
    for (uint32_t r = 0; r < atx_launch::kP2pMaxRanks; r++)
    {
        if (static_cast<int>(r) != h->rank) // the own entries are the local pointers, not mappings
        {
            if (h->peerAccum[r])
                cudaIpcCloseMemHandle(h->peerAccum[r]);
            if (h->peerFlags[r])
                cudaIpcCloseMemHandle(h->peerFlags[r]);
        }
        h->peerAccum[r] = nullptr;
        h->peerFlags[r] = nullptr;
    }
    h->p2pReady = false;
    if (freeRetired)
    {
        for (void* p : h->retired)
            cudaFree(p);
        h->retired.clear();
        h->p2pExported = false;
    }
}

// COLLECTIVE over the communicator: map every rank's accumulation buffer and flag block into this process. All
// ranks leave with the same answer (an all-reduced "it worked for me" word): either every rank uses peer memory
// from now on or every rank uses ncclAllReduce.
atx_status p2p_setup(atx_handle h)
{
    using namespace atx_launch;
    p2p_teardown(h, false);
    h->p2pFailed = true; // until proven otherwise
    const int N = h->nRanks;
    int okLocal = (N >= 2 && N <= static_cast<int>(kP2pMaxRanks)) ? 1 : 0;
    constexpr size_t kRec = 2 * sizeof(cudaIpcMemHandle_t);
    if (!h->dP2pFlags)
    {
        ATX_CUDA(cudaMalloc(&h->dP2pFlags, kP2pFlagWords * sizeof(uint32_t)));
        ATX_CUDA(cudaMalloc(&h->dP2pStage, kRec * (kP2pMaxRanks + 1) + 16));
        void* e = nullptr;
        ATX_CUDA(cudaHostAlloc(&e, sizeof(uint32_t), cudaHostAllocMapped));
        h->hP2pError = static_cast<uint32_t*>(e);
        *h->hP2pError = 0u;
    }
    ATX_CUDA(cudaMemsetAsync(h->dP2pFlags, 0, kP2pFlagWords * sizeof(uint32_t), h->stream));
    h->p2pEpoch = 0;
    cudaIpcMemHandle_t mineRec[2];
    std::memset(mineRec, 0, sizeof(mineRec));
    if (okLocal && (cudaIpcGetMemHandle(&mineRec[0], h->dAccum) != cudaSuccess || cudaIpcGetMemHandle(&mineRec[1], h->dP2pFlags) != cudaSuccess))
    {
        cudaGetLastError();
        okLocal = 0;
    }
    h->p2pExported = true; // from here on a peer may map dAccum
    // every rank's two handles to every rank
    uint8_t* send = h->dP2pStage;
    uint8_t* recv = h->dP2pStage + kRec;
    ATX_CUDA(cudaMemcpyAsync(send, mineRec, kRec, cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = nccl().AllGather(send, recv, kRec, ncclUint8, h->comm, h->stream);
    if (r != ncclSuccess)
        return fail(ATX_ERR_NCCL, "ncclAllGather: %s", nccl().GetErrorString(r));
    std::vector<cudaIpcMemHandle_t> all(2 * static_cast<size_t>(std::max(N, 1)));
    if (N <= static_cast<int>(kP2pMaxRanks))
        ATX_CUDA(cudaMemcpyAsync(all.data(), recv, kRec * N, cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    for (int q = 0; q < N && okLocal; q++)
    {
        if (q == h->rank)
        {
            h->peerAccum[q] = h->dAccum;
            h->peerFlags[q] = h->dP2pFlags;
            continue;
        }
        void *a = nullptr, *f = nullptr;
        if (cudaIpcOpenMemHandle(&a, all[2 * q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&f, all[2 * q + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
            cudaGetLastError(); // e.g. two ranks in one process, or no peer access between the devices
            if (a)
                cudaIpcCloseMemHandle(a);
            okLocal = 0;
            break;
        }
        h->peerAccum[q] = static_cast<float4*>(a);
        h->peerFlags[q] = static_cast<uint32_t*>(f);
    }
    // agreement: min over ranks
    int* word = reinterpret_cast<int*>(h->dP2pStage + kRec * (kP2pMaxRanks + 1));
    ATX_CUDA(cudaMemcpyAsync(word, &okLocal, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    r = nccl().AllReduce(word, word, 1, ncclInt32, ncclMin, h->comm, h->stream);
    if (r != ncclSuccess)
        return fail(ATX_ERR_NCCL, "ncclAllReduce: %s", nccl().GetErrorString(r));
    int okAll = 0;
    ATX_CUDA(cudaMemcpyAsync(&okAll, word, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    if (!okAll)
    {
        p2p_teardown(h, false); // closes what this rank did open
        return ATX_OK;          // p2pFailed stays set: NCCL
    }
    h->p2pReady = true;
    h->p2pFailed = false;
    return ATX_OK;
}

unsigned long long p2p_timeout_ns()
{
    static const unsigned long long timeoutMs = []() {
        const char* e = std::getenv("ATX_P2P_TIMEOUT_MS");
        const long long v = e ? std::atoll(e) : 0;
        return static_cast<unsigned long long>(v > 0 ? v : 60000);
    }();
    return timeoutMs * 1000000ull;
}

// the peer-memory kernel on the handle's stream: flag barrier, sum of `count` float4 (0: none), flag barrier
atx_status p2p_enqueue(atx_handle h, uint32_t count)
{
    atx_launch::P2pParams q;
    std::memset(&q, 0, sizeof(q));
    for (int r = 0; r < h->nRanks; r++)
    {
        q.accum[r] = h->peerAccum[r];
        q.flags[r] = h->peerFlags[r];
    }
    ATX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&q.error), h->hP2pError, 0));
    q.nRanks = static_cast<uint32_t>(h->nRanks);
    q.rank = static_cast<uint32_t>(h->rank);
    q.epoch = ++h->p2pEpoch;
    q.count = count;
    q.timeoutNs = p2p_timeout_ns();
    ATX_CUDA(atx_launch::p2p_allreduce(q, h->smCount, h->stream));
    h->launches++;
    return ATX_OK;
}

atx_status p2p_check(atx_handle h)
{
    if (h->hP2pError && *h->hP2pError)
    {
        const uint32_t code = *h->hP2pError;
        *h->hP2pError = 0u;
        return fail(ATX_ERR_NCCL, "peer-memory reduce timed out (%s): a rank of the communicator did not take part in atx_allreduce_accum",
                    code == 1u ? "a peer never announced its buffer" : "a peer never finished its stores");
    }
    return ATX_OK;
}
} // namespace

extern "C" {

atx_status atx_comm_unique_id(uint8_t id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (!id)
        return fail(ATX_ERR_INVALID, "id is null");
    if (!nccl().ok)
        return fail(ATX_ERR_NCCL, "NCCL library not loadable: %s", nccl().loadError.c_str());
    ncclUniqueId uid;
    ncclResult_t r = nccl().GetUniqueId(&uid);
    if (r != ncclSuccess)
        return fail(ATX_ERR_NCCL, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
    std::memcpy(id, &uid, 128);
    return ATX_OK;
}

atx_status atx_comm_init_rank(atx_handle h, int n_ranks, int rank, const uint8_t id[128])
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(ATX_ERR_INVALID, "bad communicator arguments");
    if (!nccl().ok)
        return fail(ATX_ERR_NCCL, "NCCL library not loadable");
    if (h->comm)
    {
        cudaStreamSynchronize(h->stream);
        p2p_teardown(h, true);
        nccl().CommDestroy(h->comm);
        h->comm = nullptr;
    }
    h->p2pFailed = false;
    ncclUniqueId uid;
    std::memcpy(&uid, id, 128);
    ncclResult_t r = nccl().CommInitRank(&h->comm, n_ranks, uid, rank);
    if (r != ncclSuccess)
        return fail(ATX_ERR_NCCL, "ncclCommInitRank: %s", nccl().GetErrorString(r));
    h->nRanks = n_ranks;
    h->rank = rank;
    return ATX_OK;
}

atx_status atx_comm_destroy(atx_handle h)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (h->comm && nccl().ok)
    {
        cudaStreamSynchronize(h->stream);
        p2p_teardown(h, true);
        nccl().CommDestroy(h->comm);
    }
    h->comm = nullptr;
    h->nRanks = 1;
    h->rank = 0;
    return ATX_OK;
}

atx_status atx_allreduce_accum(atx_handle h)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!h->comm)
        return fail(ATX_ERR_INVALID, "no communicator: call atx_comm_init_rank");
    if (!h->dAccum)
        return fail(ATX_ERR_INVALID, "no image");
    if (atx_status s = p2p_check(h))
        return s;
    // one kernel over NVLink peer memory (atx_p2p.cu) when every rank could map every other rank's buffer;
    // the set-up is collective and happens at the first reduce after atx_comm_init_rank or atx_resize
    if (h->reduceMode == 0 && !h->p2pFailed && h->nRanks >= 2)
    {
        if (!h->p2pReady)
            if (atx_status s = p2p_setup(h))
                return s;
        if (h->p2pReady)
        {
            if (atx_status s = p2p_enqueue(h, h->width * h->height))
                return s;
            h->lastReduce = 1;
            return ATX_OK;
        }
    }
    const size_t count = static_cast<size_t>(h->width) * h->height * 4;
    ncclResult_t r = nccl().AllReduce(h->dAccum, h->dAccum, count, ncclFloat32, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess)
        return fail(ATX_ERR_NCCL, "ncclAllReduce: %s", nccl().GetErrorString(r));
    h->lastReduce = 2;
    return ATX_OK;
}

// Image-tile split (the alternative to the spp split): rank r of R renders ALL requested frames of the
// 8x4 tiles r, r + R, r + 2R, ... and its kernel stores every finished pixel straight into the image of every rank over
// NVLink peer memory; a flag barrier in peer memory ends the step. No arithmetic happens on the wire, the per-pixel
// prologue (primary ray, cached first bounce) is not replicated across ranks, and the sums of a pixel are formed on
// one GPU in frame order - so the image is bit-identical to a single-GPU render of the same frames, on every rank.
atx_status atx_render_tiles(atx_handle h, uint32_t first_frame, uint32_t n_frames, int zero_first, int variant)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!h->dAccum)
        return fail(ATX_ERR_INVALID, "atx_resize has not been called");
    if (n_frames == 0)
        return fail(ATX_ERR_INVALID, "n_frames must be >= 1");
    if (atx_status s = p2p_check(h))
        return s;
    h->timed = false;
    const bool multi = h->comm && h->nRanks >= 2;
    if (multi && h->reduceMode == 0 && !h->p2pFailed && !h->p2pReady)
        if (atx_status s = p2p_setup(h)) // collective
            return s;
    const bool peer = multi && h->reduceMode == 0 && h->p2pReady;
    ATX_CUDA(cudaEventRecord(h->evStart, h->stream));
    const size_t P = static_cast<size_t>(h->width) * h->height;
    if (multi && !peer)
    {
        // no peer mappings: keep only this rank's tiles (zero everywhere else), render, sum with ncclAllReduce: x + 0 == x,
        // so the result is the same bits, at the price of a full all-reduce
        if (zero_first)
            ATX_CUDA(cudaMemsetAsync(h->dAccum, 0, P * sizeof(float4), h->stream));
        else
        {
            ATX_CUDA(atx_launch::keep_own_tiles(h->dAccum, h->width, h->height, static_cast<uint32_t>(h->nRanks), static_cast<uint32_t>(h->rank), h->stream));
            h->launches++;
        }
    }
    if (peer)
        // peers store into this rank's image while they render: nobody starts before every rank's stream has reached this
        // step (e.g. has finished reading the previous image back)
        if (atx_status s = p2p_enqueue(h, 0))
            return s;
    if (atx_status s = launch_frames(h, first_frame, n_frames, 1, zero_first != 0, variant, false, 1, multi ? static_cast<uint32_t>(h->nRanks) : 1u,
                                     multi ? static_cast<uint32_t>(h->rank) : 0u, peer))
        return s;
    h->lastFrame = first_frame + n_frames - 1;
    if (peer)
    {
        // the kernel's peer stores are complete when it ends; tell every rank, and wait until every rank has told us
        if (atx_status s = p2p_enqueue(h, 0))
            return s;
        h->lastReduce = 1;
    }
    else if (multi)
    {
        ncclResult_t r = nccl().AllReduce(h->dAccum, h->dAccum, P * 4, ncclFloat32, ncclSum, h->comm, h->stream);
        if (r != ncclSuccess)
            return fail(ATX_ERR_NCCL, "ncclAllReduce: %s", nccl().GetErrorString(r));
        h->lastReduce = 2;
    }
    ATX_CUDA(cudaEventRecord(h->evStop, h->stream));
    h->timed = true;
    return ATX_OK;
}

// one share of an image-tile split, locally: the tiles share, share + n_shares, ... of the image; every other pixel of the
// accumulation buffer is left alone
atx_status atx_render_tile_share(atx_handle h, uint32_t first_frame, uint32_t n_frames, int zero_first, int variant, uint32_t n_shares,
                                 uint32_t share)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!h->dAccum)
        return fail(ATX_ERR_INVALID, "atx_resize has not been called");
    if (n_frames == 0 || n_shares == 0 || share >= n_shares)
        return fail(ATX_ERR_INVALID, "n_frames >= 1 and share < n_shares are required");
    h->timed = false;
    ATX_CUDA(cudaEventRecord(h->evStart, h->stream));
    if (atx_status s = launch_frames(h, first_frame, n_frames, 1, zero_first != 0, variant, false, 1, n_shares, share, false))
        return s;
    h->lastFrame = first_frame + n_frames - 1;
    ATX_CUDA(cudaEventRecord(h->evStop, h->stream));
    h->timed = true;
    return ATX_OK;
}

atx_status atx_last_reduce_kind(atx_handle h, int* out)
{
    if (!h || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    *out = h->lastReduce;
    return ATX_OK;
}

// Progressive preview across ranks: the sum of every rank's accumulation buffer WITHOUT touching
// the buffers themselves, so each rank keeps adding its own frames afterwards. One out-of-place ncclAllReduce
// into a second float4 buffer (a plain device copy when the handle has no communicator).
atx_status atx_allreduce_preview(atx_handle h)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!h->dAccum)
        return fail(ATX_ERR_INVALID, "no image");
    const size_t P = static_cast<size_t>(h->width) * h->height;
    if (!h->dPreview)
        ATX_CUDA(cudaMalloc(&h->dPreview, P * sizeof(float4)));
    if (!h->comm)
    {
        ATX_CUDA(cudaMemcpyAsync(h->dPreview, h->dAccum, P * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
        return ATX_OK;
    }
    ncclResult_t r = nccl().AllReduce(h->dAccum, h->dPreview, P * 4, ncclFloat32, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess)
        return fail(ATX_ERR_NCCL, "ncclAllReduce: %s", nccl().GetErrorString(r));
    return ATX_OK;
}

atx_status atx_read_preview(atx_handle h, float* accum_dst, uint32_t* rgba_dst, uint32_t divisor)
{
    if (atx_status s = ensure_device(h))
        return s;
    if (!h->dPreview)
        return fail(ATX_ERR_INVALID, "no preview: call atx_allreduce_preview");
    if (rgba_dst && divisor == 0)
        return fail(ATX_ERR_INVALID, "the preview needs the total sample count as divisor");
    const size_t P = static_cast<size_t>(h->width) * h->height;
    if (accum_dst)
        ATX_CUDA(cudaMemcpyAsync(accum_dst, h->dPreview, P * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    if (rgba_dst)
    {
        // the display buffer is scratch between renders: the next render or atx_read_rgba8 rewrites it
        ATX_CUDA(atx_launch::resolve_rgba(h->dPreview, h->dRgba, static_cast<uint32_t>(P), divisor, h->stream));
        h->launches++;
        ATX_CUDA(cudaMemcpyAsync(rgba_dst, h->dRgba, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    ATX_CUDA(cudaStreamSynchronize(h->stream));
    return ATX_OK;
}

// ---- host math helpers for the header-only C++ mirror -------------------------
// Compiled here (not in the caller) so the caller's compiler flags cannot change
// the numerics that parity depends on.

atx_status atx_host_camera_matrices(const float position[3], const float direction[3], float fov_degrees,
                                    float near_clip, float far_clip, uint32_t width, uint32_t height,
                                    float projection[16], float view[16], float inv_projection[16], float inv_view[16])
{
    if (!position || !direction || width == 0 || height == 0)
        return fail(ATX_ERR_INVALID, "bad camera arguments");
    atx::mat4 p, v, ip, iv;
    camera_matrices(position, direction, fov_degrees, near_clip, far_clip, width, height, p, v, ip, iv);
    if (projection) std::memcpy(projection, &p, 64);
    if (view) std::memcpy(view, &v, 64);
    if (inv_projection) std::memcpy(inv_projection, &ip, 64);
    if (inv_view) std::memcpy(inv_view, &iv, 64);
    return ATX_OK;
}

atx_status atx_host_ray_directions(const float inv_projection[16], const float inv_view[16], uint32_t width,
                                   uint32_t height, float* out)
{
    // Camera::UpdateRayDirection (Camera.cpp:161-195) for callers of Camera::getRayDirection():
    // the renderer itself never reads this table — it generates rays in-kernel.
    if (!inv_projection || !inv_view || !out || width == 0 || height == 0)
        return fail(ATX_ERR_INVALID, "bad arguments");
    atx::mat4 ip, iv;
    std::memcpy(&ip, inv_projection, 64);
    std::memcpy(&iv, inv_view, 64);
    const int nThreads = std::max(1u, std::min(std::thread::hardware_concurrency(), height));
    std::vector<std::thread> threads;
    const uint32_t rows = height / nThreads;
    for (int t = 0; t < nThreads; t++)
    {
        const uint32_t y0 = t * rows, y1 = (t == nThreads - 1) ? height : y0 + rows;
        threads.emplace_back([=]() {
            for (uint32_t y = y0; y < y1; y++)
                for (uint32_t x = 0; x < width; x++)
                {
                    atx::vec2 coord(static_cast<float>(x) / static_cast<float>(width),
                                    static_cast<float>(y) / static_cast<float>(height));
                    coord = coord * 2.0f - 1.0f;
                    const atx::vec4 target = ip * atx::vec4(coord.x, coord.y, 1.0f, 1.0f);
                    const atx::vec3 n = atx::normalize(atx::vec3(target.x, target.y, target.z) / target.w);
                    const atx::vec4 r = iv * atx::vec4(n, 0.0f);
                    const atx::vec3 d = atx::normalize(atx::vec3(r.x, r.y, r.z));
                    float* o = out + 3ull * (x + static_cast<size_t>(y) * width);
                    o[0] = d.x; o[1] = d.y; o[2] = d.z;
                }
        });
    }
    for (auto& th : threads)
        th.join();
    return ATX_OK;
}

atx_status atx_host_node_transform(const float parent[16], const float position[3], const float rotation_xyzw[4],
                                   const float scale[3], float local[16], float global[16])
{
    // SceneNode::updateGlobalTransform (SceneNode.cpp:42-59)
    if (!parent || !position || !rotation_xyzw || !scale)
        return fail(ATX_ERR_INVALID, "null argument");
    atx::mat4 par;
    std::memcpy(&par, parent, 64);
    const atx::quat q(rotation_xyzw[3], rotation_xyzw[0], rotation_xyzw[1], rotation_xyzw[2]);
    const atx::mat4 loc = atx::translate(atx::mat4(1.0f), atx::vec3(position[0], position[1], position[2])) *
                          atx::mat4_cast(q) * atx::scale(atx::mat4(1.0f), atx::vec3(scale[0], scale[1], scale[2]));
    const atx::mat4 glob = par * loc;
    if (local) std::memcpy(local, &loc, 64);
    if (global) std::memcpy(global, &glob, 64);
    return ATX_OK;
}

atx_status atx_host_mat4_mul(const float a[16], const float b[16], float out[16])
{
    if (!a || !b || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    atx::mat4 A, B;
    std::memcpy(&A, a, 64);
    std::memcpy(&B, b, 64);
    const atx::mat4 R = A * B;
    std::memcpy(out, &R, 64);
    return ATX_OK;
}

atx_status atx_host_transform_sphere(const float global[16], const atx_sphere* in, atx_sphere* out)
{
    // Renderer::traverseSceneGraph, per sphere (Renderer.cu:77-88)
    if (!global || !in || !out)
        return fail(ATX_ERR_INVALID, "null argument");
    atx::mat4 g;
    std::memcpy(&g, global, 64);
    const atx::vec4 c = g * atx::vec4(in->center[0], in->center[1], in->center[2], 1.0f);
    const atx::vec3 c3 = atx::vec3(c.x, c.y, c.z) / c.w;
    const float sx = atx::length(atx::vec3(g[0].x, g[0].y, g[0].z));
    const float sy = atx::length(atx::vec3(g[1].x, g[1].y, g[1].z));
    const float sz = atx::length(atx::vec3(g[2].x, g[2].y, g[2].z));
    const float uniformScale = (sx + sy + sz) / 3.0f;
    *out = *in;
    out->center[0] = c3.x; out->center[1] = c3.y; out->center[2] = c3.z;
    out->radius = in->radius * uniformScale; // "radius *= uniformScale"
    return ATX_OK;
}

atx_status atx_host_camera_update(float position[3], float direction[3], float last_mouse[2], const atx_camera_input* input,
                                  float dt, int* moved)
{
    // Camera::onUpdate (Camera.cpp:30-108), statement for statement
    if (!position || !direction || !last_mouse || !input)
        return fail(ATX_ERR_INVALID, "null argument");
    atx::vec3 pos(position[0], position[1], position[2]), dir(direction[0], direction[1], direction[2]);
    const atx::vec2 mousePos(input->mouse_x, input->mouse_y);
    const atx::vec2 mouseDelta = (mousePos - atx::vec2(last_mouse[0], last_mouse[1])) * 0.002f;
    last_mouse[0] = mousePos.x;
    last_mouse[1] = mousePos.y;
    if (moved)
        *moved = 0;
    if (!input->right_button)
        return ATX_OK; // :36-40
    bool didMove = false;
    const atx::vec3 up(0.0f, 1.0f, 0.0f);
    const atx::vec3 right = atx::cross(dir, up);
    const float speed = 5.0f;
    if (input->keys & ATX_KEY_W) { pos += dir * speed * dt; didMove = true; }
    else if (input->keys & ATX_KEY_S) { pos -= dir * speed * dt; didMove = true; }
    if (input->keys & ATX_KEY_A) { pos -= right * speed * dt; didMove = true; }
    else if (input->keys & ATX_KEY_D) { pos += right * speed * dt; didMove = true; }
    if (input->keys & ATX_KEY_Q) { pos -= up * speed * dt; didMove = true; }
    else if (input->keys & ATX_KEY_E) { pos += up * speed * dt; didMove = true; }
    if (mouseDelta.x != 0.0f || mouseDelta.y != 0.0f)
    {
        const float yaw = mouseDelta.x * 0.3f;   // getRotationSpeed(), Camera.cpp:129-132
        const float pitch = mouseDelta.y * 0.3f;
        const atx::quat orientation =
            atx::normalize(atx::cross(atx::angleAxis(-pitch, right), atx::angleAxis(-yaw, atx::vec3(0.0f, 1.0f, 0.0f))));
        dir = atx::rotate(orientation, dir);
        didMove = true;
    }
    position[0] = pos.x; position[1] = pos.y; position[2] = pos.z;
    direction[0] = dir.x; direction[1] = dir.y; direction[2] = dir.z;
    if (moved)
        *moved = didMove ? 1 : 0;
    return ATX_OK;
}

} // extern "C"
