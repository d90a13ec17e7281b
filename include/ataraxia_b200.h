/*
 * ataraxia_b200.h — C-ABI of the B200-native path-tracing hot path.
 *
 * This is the drop-in boundary. The reference (1neskk/Ataraxia)
 * has no FFI layer: its boundary is the C++ class API in Engine/include. The
 * C++ mirror of that API (include/ataraxia/Ataraxia.h: Renderer, Scene, SceneNode,
 * Camera, Material, Light, Settings, Utils) is a thin host layer over the
 * functions declared here, and every function cites the reference code it
 * replaces. Plain pointers and sizes only; no torch / glm / STL types.
 *
 * Conventions
 *   - every function returns an atx_status (0 = ok, negative = error); the
 *     message of the last error on the calling thread is atx_last_error().
 *     Nothing here ever calls exit() (the reference's CUDA_CHECK does,
 *     Engine/include/DeviceMemory.h:7-16).
 *   - one handle = one device + one non-default CUDA stream; handles are
 *     independent; a handle is not thread-safe.
 *   - host arrays passed in are copied during the call; output pointers are
 *     caller-allocated HOST memory unless the name says "_device".
 *   - there is NO CPU fallback: without a CUDA device atx_create fails.
 */
#ifndef ATARAXIA_B200_H
#define ATARAXIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ATX_API __declspec(dllexport)
#else
#define ATX_API __attribute__((visibility("default")))
#endif

typedef int atx_status;
enum {
    ATX_OK = 0,
    ATX_ERR_INVALID = -1,   /* bad argument / bad state */
    ATX_ERR_CUDA = -2,      /* CUDA runtime error (message has the string) */
    ATX_ERR_NCCL = -3,      /* NCCL error or NCCL library not loadable */
    ATX_ERR_NO_DEVICE = -4, /* no CUDA device: there is no CPU fallback */
    ATX_ERR_ALLOC = -5
};

/* POD scene records. Byte-identical to the reference's structs so a caller can
 * pass std::vector<Sphere>::data() etc. unchanged. */
typedef struct atx_sphere {   /* == Sphere, Engine/include/SceneNode.h:11-21 (20 B) */
    float center[3];
    float radius;
    int32_t material;         /* Sphere::id = material index */
} atx_sphere;

typedef struct atx_material { /* == Material, Engine/include/Scene.h:28-47 (52 B) */
    float albedo[3];
    float roughness;
    float metallic;
    float F0[3];
    float emissionColor[3];
    float emissionIntensity;
    int32_t id;
} atx_material;

typedef struct atx_light {    /* == Light, Engine/include/Scene.h:17-26 (28 B) */
    float position[3];
    float color[3];
    float intensity;
} atx_light;

typedef struct atx_counters {
    uint64_t paths;           /* perPixel evaluations (pixel x sample) */
    uint64_t rays;            /* traceRay calls of the reference (closest-hit + shadow) */
    uint64_t sphere_tests;    /* rays x numSpheres (brute force, Renderer.cu:256) */
    uint64_t launches;        /* kernels of this library launched on the handle */
    uint64_t rays_traced;     /* rays whose sphere loop ran on the device: the primary ray of a
                                 pixel is the same every frame (no jitter, Camera.cpp:176-187), so
                                 it is traced once per launch and its hit reused by every frame */
    uint64_t sphere_tests_executed; /* rays_traced x numSpheres: what the roofline is computed from */
} atx_counters;

/* kernel family for atx_render* */
enum {
    ATX_VARIANT_AUTO = 0,      /* the faster of the two as measured by atx_calibrate */
    ATX_VARIANT_MEGAKERNEL = 1,/* persistent path loops in registers, pixel pool, path regeneration */
    ATX_VARIANT_WAVEFRONT = 2  /* path records in HBM, one launch per bounce, ray compaction */
};

typedef struct atx_renderer* atx_handle;

/* ---- lifetime ------------------------------------------------------------ */

/* Create a renderer on CUDA device `device_ordinal` (reference: Renderer::Renderer,
 * Renderer.cu:13-15, which implies device 0). Fails with ATX_ERR_NO_DEVICE when
 * no CUDA device is present. */
ATX_API atx_status atx_create(int device_ordinal, atx_handle* out);

/* Renderer::~Renderer, Renderer.cu:17-23. */
ATX_API atx_status atx_destroy(atx_handle h);

/* Message of the last failing call on this thread ("" if none). */
ATX_API const char* atx_last_error(void);

/* Library/ABI version and the SM architecture the kernels were built for. */
ATX_API const char* atx_version(void);

/* ---- state --------------------------------------------------------------- */

/* Renderer::onResize, Renderer.cu:98-146: (re)allocates the RGBA8 image and the
 * float4 accumulation buffer and resets frameIndex to 1. No-op when the size is
 * unchanged. Zero width/height is ATX_ERR_INVALID (reference prints and returns,
 * Camera.cpp:112-116). */
ATX_API atx_status atx_resize(atx_handle h, uint32_t width, uint32_t height);

/* Renderer::allocateDeviceMemory, Renderer.cu:25-57, for an already flattened
 * scene (world-space spheres from Renderer::traverseSceneGraph, :67-96).
 * Sphere material indices >= n_materials (as unsigned) are clamped to 0 exactly
 * as Renderer.cu:30-37 does. The library packs SoA records for shared memory. */
ATX_API atx_status atx_upload_scene(atx_handle h,
                                    const atx_sphere* spheres, size_t n_spheres,
                                    const atx_material* materials, size_t n_materials,
                                    const atx_light* lights, size_t n_lights);

/* Camera(fov, near, far, position, direction) + Camera::Resize(w,h) of the
 * current size: builds projection/view and their inverses on the host with the
 * reference's evaluation order (Camera.cpp:14-29, :134-159) — primary rays are
 * then generated in-kernel, bit-identical to Camera::UpdateRayDirection
 * (Camera.cpp:161-195). Replaces Camera::allocateDevice/freeDevice (:197-220). */
ATX_API atx_status atx_set_camera(atx_handle h, const float position[3], const float direction[3],
                                  float fov_degrees, float near_clip, float far_clip);

/* Same, from explicit column-major inverse-projection and inverse-view matrices
 * (Camera::getInverseProjectionMatrix / getInverseViewMatrix, Camera.h:49-50). */
ATX_API atx_status atx_set_camera_matrices(atx_handle h, const float position[3],
                                           const float inv_projection[16], const float inv_view[16]);

/* Renderer::setSettings, Renderer.h:28; Settings = Scene.h:49-56. */
ATX_API atx_status atx_set_settings(atx_handle h, int accumulation, int sky_light, int max_bounces);

/* Tuning knobs that never change results (parity tests sweep them). */
enum {
    ATX_TUNE_CHUNK_SPHERES = 1, /* spheres per shared-memory chunk; 0 = automatic. Forces the
                                   chunked (double-buffered) staging path when < n_spheres. */
    ATX_TUNE_MEGA_KIND = 2,     /* megakernel form: 0 = by sphere count, 1 = while-while (one pixel
                                   per thread, hits gathered before the shading phase), 2 = two-slot
                                   packed (two pixels per thread, f32x2 sphere loop), 3 = warp-queue
                                   (one 8x4 tile per warp, hits queued in shared memory and bounced 32
                                   at a time; the automatic choice for <= 16 spheres, at most one light
                                   and >= 4 frames per launch), 4 = two-slot packed in lockstep (every
                                   slot of a CTA alternates closest-hit and shadow traces together, so the
                                   shading between traces runs warp-wide; the automatic choice above 16
                                   spheres from 4 frames per launch) */
    ATX_TUNE_PARK_THRESHOLD = 3 /* while-while form: parked hits per warp (1..32) that trigger the
                                   bounce phase (default 8) */,
    ATX_TUNE_CLAIM_THRESHOLD = 4,/* idle lanes per warp (1..32) that trigger a batched claim from the
                                   pixel pool; 0 (default) = per form: 32 while-while (whole 8x4
                                   tiles), 3 warp-queue, 2 two-slot packed */
    ATX_TUNE_REDUCE = 5          /* atx_allreduce_accum: 0 (default) = one kernel over NVLink peer memory
                                   when every rank can map every other rank's buffer, 1 = ncclAllReduce.
                                   Must be set alike on all ranks. */
};
ATX_API atx_status atx_set_tuning(atx_handle h, int key, int64_t value);

/* Renderer::resetFrameIndex, Renderer.h:30 (the next render zeroes the
 * accumulation buffer, Renderer.cu:181-182). */
ATX_API atx_status atx_reset(atx_handle h);

/* Current frameIndex (1 after a reset; Renderer.cu:245-248). */
ATX_API atx_status atx_frame_index(atx_handle h, uint32_t* out);

/* ---- the hot path -------------------------------------------------------- */

/* n_frames x { Renderer::Render, Renderer.cu:173-249 } in ONE launch: frames
 * frameIndex .. frameIndex+n_frames-1 are accumulated in registers in the
 * reference's order (so the float4 sums are bit-identical to n_frames
 * sequential reference frames), one 16 B read + one 16 B write per pixel.
 * frameIndex advances by n_frames (or stays 1 when accumulation is off, :245-248).
 * Asynchronous on the handle's stream; atx_sync or any atx_read_* waits. */
ATX_API atx_status atx_render(atx_handle h, uint32_t n_frames, int variant);

/* Multi-GPU spp split: accumulate frame indices
 * first, first+stride, ... (n_frames of them) WITHOUT touching frameIndex.
 * zero_first != 0 starts from a zeroed accumulation buffer. */
ATX_API atx_status atx_render_frames(atx_handle h, uint32_t first_frame, uint32_t n_frames,
                                     uint32_t frame_stride, int zero_first, int variant);

/* Decide what ATX_VARIANT_AUTO means for the current scene, camera, size and settings by
 * measurement: renders n_frames frames with each variant into a scratch buffer (the
 * accumulation buffer and frameIndex are untouched), times them with CUDA events and keeps
 * the faster. The variants are bit-identical, so the choice never changes a result. The
 * lane divergence of the megakernel against the HBM traffic of the wavefront queues is
 * exactly what the two times weigh. Either output may be NULL; wavefront_ms is -1 when
 * that variant cannot run the configuration. Without a calibration AUTO = megakernel. */
ATX_API atx_status atx_calibrate(atx_handle h, uint32_t n_frames, float* megakernel_ms, float* wavefront_ms);

/* Wait for everything queued on the handle's stream. */
ATX_API atx_status atx_sync(atx_handle h);

/* Device time (ms, CUDA events on the handle's stream) of the last atx_render*
 * call's kernels. Synchronises. */
ATX_API atx_status atx_last_render_ms(atx_handle h, float* out_ms);

/* Megakernel form the last launch used (1 while-while, 2 two-slot packed, 3 warp-queue, 4 two-slot packed in lockstep; 0 before any launch
 * or after a wavefront launch): what ATX_TUNE_MEGA_KIND = 0 resolved to. */
ATX_API atx_status atx_last_mega_kind(atx_handle h, int* out);

/* General-purpose device timers on the handle's stream (CUDA events), e.g. to bracket
 * render + all-reduce. slot in [0, 8). elapsed synchronises on the later event. */
ATX_API atx_status atx_event_record(atx_handle h, int slot);
ATX_API atx_status atx_event_elapsed_ms(atx_handle h, int slot_begin, int slot_end, float* out_ms);

/* ---- results ------------------------------------------------------------- */

/* float4 accumulation buffer (Renderer::d_accumulation_, Renderer.h:55):
 * width*height*4 floats, row-major x + y*width, .w = exact sample count. */
ATX_API atx_status atx_read_accum(atx_handle h, float* dst);

/* Overwrite the accumulation buffer from host data and set frameIndex — the
 * resumable-render state. */
ATX_API atx_status atx_write_accum(atx_handle h, const float* src, uint32_t next_frame_index);

/* ---- resumable renders on disk ---------------------------- */

/* The complete state of a progressive render is the accumulation buffer and the next frame index
 * (Renderer.cu:165-168, :181-182, :245-248: the RNG is a pure function of pixel and frameIndex).
 * atx_save_checkpoint writes both to `path` (160-byte header + width*height float4), bound to the scene and
 * camera by a SHA-256 over the uploaded records and the camera matrices, with a SHA-256 of the payload.
 * next_frame_index = 0 means "the handle's frameIndex" (what atx_render continues with); callers that drive
 * atx_render_frames themselves (one rank's share of an spp-split render) pass their own next frame index and
 * frame_stride (0 = 1). Written to `path`.part and renamed. */
ATX_API atx_status atx_save_checkpoint(atx_handle h, const char* path, uint32_t next_frame_index, uint32_t frame_stride);

/* Load a checkpoint into the accumulation buffer and set frameIndex. The renderer must already have the
 * image size, scene, camera and settings the file was rendered with: a different size, maxBounces, skyLight,
 * or scene/camera hash is refused with ATX_ERR_INVALID (continuing would mix two images), as is a truncated or
 * corrupt file. Continuing after a load is bit-identical to never having stopped. */
ATX_API atx_status atx_load_checkpoint(atx_handle h, const char* path, uint32_t* next_frame_index, uint32_t* frame_stride);

/* SHA-256 of what a checkpoint binds to: uploaded scene records, camera position and the two inverse matrices. */
ATX_API atx_status atx_scene_sha256(atx_handle h, uint8_t out[32]);

/* SHA-256 of a host buffer (the digest used above; exposed so callers can verify read-backs). */
ATX_API atx_status atx_host_sha256(const void* data, size_t bytes, uint8_t out[32]);

/* RGBA8 image as Renderer::Render leaves it in h_imageData_ (Renderer.cu:165-168,
 * :240; packing colorUtils::vec4ToRGBA, Renderer.h:70-78): clamp(acc/divisor,0,1),
 * truncating, alpha from acc.w. divisor = 0 means "frameIndex of the last
 * rendered frame". */
ATX_API atx_status atx_read_rgba8(atx_handle h, uint32_t* dst, uint32_t divisor);

/* Parity/debug: index of the closest sphere hit by each primary ray (-1 = miss),
 * via the same intersection code as the path loop (Renderer::traceRay,
 * Renderer.cu:251-285). width*height int32. */
ATX_API atx_status atx_read_hit_ids(atx_handle h, int32_t* dst);

/* Parity/debug: the in-kernel primary ray directions, width*height*3 floats
 * (the table Camera::getRayDirection returns, Camera.h:60). */
ATX_API atx_status atx_read_ray_directions(atx_handle h, float* dst);

/* Exact device counters since the last atx_reset_counters. */
ATX_API atx_status atx_get_counters(atx_handle h, atx_counters* out);
ATX_API atx_status atx_reset_counters(atx_handle h);

/* Raw device pointers for interop (torch tensors, NCCL, peer access). */
ATX_API atx_status atx_accum_device_ptr(atx_handle h, void** out);
ATX_API atx_status atx_stream(atx_handle h, void** out_cuda_stream);

/* Page-locked host memory for read-back destinations (atx_read_rgba8, atx_read_accum): the copy becomes one DMA
 * instead of a staged one. Optional: every read function takes any host pointer. Needs a CUDA device. */
ATX_API atx_status atx_host_alloc(size_t bytes, void** out);
ATX_API atx_status atx_host_free(void* p);

/* ---- multi-GPU: spp split + sum of accumulation buffers over NVLink ------- */

/* 128-byte NCCL unique id (rank 0 creates, the host broadcasts it). */
ATX_API atx_status atx_comm_unique_id(uint8_t id[128]);
ATX_API atx_status atx_comm_init_rank(atx_handle h, int n_ranks, int rank, const uint8_t id[128]);
ATX_API atx_status atx_comm_destroy(atx_handle h);

/* Sum of the accumulation buffers of all ranks, in place, on the handle's stream (4*width*height floats; .w sums to
 * the exact total spp). COLLECTIVE. Default transport: ONE kernel over NVLink peer memory (atx_p2p.cu): every rank's
 * buffer is mapped into every process (CUDA IPC, set up collectively at the first call after atx_comm_init_rank or
 * atx_resize), rank r sums the r-th slice of all buffers in rank order and stores it to all ranks; two flag barriers
 * in peer memory bracket it. Every rank ends with the same bits. Falls back to ncclAllReduce(float32, sum) when the
 * buffers cannot be mapped (ranks sharing a process, no peer access) or ATX_TUNE_REDUCE = 1. A rank that never joins
 * makes the others fail with ATX_ERR_NCCL after ATX_P2P_TIMEOUT_MS (environment, default 60000) at their next
 * atx_sync / atx_read_accum instead of hanging. */
ATX_API atx_status atx_allreduce_accum(atx_handle h);

/* Image-tile split (the alternative to the spp split). COLLECTIVE. Rank r of R renders ALL of the frames
 * first_frame .. first_frame + n_frames - 1 for the 8x4-pixel tiles r, r + R, r + 2R, ... (row-major tile order: a
 * fine interleave, so sky and geometry spread evenly), and its kernel stores every finished pixel into the image of
 * EVERY rank over NVLink peer memory while it renders; flag barriers in peer memory open and close the step. No
 * arithmetic on the wire and each pixel is summed on one GPU in frame order: every rank ends with an image that is
 * bit-identical to a single-GPU atx_render_frames(first_frame, n_frames, 1, zero_first). With zero_first = 0 the
 * frames are added to the image every rank already holds (all ranks must hold the same one: the result of an earlier
 * atx_render_tiles or atx_allreduce_accum). Without peer mappings (see atx_allreduce_accum) the ranks' tiles are
 * combined with ncclAllReduce instead (same bits, more traffic). Without a communicator it renders the whole image. */
ATX_API atx_status atx_render_tiles(atx_handle h, uint32_t first_frame, uint32_t n_frames, int zero_first, int variant);

/* One share of an image-tile split rendered locally, no communication: the 8x4 tiles share, share + n_shares, ... of
 * the image get the frames first_frame .. first_frame + n_frames - 1 (zero_first: starting from zero instead of the
 * stored sums); every other pixel of the accumulation buffer is left untouched. The shares 0 .. n_shares-1 together
 * are bit-identical to one atx_render_frames of the whole image. */
ATX_API atx_status atx_render_tile_share(atx_handle h, uint32_t first_frame, uint32_t n_frames, int zero_first, int variant,
                                         uint32_t n_shares, uint32_t share);

/* Transport the last atx_allreduce_accum / atx_render_tiles used: 0 none yet, 1 peer memory, 2 ncclAllReduce. */
ATX_API atx_status atx_last_reduce_kind(atx_handle h, int* out);

/* Progressive preview across ranks: the sum of every rank's accumulation buffer into a separate preview buffer
 * (one out-of-place ncclAllReduce on the handle's stream; a device copy without a communicator). The ranks' own
 * buffers are not touched, so rendering continues afterwards — what a live multi-GPU view of a long render
 * needs (the reference shows its single buffer every frame, Renderer.cu:165-168, main.cpp:185-187). */
ATX_API atx_status atx_allreduce_preview(atx_handle h);
/* Read the preview: the float4 sums (accum_dst, may be NULL) and/or the RGBA8 image resolved with
 * `divisor` = total samples per pixel over all ranks (rgba_dst, may be NULL). Overwrites the display buffer. */
ATX_API atx_status atx_read_preview(atx_handle h, float* accum_dst, uint32_t* rgba_dst, uint32_t divisor);

/* ---- host math helpers ------------------------------------------------------
 * CPU-side pieces of the reference's HOST API (they run on the CPU in the reference
 * too), compiled inside this library so the caller's compiler flags cannot perturb
 * the exact glm evaluation order parity depends on. Used by the header-only C++
 * mirror (include/ataraxia/Ataraxia.h). None of them is on the render path. Matrices are
 * column-major float[16] (glm::mat4 layout). */

/* Camera::UpdateProjectionMatrix + UpdateViewMatrix, Camera.cpp:134-159. Any output
 * pointer may be NULL. */
ATX_API atx_status atx_host_camera_matrices(const float position[3], const float direction[3], float fov_degrees,
                                            float near_clip, float far_clip, uint32_t width, uint32_t height,
                                            float projection[16], float view[16],
                                            float inv_projection[16], float inv_view[16]);

/* Camera::UpdateRayDirection, Camera.cpp:161-195 (multithreaded like the reference):
 * the table Camera::getRayDirection() exposes. out = width*height*3 floats. */
ATX_API atx_status atx_host_ray_directions(const float inv_projection[16], const float inv_view[16],
                                           uint32_t width, uint32_t height, float* out);

/* SceneNode::updateGlobalTransform, SceneNode.cpp:42-59: local = T * R * S,
 * global = parent * local. rotation is (x, y, z, w). */
ATX_API atx_status atx_host_node_transform(const float parent[16], const float position[3],
                                           const float rotation_xyzw[4], const float scale[3],
                                           float local[16], float global[16]);

/* out = a * b in glm's column-by-column order (type_mat4x4.inl mul4x4), e.g. the
 * non-dirty branch "m_globalTransform = parentTransform * m_localTransform",
 * SceneNode.cpp:51-54. */
ATX_API atx_status atx_host_mat4_mul(const float a[16], const float b[16], float out[16]);

/* One sphere of Renderer::traverseSceneGraph, Renderer.cu:77-88. */
ATX_API atx_status atx_host_transform_sphere(const float global[16], const atx_sphere* in, atx_sphere* out);

/* Camera::onUpdate, Camera.cpp:30-108, with the GLFW queries (Core/src/input/Input.cpp:9-35) replaced
 * by an explicit input record, so that camera motion can be scripted without a window. keys is a
 * bit set of the six keys the camera reads; mouse_x/mouse_y is the cursor position
 * (Input::GetMousePosition), right_button whether MouseButton::Right is held. Updates position,
 * direction and last_mouse in place exactly as the reference does (same glm evaluation order: the
 * delta is taken and last_mouse advanced even when the button is up, then W/S, A/D, Q/E at speed 5,
 * then the yaw/pitch quaternion at rotation speed 0.3) and reports what onUpdate returns in *moved.
 * The caller rebuilds the view matrix when *moved is set (Camera.cpp:101-105). */
typedef struct atx_camera_input {
    uint32_t keys;          /* ATX_KEY_* bits */
    uint32_t right_button;  /* non-zero: held */
    float mouse_x, mouse_y;
} atx_camera_input;
enum { ATX_KEY_W = 1, ATX_KEY_S = 2, ATX_KEY_A = 4, ATX_KEY_D = 8, ATX_KEY_Q = 16, ATX_KEY_E = 32 };
ATX_API atx_status atx_host_camera_update(float position[3], float direction[3], float last_mouse[2],
                                          const atx_camera_input* input, float dt, int* moved);

#ifdef __cplusplus
}
#endif
#endif /* ATARAXIA_B200_H */
