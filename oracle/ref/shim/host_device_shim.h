// TEST INFRASTRUCTURE (oracle build only). Force-included (-include) ahead of the
// UNMODIFIED reference TUs Renderer.cu / BRDF.cu so that every function they mark
// __device__ is ALSO emitted for the host: that is the north_star's "per-pixel
// shading compiled host-side" CPU baseline, built from the reference's own
// source text. The toolkit/glm/curand headers are pulled in first (with the
// stock meaning of __device__) so only the reference's own declarations change.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <device_launch_parameters.h>
#include <random>
#include <algorithm>
#include <iostream>
#include <cfloat>
#include <cmath>
#include <vector>
#include <memory>
#include <string>
#define GLM_FORCE_CUDA
#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/ext/scalar_constants.hpp>
#undef __device__
#define __device__ __location__(host) __location__(device)
// Random.h:14-52 wraps cuRAND device functions that the reference never calls
// (dead code). They cannot exist on the host, so in THIS build
// only they are redirected to inert stubs; PcgHash/PcgFloat (the RNG actually
// used, Random.h:59-70) are untouched.
__host__ __location__(device) inline void atx_dead_curand_init(unsigned long long, int, int, curandState*) {}
__host__ __location__(device) inline unsigned int atx_dead_curand(curandState*) { return 0u; }
__host__ __location__(device) inline float atx_dead_curand_uniform(curandState*) { return 0.0f; }
#define curand_init atx_dead_curand_init
#define curand atx_dead_curand
#define curand_uniform atx_dead_curand_uniform
