// atx_exact.cuh — the arithmetic vocabulary of the path-tracing kernels.
//
// Parity contract: primary hit indices and sample counts must be
// bit-exact against the reference's CUDA renderer, which is compiled with
// -use_fast_math (Engine/CMakeLists.txt:8). That flag fixes, per source
// expression, one PTX instruction: add/sub/mul/fma.rn.ftz, div.approx.ftz,
// sqrt.approx.ftz, rsqrt.approx.ftz, lg2/ex2/sin/cos.approx.ftz — and which
// mul+add pairs nvcc contracts into fma. The kernels therefore never write
// `a*b+c`; every floating-point operation is one of the functions below, each
// exactly one PTX instruction that the compiler can neither contract, split nor
// reassociate. Ground truth for the op sequences is the reference's sm_100a SASS
// (cuobjdump of oracle/_ref/ref_headless), not its PTX: the PTX carries mul/add/sub
// without ".rn", which ptxas may — and in five places does — contract further
// (listed function by function in DESIGN.md §4).
//
// Two families:
//   f*    : the device's fast-math family (round-to-nearest, flush-to-zero,
//           approximate MUFU ops) — the path loop (Renderer.cu:251-409, BRDF.cu).
//   ieee_*: IEEE-754 round-to-nearest WITHOUT flush and WITHOUT contraction —
//           primary-ray generation, which the reference runs on the host in plain
//           g++ float arithmetic (Camera.cpp:161-195).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace atxk
{
#define ATX_DEV __device__ __forceinline__

// ---- fast-math family (.ftz) ------------------------------------------------
ATX_DEV float fadd(float a, float b) { float r; asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float fsub(float a, float b) { float r; asm("sub.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float fmul(float a, float b) { float r; asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float ffma(float a, float b, float c) { float r; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
ATX_DEV float fneg(float a) { float r; asm("neg.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float fabs_(float a) { float r; asm("abs.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float fmin_(float a, float b) { float r; asm("min.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float fmax_(float a, float b) { float r; asm("max.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float fdiv_approx(float a, float b) { float r; asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float fsqrt_approx(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float frsqrt_approx(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float flg2_approx(float a) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float fex2_approx(float a) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float fsin_approx(float a) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float fcos_approx(float a) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ATX_DEV float u32_to_f32_rn(uint32_t a) { float r; asm("cvt.rn.f32.u32 %0, %1;" : "=f"(r) : "r"(a)); return r; }
ATX_DEV uint32_t f32_to_u32_rz_ftz(float a) { uint32_t r; asm("cvt.rzi.ftz.u32.f32 %0, %1;" : "=r"(r) : "f"(a)); return r; }

// comparisons under .ftz (setp.*.ftz.f32): a denormal operand compares as zero.
ATX_DEV bool flt(float a, float b) { int r; asm("{ .reg .pred p; setp.lt.ftz.f32 p, %1, %2; selp.s32 %0, 1, 0, p; }" : "=r"(r) : "f"(a), "f"(b)); return r != 0; }
ATX_DEV bool fgt(float a, float b) { int r; asm("{ .reg .pred p; setp.gt.ftz.f32 p, %1, %2; selp.s32 %0, 1, 0, p; }" : "=r"(r) : "f"(a), "f"(b)); return r != 0; }

// ---- packed pairs (sm_100 f32x2: FADD2 / FMUL2 / FFMA2) --------------------------
// Two independent fp32 lanes in one 64-bit register, each rounded exactly like the
// scalar .rn.ftz op, so a packed op is bit-identical to two scalar ops. One issue slot
// per two flops-pairs: the FMA pipe keeps its 128 lane-ops/clk/SM while the issue port
// has room for the loads, shifts and branches of the sphere loop (measured with
// tools/ubench_fp32.cu: ffma 122, ffma2 127 lane-ops/clk/SM at half the issue slots).
// ptxas turns pk2(x, x) into a scalar-broadcast operand (Rn.F32), so uniform operands
// cost no extra registers.
typedef unsigned long long f32x2;
ATX_DEV f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
ATX_DEV float lo2(f32x2 v) { return __uint_as_float(static_cast<uint32_t>(v)); }
ATX_DEV float hi2(f32x2 v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
ATX_DEV f32x2 fadd2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ATX_DEV f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ATX_DEV f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// mask = (mask << 1) | sign(v): one SHF (ALU pipe) per test, no predicate, no branch
ATX_DEV uint32_t shift_in_sign(uint32_t mask, float v)
{
    uint32_t r;
    asm("shf.l.clamp.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(__float_as_uint(v)), "r"(mask));
    return r;
}

// dot(a,b) as nvcc contracts glm's (x*x' + y*y') + z*z' for the reference:
// mul(y,y') -> fma(x,x',.) -> fma(z,z',.)   (Renderer PTX, every dot product)
ATX_DEV float fdot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return ffma(az, bz, ffma(ax, bx, fmul(ay, by)));
}

// ---- IEEE family (no ftz, no contraction): host-equivalent arithmetic --------
ATX_DEV float ieee_add(float a, float b) { float r; asm("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float ieee_sub(float a, float b) { float r; asm("sub.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float ieee_mul(float a, float b) { float r; asm("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float ieee_div(float a, float b) { float r; asm("div.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
ATX_DEV float ieee_sqrt(float a) { float r; asm("sqrt.rn.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// ---- RNG: Random::PcgHash / PcgFloat, Core/include/Random.h:59-70 ------------
ATX_DEV uint32_t pcg_hash(uint32_t seed)
{
    const uint32_t state = seed * 747796405u + 2891336453u;
    const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

// seed = PcgHash(seed); float(seed) / float(UINT32_MAX)  ->  cvt.rn.f32.u32 ; div.approx.ftz by 2^32
ATX_DEV float pcg_float(uint32_t& seed)
{
    seed = pcg_hash(seed);
    return fdiv_approx(u32_to_f32_rn(seed), 4294967296.0f);
}

} // namespace atxk
