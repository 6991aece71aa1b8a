// Host/device portability shim + the carry-chain primitives every field routine is built from.
//
// On the device each primitive is ONE PTX instruction of the add.cc / addc / mad.lo.cc / madc.hi.cc
// family (ptxas turns adjacent lo/hi pairs into IMAD.WIDE.U32 with an explicit carry predicate).
// When the same headers are compiled by a host compiler (tests/host_emul.cpp only -- the product
// library never does that) the primitives are emulated with a thread-local carry flag so the exact
// limb-level algorithms can be unit-tested on a machine without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_D inline
#endif

namespace zk {

#if defined(__CUDA_ARCH__)

// ---- device: real PTX ---------------------------------------------------------------------------
#define ZK_ASM asm volatile
ZK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#undef ZK_ASM

#else

// ---- host: emulation with an explicit carry flag (unit tests only) -------------------------------
namespace detail { inline uint32_t& cc() { static thread_local uint32_t f = 0; return f; } }
inline uint32_t emu_add(uint64_t a, uint64_t b, uint32_t cin, bool set) {
  uint64_t s = a + b + cin; if (set) detail::cc() = (uint32_t)(s >> 32); return (uint32_t)s;
}
inline uint32_t emu_sub(uint64_t a, uint64_t b, uint32_t bin, bool set) {
  uint64_t s = a - b - bin; if (set) detail::cc() = (uint32_t)((s >> 32) & 1); return (uint32_t)s;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, detail::cc(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add(a, b, detail::cc(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, detail::cc(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub(a, b, detail::cc(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, detail::cc(), true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, detail::cc(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, detail::cc(), false); }

#endif

}  // namespace zk
