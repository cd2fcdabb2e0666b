// ptx_ops.cuh -- the carry-chain integer primitives every field kernel is built from.
//
// On the device each primitive is ONE PTX instruction (add.cc / addc.cc / sub.cc /
// subc.cc / mad.lo.cc / madc.hi.cc ...) issued as its own `asm volatile` statement;
// the condition-code register carries between consecutive statements.  ptxas fuses
// an adjacent {mad.lo.cc, madc.hi.cc} pair on the same operands into a single
// IMAD.WIDE.U32(.X) with a predicate carry -- the unit the integer-pipe roofline is
// counted in.  sm_100a has no 64-bit
// multiplier, so 8 x 32-bit limbs is the native shape; the reference's 4 x u64
// `mac/adc/sbb` (src/util.rs:3-20) are what these chains replace.
//
// When the header is compiled by a plain host compiler (JJ_HOST_EMUL, used only by
// tests/emul to unit-test the kernel arithmetic source without a GPU) the same
// primitives are emulated in C with an explicit carry flag.  The product library
// never defines JJ_HOST_EMUL.
#pragma once
#include <stdint.h>

// Arithmetic formulation switches (all bit-identical results; JJ_BASELINE_ARITH restores the round-1
// first-half formulation for A/B measurements):
//   JJ_OPAQUE_ZERO   `x + carry` / `0 - x` take a never-written constant-bank zero as second source so
//                    ptxas keeps them IADD3(.X) on the ALU pipe instead of IMAD.X / IMAD.MOV
//   JJ_PRED_SUB      fe_sub adds the modulus back with 8 predicated IADD3.X (no mask AND)
//   JJ_PRED_FOLD     the squaring's q - a fold is a predicated negate (no selects)
//   JJ_REDC_M1_IMAD  Fq reduction row: k*m1 + c0 on the multiplier (1 IMAD.WIDE.X for 3 ALU instructions)
//   JJ_LAST_ROW_SUB  Fq: the last reduction row subtracts E[0]*q, the result is in (-q, q) and is fixed by a
//                    sign-predicated add (fe.cuh, redc_row_fq_last)
#if !defined(JJ_BASELINE_ARITH)
#define JJ_OPAQUE_ZERO 1
#define JJ_PRED_SUB 1
#define JJ_PRED_FOLD 1
#if !defined(JJ_NO_REDC_M1_IMAD)
#define JJ_REDC_M1_IMAD 1
#endif
#if !defined(JJ_NO_LAST_ROW_SUB)
#define JJ_LAST_ROW_SUB 1
#endif
#endif

#if defined(JJ_HOST_EMUL)
#define JJ_DEVICE static inline
#define JJ_DEVICE_SPEC inline
#define JJ_CONST_FN static constexpr
#define JJ_HD static inline
namespace jj {
static thread_local uint32_t g_cf = 0;
JJ_DEVICE void add_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; d = (uint32_t)t; g_cf = (uint32_t)(t >> 32); }
JJ_DEVICE void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cf; d = (uint32_t)t; g_cf = (uint32_t)(t >> 32); }
JJ_DEVICE void addc(uint32_t& d, uint32_t a, uint32_t b) { d = a + b + g_cf; }
JJ_DEVICE void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; d = (uint32_t)t; g_cf = (uint32_t)(t >> 63); }
JJ_DEVICE void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cf; d = (uint32_t)t; g_cf = (uint32_t)(t >> 63); }
JJ_DEVICE void subc(uint32_t& d, uint32_t a, uint32_t b) { d = a - b - g_cf; }
JJ_DEVICE void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c; d = (uint32_t)t; g_cf = (uint32_t)(t >> 32); }
JJ_DEVICE void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + g_cf; d = (uint32_t)t; g_cf = (uint32_t)(t >> 32); }
JJ_DEVICE void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + g_cf; d = (uint32_t)t; g_cf = (uint32_t)(t >> 32); }
JJ_DEVICE void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { d = (uint32_t)(((uint64_t)a * b) >> 32) + c + g_cf; }
JJ_DEVICE uint32_t umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
JJ_DEVICE uint32_t oz() { return 0u; }
}  // namespace jj
// immediate-operand forms: on the host they are ordinary values
#define JJ_MAD_LO_CC_I(d, a, IMM, c) jj::mad_lo_cc(d, a, (uint32_t)(IMM), c)
#define JJ_MADC_LO_CC_I(d, a, IMM, c) jj::madc_lo_cc(d, a, (uint32_t)(IMM), c)
#define JJ_MADC_HI_CC_I(d, a, IMM, c) jj::madc_hi_cc(d, a, (uint32_t)(IMM), c)
#define JJ_MADC_HI_I(d, a, IMM, c) jj::madc_hi(d, a, (uint32_t)(IMM), c)
#define JJ_SUB_CC_I(d, a, IMM) jj::sub_cc(d, a, (uint32_t)(IMM))
#define JJ_SUBC_CC_I(d, a, IMM) jj::subc_cc(d, a, (uint32_t)(IMM))
#define JJ_ADD_CC_I(d, a, IMM) jj::add_cc(d, a, (uint32_t)(IMM))
#define JJ_ADDC_CC_I(d, a, IMM) jj::addc_cc(d, a, (uint32_t)(IMM))
#define JJ_ADDC_I(d, a, IMM) jj::addc(d, a, (uint32_t)(IMM))

#else  // ---------------------------------------------------------------- device
#define JJ_DEVICE __device__ __forceinline__
#define JJ_DEVICE_SPEC __device__ __forceinline__
#define JJ_CONST_FN __host__ __device__ static constexpr
#define JJ_HD __host__ __device__ __forceinline__
namespace jj {
JJ_DEVICE void add_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
JJ_DEVICE void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
JJ_DEVICE void addc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
JJ_DEVICE void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
JJ_DEVICE void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
JJ_DEVICE void subc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
JJ_DEVICE void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
JJ_DEVICE void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
JJ_DEVICE void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
JJ_DEVICE void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
JJ_DEVICE uint32_t umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// "Opaque zero": a constant-bank word that is never written, so it reads as 0 but ptxas cannot
// fold it.  `x + carry` written as addc(x, x, 0) is free to become IMAD.X (and 0 - x IMAD.MOV) on the
// FMA-heavy pipe -- the pipe IMAD.WIDE saturates; with oz() as the addend the instruction has two
// register/constant sources and stays an IADD3(.X) on the ALU pipe, without a false dependency.
#if defined(JJ_OPAQUE_ZERO)
__constant__ uint32_t g_opaque_zero;
JJ_DEVICE uint32_t oz() { return g_opaque_zero; }
#else
JJ_DEVICE uint32_t oz() { return 0u; }
#endif
}  // namespace jj
#define JJ_MAD_LO_CC_I(d, a, IMM, c) asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "n"(IMM), "r"(c))
#define JJ_MADC_LO_CC_I(d, a, IMM, c) asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "n"(IMM), "r"(c))
#define JJ_MADC_HI_CC_I(d, a, IMM, c) asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "n"(IMM), "r"(c))
#define JJ_MADC_HI_I(d, a, IMM, c) asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "n"(IMM), "r"(c))
#define JJ_SUB_CC_I(d, a, IMM) asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "n"(IMM))
#define JJ_SUBC_CC_I(d, a, IMM) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "n"(IMM))
#define JJ_ADD_CC_I(d, a, IMM) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "n"(IMM))
#define JJ_ADDC_CC_I(d, a, IMM) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "n"(IMM))
#define JJ_ADDC_I(d, a, IMM) asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "n"(IMM))
#endif
