// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (modconv_tc.cu, modconv_tc3.cu): mbarrier, TMA tensor loads,
// UMMA shared-memory / instruction descriptors, tcgen05.mma / commit / ld.  Everything is internal (anonymous namespace).
#pragma once
#include <cuda.h>
#include <cstdint>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 24)) __trap();
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64): SWIZZLE_128B = 2 (K-major operand),
// SWIZZLE_128B_BASE32B = 1 (32-bit MN-major operand: 32-byte swizzle granules, 4-row atoms).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;

__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes, uint32_t layout = kLayoutSw128)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lboBytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sboBytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- warp-uniform issue -------------------------------------------------------------------------------------------
// Measured (tools/tc_mma_bench.cu): when tcgen05.mma sits in a single-thread branch (`if (lane == 0)`), ptxas wraps every
// UTCHMMA in an ELECT/branch waterfall to move its operands into uniform registers and one MMA costs ~140 clk of issue
// time -- more than the 64..128 clk the tensor core needs for N = 128..256.  The helpers below are called by ALL lanes of
// the (converged) MMA warp; elect.sync picks the issuing lane, the operands stay in uniform registers and four MMAs go
// out back to back (N = 64: 48 clk, N = 96: 56 clk, N >= 128: N/2 clk per MMA, the tensor-core floor).
//
// Descriptors are passed as 32-bit halves: hi = SBO | version | layout (constant per operand kind), lo = LBO | start >> 4;
// a K step only adds to the start field (AS / BS in 16-byte units: 2 for a K-major SWIZZLE_128B operand, 64 for the
// 32-bit MN-major one).
template <int AS, int BS>
__device__ __forceinline__ void umma_tf32_x4(uint32_t d, uint32_t aLo, uint32_t aHi, uint32_t bLo, uint32_t bHi, uint32_t idesc,
                                             uint32_t accumulateFirst)
{
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, %5, %5;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\tadd.u32 bl, %3, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, t;\n\t"
        "add.u32 al, %1, %9;\n\tadd.u32 bl, %3, %10;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, t;\n\t"
        "add.u32 al, %1, %11;\n\tadd.u32 bl, %3, %12;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, t;\n\t}"
        ::"r"(d), "r"(aLo), "r"(aHi), "r"(bLo), "r"(bHi), "r"(idesc), "r"(accumulateFirst),
          "n"(AS), "n"(BS), "n"(2 * AS), "n"(2 * BS), "n"(3 * AS), "n"(3 * BS) : "memory");
}

// Same for kind::f16 (fp16 operands, fp32 accumulate; K = 16 per instruction).
template <int AS, int BS>
__device__ __forceinline__ void umma_f16_x4(uint32_t d, uint32_t aLo, uint32_t aHi, uint32_t bLo, uint32_t bHi, uint32_t idesc,
                                            uint32_t accumulateFirst)
{
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, %5, %5;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "add.u32 al, %1, %7;\n\tadd.u32 bl, %3, %8;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, t;\n\t"
        "add.u32 al, %1, %9;\n\tadd.u32 bl, %3, %10;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, t;\n\t"
        "add.u32 al, %1, %11;\n\tadd.u32 bl, %3, %12;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, t;\n\t}"
        ::"r"(d), "r"(aLo), "r"(aHi), "r"(bLo), "r"(bHi), "r"(idesc), "r"(accumulateFirst),
          "n"(AS), "n"(BS), "n"(2 * AS), "n"(2 * BS), "n"(3 * AS), "n"(3 * BS) : "memory");
}

// tcgen05.commit from the elected lane of a converged warp.
__device__ __forceinline__ void umma_commit_elect(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}

// TMA issue from the elected lane of a converged warp (same reason: no ELECT waterfall around UTMALDG).
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load_3d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

}  // namespace
