// Thin inline-PTX wrappers for the Blackwell (sm_100a) asynchronous machinery used by the dense-BA kernels:
// mbarrier, 1-D bulk copies (cp.async.bulk, SASS UBLKCP), tensor memory (tcgen05.alloc/ld, SASS LDTM) and the
// 5th-generation tensor-core MMA (tcgen05.mma kind::tf32, SASS UTCMMA) with shared-memory operand descriptors.
// Bit layouts of the descriptors follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vba {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out, so a waiting warp
// takes (almost) no issue slots
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
// A wait that can never complete (a pipeline bug) must not hang the GPU: after ~4 s it reports and traps.
static __device__ __noinline__ void mbar_timeout(uint64_t *bar, uint32_t parity) {
    printf("mbarrier wait timed out: block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
    __trap();
}
// The polling loops are kept to a handful of instructions (they compete for issue slots with the warps doing the work); a
// wait that can never complete traps after ~2^26 polls instead of reading the clock on every poll.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) mbar_timeout(bar, parity);
    }
}

// for roles that expect to wait long (MMA issuer, flush warps): back off between polls
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(100);
        if (++spins > (1u << 24)) mbar_timeout(bar, parity);
    }
}

// ------------------------------------------------------------------------------------------------ bulk copy (TMA, 1-D)
// global -> shared, completion counted in bytes on `bar`.  dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core / bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ tensor memory
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the one that allocated)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive columns: thread i of the warp receives lane (base_lane + i), columns [col, col + 16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ tcgen05.mma
// Shared-memory operand descriptor, K-major, 128-byte swizzle: rows of 128 bytes (32 tf32), 8-row groups 1024 bytes
// apart (SBO), the 16-byte chunk index of a row XORed with (row % 8).  The operand base must be 1024-byte aligned;
// K advances inside the swizzle atom by adding the byte offset to the start address.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups, bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
    return d;
}
// physical byte offset of element (row, k) inside one [rows x 32] tf32 swizzle-128B K-major block
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}
// instruction descriptor: tf32 x tf32 -> fp32, both operands K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4)      // D format fp32
           | (2u << 7)    // A format tf32
           | (2u << 10)   // B format tf32
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one K = 8 step.  Issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// fp32 -> (hi, lo): hi = x rounded to tf32's 10 mantissa bits (round half away, two integer ops; cvt.rna.tf32 is emulated
// with a dozen), lo = x - hi exactly (<= 13 significant bits, of which the tensor core keeps the leading 11:
// 2^-22 |x| relative).  Finite inputs only.
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

}  // namespace sm100
}  // namespace vba
