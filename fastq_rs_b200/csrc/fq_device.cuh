// fq_device.cuh -- device-side helpers shared by the sm_100a kernels (PTX wrappers for the TMA bulk
// copy + mbarrier, the SWAR newline detector, warp reductions).
#pragma once
#include "fq_common.cuh"

namespace fq {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 prefetch of a global range (16-byte granular; no completion to wait for)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// potentially blocking probe: the warp is suspended until the phase completes or the time hint
// (ns) runs out, so waiting warps do not burn issue slots
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the warp for a system-defined time)
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 0x80 in every byte of w that equals '\n' (exact: no false positives, no cross-byte carries)
__device__ __forceinline__ uint32_t nlbits(uint32_t w)
{
    uint32_t y = (w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu;
    uint32_t t = y + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}
// 16-bit mask, bit i = byte i of the 16-byte piece is '\n'
__device__ __forceinline__ uint32_t nlmask16(const uint4& v)
{
    uint32_t m0 = nlbits(v.x), m1 = nlbits(v.y), m2 = nlbits(v.z), m3 = nlbits(v.w);
    uint32_t c01 = m1 | (m0 >> 4);
    uint32_t c23 = m3 | (m2 >> 4);
    // multiply-gather: bits {3,11,19,27} -> 24..27, bits {7,15,23,31} -> 28..31
    uint32_t b01 = (c01 * 0x00204081u) >> 24;
    uint32_t b23 = (c23 * 0x00204081u) >> 24;
    return b01 | (b23 << 8);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

struct Acc {  // per-thread accumulators, reduced once at the end of the kernel
    unsigned long long n_records, n_bases, clip_seq, clip_qual;
};

}  // namespace fq
