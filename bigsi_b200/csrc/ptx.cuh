// Thin inline-PTX helpers for sm_100a: mbarrier, 1-D bulk async copy (TMA, SASS UBLKCP),
// LOP3, cache-hinted vector loads.  No CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bigsi {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier.init visible to the async proxy (TMA) before first use
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// all state spaces: generic-proxy writes to GLOBAL memory (made visible by a fence / barrier) are ordered
// before this thread's later async-proxy (bulk copy) reads of them
__device__ __forceinline__ void fence_proxy_async_all()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t tx_bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine;
// SASS: UBLKCP).  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// same with an L2 eviction-priority hint (rows are streamed once: evict_first)
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint4 lds128(const void *p)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}
// streaming 256-bit global load (sm_100: LDG.256), no L1 allocation; p 32-byte aligned
__device__ __forceinline__ void ldg256_stream(const void *p, uint32_t (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// 256-bit global store (STG.256): one full 32-byte sector per lane -- no partial-sector write, no fill read
__device__ __forceinline__ void stg256(void *p, const uint32_t (&v)[8])
{
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// streaming 128-bit global load, no L1 allocation
__device__ __forceinline__ uint4 ldg128_stream(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

template <int IMM>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(IMM));
    return d;
}
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return lop3<0x96>(a, b, c); }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return lop3<0xE8>(a, b, c); }
__device__ __forceinline__ uint32_t and3(uint32_t a, uint32_t b, uint32_t c) { return lop3<0x80>(a, b, c); }

// Named barriers with COMPILE-TIME ids: with a register operand ptxas reserves all 16 hardware barriers of the SM
// for the kernel ("used 16 barriers"), and no other CTA that needs even one barrier can become resident beside it --
// which is exactly what the streamed path needs (a reduce CTA beside every gather CTA).
template <int ID>
__device__ __forceinline__ void named_bar_sync(int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
// signal a named barrier without waiting on it (the waiting side uses named_bar_sync with the same count)
template <int ID>
__device__ __forceinline__ void named_bar_arrive(int nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Sticky error word of a handle: the first bounded wait that times out stores (query number << 8 | code) in the
// device word and in its mapped host mirror; every later kernel of the handle exits at once when it sees it.
__device__ __forceinline__ void raise_abort(unsigned long long *abort_word, unsigned long long *host_abort, uint32_t code,
                                            unsigned long long seq)
{
    if (!abort_word) return;
    const unsigned long long v = (seq << 8) | code;
    if (atomicCAS(abort_word, 0ull, v) == 0ull && host_abort) {
        *reinterpret_cast<volatile unsigned long long *>(host_abort) = v;
        __threadfence_system();
    }
}
// Spin until cond() holds, but never longer than timeout_ns and never past an abort raised by somebody else.
// Returns false when it gave up (after raising the abort word with `code`): the caller carries on with whatever
// it has -- results of an aborted handle are never reported, the host turns the abort word into an error code.
template <typename Cond>
__device__ __forceinline__ bool bounded_wait(unsigned long long *abort_word, unsigned long long *host_abort,
                                             unsigned long long timeout_ns, uint32_t code, unsigned long long seq, Cond cond)
{
    if (cond()) return true;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (cond()) return true;
        if ((spins & 255u) == 0) {
            if (abort_word && ld_volatile_u64(abort_word) != 0ull) return false;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) {
                raise_abort(abort_word, host_abort, code, seq);
                return false;
            }
        }
    }
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace bigsi
