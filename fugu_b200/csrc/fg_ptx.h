// Inline-PTX primitives and launch spellings of the sm_100a kernels. The kernel sources include this
// header as <fg_ptx.h>; nothing else in the product knows about any other implementation of it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FG_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define FG_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// 2^23 as an opaque register value: PRMT keeps a register operand, no constant folding
#define FG_MAGIC_2P23(m) asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(m))

namespace fg {

// nanosecond timer shared by all SMs (dev tool: per-item timing of the lead kernel)
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// 1 / x with MUFU.RCP (<= 1 ulp; callers keep x >= 0.3)
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// PRMT with an immediate selector 0x765J: byte J of w into the low byte of `magic`
template <int J>
__device__ __forceinline__ uint32_t prmt_byte(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7650 | J));
    return r;
}
// shared-memory atomic add through a 32-bit shared address (ATOMS.ADD without generic addressing)
__device__ __forceinline__ uint32_t smem_atomic_inc(uint32_t* p) {
    uint32_t r;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(a) : "memory");
    return r;
}

// start bringing the 128-byte line at p into L2 (no register, no fault on a bad address)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- 1-D bulk copies global -> shared memory (TMA engine, SASS UBLKCP) guarded by an mbarrier ----
// A posting block's payload is 16*(bd+bt) contiguous bytes, 16-byte aligned: one bulk copy stages it.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// arm the barrier with the byte count of the copy that follows, then issue the copy (one thread)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
// block until the barrier's phase with the given parity has completed
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity)
        : "memory");
}

}  // namespace fg
