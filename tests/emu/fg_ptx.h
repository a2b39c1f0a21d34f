// TEST INFRASTRUCTURE ONLY (tests/emu): plain-C++ twins of the product's inline-PTX primitives
// (fugu_b200/csrc/fg_ptx.h). The emulated build puts this directory first on the include path, so the
// kernel sources' `#include <fg_ptx.h>` resolves here; the product build never sees this file.
#pragma once
#include "cuda_runtime.h"

#define FG_DYN_SMEM(name) unsigned char* name = fgemu::dyn_smem()
#define FG_LAUNCH(kernel, grid, block, smem, stream, ...) (fgemu::g_kernel_name = #kernel, fgemu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); }))
#define FG_MAGIC_2P23(m) m = 0x4B000000u

namespace fg {

inline unsigned long long global_timer_ns() { return 0ull; }
inline float rcp_approx(float x) { return 1.0f / x; }
template <int J>
inline uint32_t prmt_byte(uint32_t w, uint32_t magic) { return __byte_perm(w, magic, 0x7650u | (uint32_t)J); }
inline uint32_t smem_atomic_inc(uint32_t* p) { return atomicAdd(p, 1u); }

inline void prefetch_l2(const void*) {}

// mbarrier + bulk copy: the copy completes at once; the barrier word counts completed phases
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void fence_mbar_init() {}
inline void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    if (bytes & 15u) fgemu::die("bulk copy size must be a multiple of 16");
    if (((uintptr_t)smem_dst | (uintptr_t)gsrc) & 15u) fgemu::die("bulk copy addresses must be 16-byte aligned");
    memcpy(smem_dst, gsrc, bytes);
    *bar += 1;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((uint32_t)(*bar & 1u) == (parity & 1u)) fgemu::yield();
}

}  // namespace fg
