// Device helpers shared by the kernel translation units (fg_kernels.cu, fg_lead.cu): the TopDocs key
// order, the per-warp register top-k queue, warp reductions and the horizontal bit-stream unpack of a
// posting block (layout in fg_internal.h).
#pragma once
#include <fg_ptx.h>
#include <stdint.h>

#include "fg_internal.h"

namespace fg {
namespace dev {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t sortable(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unsortable(uint32_t s) {
    uint32_t b = (s & 0x80000000u) ? (s & 0x7FFFFFFFu) : ~s;
    return __uint_as_float(b);
}
// larger key = better hit: score descending, then doc ascending
__device__ __forceinline__ uint64_t make_key(float score, uint32_t doc) {
    return ((uint64_t)sortable(score) << 32) | (uint32_t)(~doc);
}

// tf / (tf + norm) with one MUFU.RCP + one FMUL. __fdividef also guards denominators below 2^-126
// (two compares + four scalings per call); here tf + norm >= 0.3. Error <= 2 ulp, inside the 1e-5 bar.
__device__ __forceinline__ float tf_factor(float t, float n) {
    return t * rcp_approx(t + n);
}

// Per-warp top-k in registers: rank r lives in lane r%32, row r/32; sorted best first.
template <int KS>
struct WarpTopK {
    uint64_t q[KS];
    uint64_t theta;  // a candidate must be > theta to enter (key of rank k-1, 0 while not full)

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < KS; s++) q[s] = 0;
        theta = 0;
    }
    __device__ __forceinline__ void insert_body(uint64_t c, int k, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < KS; s++) pos += __popc(__ballot_sync(FULL, q[s] > c));
        uint64_t carry = 0;
#pragma unroll
        for (int s = 0; s < KS; s++) {
            uint64_t up = __shfl_up_sync(FULL, q[s], 1);
            uint64_t last = __shfl_sync(FULL, q[s], 31);
            if (lane == 0) up = carry;
            int r = s * 32 + lane;
            q[s] = r < pos ? q[s] : (r == pos ? c : up);
            carry = last;
        }
        uint64_t t = 0;
#pragma unroll
        for (int s = 0; s < KS; s++) {
            uint64_t v = __shfl_sync(FULL, q[s], (k - 1) & 31);
            if (s == ((k - 1) >> 5)) t = v;
        }
        theta = t;
    }
    // KS = 32 (deep pagination) would inline 32-row shifts at every offer() site (1.4 MB of SASS per
    // kernel, minutes of compile time): call it out of line there; the hot KS <= 4 variants inline it
    __device__ __noinline__ void insert_call(uint64_t c, int k, int lane) { insert_body(c, k, lane); }
    __device__ __forceinline__ void insert(uint64_t c, int k, int lane) {
        if (KS > 4) insert_call(c, k, lane); else insert_body(c, k, lane);
    }
    // warp-collective: every lane may bring one candidate
    __device__ __forceinline__ void offer(bool valid, uint64_t key, int k, int lane) {
        unsigned m = __ballot_sync(FULL, valid && key > theta);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            uint64_t c = __shfl_sync(FULL, key, src);
            if (c > theta) insert(c, k, lane);
        }
    }
};

__device__ __forceinline__ uint32_t warp_min(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane) {
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    return x - v;
}

// 4 consecutive values of a horizontal little-endian bit stream: values 4*lane .. 4*lane+3
__device__ __forceinline__ void unpack4(const uint32_t* __restrict__ w, int lane, uint32_t b,
                                        uint32_t v[4]) {
    if (b == 0) {
        v[0] = v[1] = v[2] = v[3] = 0;
        return;
    }
    const uint32_t bit0 = (uint32_t)lane * 4u * b;
    if (b <= 8) {
        const uint32_t wi = bit0 >> 5, sh = bit0 & 31;
        const uint32_t lo = __ldg(w + wi), hi = __ldg(w + wi + 1);
        const uint32_t x = __funnelshift_r(lo, hi, sh);
        const uint32_t m = (1u << b) - 1u;
        v[0] = x & m;
        v[1] = (x >> b) & m;
        v[2] = (x >> (2 * b)) & m;
        v[3] = (x >> (3 * b)) & m;
    } else if (b <= 16) {
        // 4 values = 4*b <= 64 bits starting at bit0: at most three consecutive words (sparse lists: 9..14 bits)
        const uint32_t wi = bit0 >> 5, sh = bit0 & 31;
        const uint32_t w0 = __ldg(w + wi), w1 = __ldg(w + wi + 1), w2 = __ldg(w + wi + 2);
        const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);  // bits [bit0, bit0+64)
        const uint32_t m = (1u << b) - 1u;
        v[0] = x0 & m;
        v[1] = __funnelshift_r(x0, x1, b) & m;
        v[2] = __funnelshift_rc(x0, x1, 2 * b) & m;  // (clamped form: 2*b may be 32)
        v[3] = 3 * b >= 32 ? (x1 >> (3 * b - 32)) & m : __funnelshift_r(x0, x1, 3 * b) & m;
    } else {
        const uint32_t m = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t bp = bit0 + j * b;
            const uint32_t wi = bp >> 5, sh = bp & 31;
            const uint32_t lo = __ldg(w + wi), hi = __ldg(w + wi + 1);
            v[j] = __funnelshift_r(lo, hi, sh) & m;
        }
    }
}

// the same from a payload staged in shared memory (plain loads; the staging slot is padded by one word)
__device__ __forceinline__ void unpack4_smem(const uint32_t* w, int lane, uint32_t b, uint32_t v[4]) {
    if (b == 0) {
        v[0] = v[1] = v[2] = v[3] = 0;
        return;
    }
    const uint32_t bit0 = (uint32_t)lane * 4u * b;
    if (b <= 8) {
        const uint32_t wi = bit0 >> 5, sh = bit0 & 31;
        const uint32_t x = __funnelshift_r(w[wi], w[wi + 1], sh);
        const uint32_t m = (1u << b) - 1u;
        v[0] = x & m;
        v[1] = (x >> b) & m;
        v[2] = (x >> (2 * b)) & m;
        v[3] = (x >> (3 * b)) & m;
    } else if (b <= 16) {
        const uint32_t wi = bit0 >> 5, sh = bit0 & 31;
        const uint32_t w0 = w[wi], w1 = w[wi + 1], w2 = w[wi + 2];
        const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
        const uint32_t m = (1u << b) - 1u;
        v[0] = x0 & m;
        v[1] = __funnelshift_r(x0, x1, b) & m;
        v[2] = __funnelshift_rc(x0, x1, 2 * b) & m;
        v[3] = 3 * b >= 32 ? (x1 >> (3 * b - 32)) & m : __funnelshift_r(x0, x1, 3 * b) & m;
    } else {
        const uint32_t m = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t bp = bit0 + j * b;
            const uint32_t wi = bp >> 5, sh = bp & 31;
            v[j] = __funnelshift_r(w[wi], w[wi + 1], sh) & m;
        }
    }
}

}  // namespace dev
}  // namespace fg
