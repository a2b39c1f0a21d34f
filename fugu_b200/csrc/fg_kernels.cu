// sm_100a kernels of the fugu query hot path: posting-block decode -> boolean AND/OR ->
// BM25 -> top-k, i.e. what tantivy does inside `searcher.search(&q, &TopDocs::with_limit(k))`
// (/root/reference/src/db/search.rs:162; semantics in SURVEY.md Appendix A).
//
// One CTA evaluates one work item = (query, contiguous doc-id range). It walks the range in
// rounds. A round owns a set of per-doc slots in shared memory (f32 score accumulator + 1-byte
// clause mask):
//   * dense mode: slot = doc - round_lo over a window of DW docs            (frequent terms)
//   * hash mode : slot = open-addressing hash of the doc id, HS slots, the round's doc span is
//                 chosen from the skip tables so that at most HBLK blocks are inserted (rare terms)
// Leaves are applied one after the other (barrier in between), so a slot is updated by at most
// one thread per phase: no floating-point atomics, bit-reproducible sums in leaf order.
// Each leaf phase (a) scans the leaf's 16-byte skip entries lane-parallel, keeps the blocks that
// overlap the round and -- for filter leaves (non-lead Must, Should-under-Must, MustNot) -- whose
// doc range contains a candidate in the round's candidate bitmap, (b) decodes the surviving
// 128-posting blocks one warp per block: bit-unpack 4 gaps + 4 tfs per lane, warp prefix sum,
// fieldnorm gather, BM25, slot update. After the last leaf every slot is tested against the
// query's clause mask and offered to a per-warp register top-k queue ordered like tantivy's
// TopDocs (score desc, doc asc). Per-item lists are merged per query by merge_kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fg_internal.h"

namespace fg {

namespace {

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
    __device__ __forceinline__ void insert(uint64_t c, int k, int lane) {
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
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
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

// any bit set in the inclusive bit range [a, z] of bitmap cb
__device__ __forceinline__ bool cb_any(const uint32_t* cb, uint32_t a, uint32_t z) {
    const uint32_t wa = a >> 5, wz = z >> 5;
    const uint32_t ma = 0xFFFFFFFFu << (a & 31), mz = 0xFFFFFFFFu >> (31 - (z & 31));
    if (wa == wz) return (cb[wa] & ma & mz) != 0;
    if (cb[wa] & ma) return true;
    for (uint32_t w = wa + 1; w < wz; w++)
        if (cb[w]) return true;
    return (cb[wz] & mz) != 0;
}

__device__ __forceinline__ uint32_t hash_doc(uint32_t d) { return (d * 2654435761u) >> (32 - 12); }
static_assert(HS == 4096, "hash_doc assumes 4096 slots");

struct Shared {
    DevLeaf leaf[MAX_LEAVES];
    uint32_t cur[MAX_LEAVES];
    uint32_t cur_next[MAX_LEAVES];
    uint32_t quota[MAX_LEAVES];
    uint32_t wl_count;
    uint32_t rlo, rhi, shift, done;
    uint32_t match;
    unsigned long long st_blocks, st_redecode, st_scored;
};

template <int KS, bool DENSE>
__device__ __forceinline__ void run_item(const SearchParams& p, const DevItem& it, const DevQuery& q,
                                         Shared& S, float* acc, uint32_t* keys, uint8_t* msk,
                                         uint32_t* cb, uint32_t* wl, uint64_t* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nl = (int)q.n_leaves;
    const uint4* __restrict__ skip = p.ix.skip;
    const int k = (int)q.k;

    // ---- per-item init ----
    if (tid < nl) {
        const DevLeaf& L = S.leaf[tid];
        // first block whose last_doc >= doc_lo
        uint32_t a = 0, b = L.n_blocks;
        while (a < b) {
            uint32_t m = (a + b) >> 1;
            uint32_t last = __ldg(&skip[L.blk_begin + m]).x;
            if (last >= it.doc_lo) b = m; else a = m + 1;
        }
        S.cur[tid] = a;
        S.cur_next[tid] = a;
    }
    if (!DENSE && warp == 0) {
        // share HBLK insert blocks per round among the insert leaves, proportional to list length
        uint32_t nb = lane < (int)q.n_insert ? S.leaf[lane].n_blocks : 0;
        uint32_t tot = warp_sum(nb);
        if (lane < (int)q.n_insert)
            S.quota[lane] = 1 + (uint32_t)(((unsigned long long)(HBLK - q.n_insert) * nb) / (tot ? tot : 1));
    }
    for (int i = tid; i < DW; i += NT) acc[i] = 0.f;  // hash keys alias acc[HS..]
    for (int i = tid; i < DW / 4; i += NT) reinterpret_cast<uint32_t*>(msk)[i] = 0;
    __syncthreads();
    if (!DENSE)
        for (int i = tid; i < HS; i += NT) keys[i] = EMPTY;
    if (tid == 0) { S.match = 0; S.st_blocks = 0; S.st_redecode = 0; S.st_scored = 0; }
    __syncthreads();

    WarpTopK<KS> tk;
    tk.init();
    uint32_t lo = it.doc_lo;
    const uint32_t end = it.doc_hi;
    uint32_t my_matches = 0, my_scored = 0;
    unsigned long long my_blocks = 0, my_redecode = 0;

    while (true) {
        // ---- round setup ----
        if (warp == 0) {
            uint32_t fb = EMPTY, chi = end;
            if (lane < (int)q.n_insert) {
                const DevLeaf& L = S.leaf[lane];
                const uint32_t c = S.cur[lane];
                if (c < L.n_blocks) {
                    fb = __ldg(&skip[L.blk_begin + c]).y;
                    if (!DENSE) {
                        const uint32_t idx = c + S.quota[lane] - 1;
                        if (idx < L.n_blocks) chi = min(chi, __ldg(&skip[L.blk_begin + idx]).x + 1);
                    }
                }
            }
            fb = warp_min(fb);
            chi = warp_min(chi);
            if (lane == 0) {
                uint32_t rlo = fb == EMPTY ? end : max(lo, fb);
                uint32_t rhi = DENSE ? (uint32_t)min((unsigned long long)rlo + DW, (unsigned long long)end) : chi;
                S.done = rlo >= end;
                if (rhi < rlo) rhi = rlo;
                S.rlo = rlo;
                S.rhi = rhi;
                const uint32_t span = rhi - rlo;
                int sh = span > 1 ? (32 - __clz(span - 1)) - 13 : 0;
                S.shift = sh > 0 ? sh : 0;
            }
        }
        __syncthreads();
        if (S.done) break;
        const uint32_t rlo = S.rlo, rhi = S.rhi, shift = S.shift;

        // ---- leaf phases ----
        for (int l = 0; l < nl; l++) {
            const DevLeaf L = S.leaf[l];
            const bool filter = L.role != ROLE_INSERT;
            uint32_t base_b = S.cur[l];
            while (true) {
                if (tid == 0) S.wl_count = 0;
                __syncthreads();
                // (a) lane-parallel scan of skip entries
                const uint32_t b = base_b + tid;
                bool in_range = false, needed = false;
                if (b < L.n_blocks) {
                    const uint4 e = __ldg(&skip[L.blk_begin + b]);
                    in_range = e.y < rhi;
                    if (in_range) {
                        if (e.x < rhi) atomicMax(&S.cur_next[l], b + 1);
                        if (e.x >= rlo) {
                            needed = true;
                            if (filter) {
                                const uint32_t a = max(e.y, rlo) - rlo, z = min(e.x, rhi - 1) - rlo;
                                needed = cb_any(cb, a >> shift, z >> shift);
                                if (needed && p.exact_filter && !DENSE) {
                                    // exact accounting: a candidate doc id inside [e.y, e.x]?
                                    bool any = false;
                                    for (int i = 0; i < HS && !any; i++) {
                                        const uint32_t kd = keys[i];
                                        any = kd != EMPTY && kd >= e.y && kd <= e.x &&
                                              (msk[i] & L.req) == L.req;
                                    }
                                    needed = any;
                                }
                            }
                        }
                    }
                }
                const unsigned nm = __ballot_sync(FULL, needed);
                if (nm) {
                    uint32_t pos = 0;
                    if (lane == 0) pos = atomicAdd(&S.wl_count, (uint32_t)__popc(nm));
                    pos = __shfl_sync(FULL, pos, 0);
                    if (needed) wl[pos + __popc(nm & ((1u << lane) - 1u))] = b;
                }
                const int more = __syncthreads_or(tid == NT - 1 && in_range);
                // (b) one warp per surviving block
                const uint32_t nwl = S.wl_count;
                for (uint32_t i = warp; i < nwl; i += NW) {
                    const uint32_t bb = wl[i];
                    const uint4 e = __ldg(&skip[L.blk_begin + bb]);
                    const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
                    const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e.z * 16u);
                    uint32_t g[4], t[4];
                    unpack4(wd, lane, bd, g);
                    unpack4(wd + 4 * bd, lane, bt, t);
                    g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
                    const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
                    if (lane == 0) {
                        const unsigned long long by = ((n * bd + 7) >> 3) + ((n * bt + 7) >> 3) + 16;
                        if (e.y >= lo) my_blocks += by; else my_redecode += by;
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t d = off + g[j] + j;
                        if (4u * lane + j < n && d >= rlo && d < rhi) {
                            int slot;
                            if (DENSE) {
                                slot = (int)(d - rlo);
                                if (filter && (msk[slot] & L.req) != L.req) slot = -1;
                                if (L.role == ROLE_NOT && msk[slot < 0 ? 0 : slot] == 0) slot = -1;
                            } else {
                                uint32_t h = hash_doc(d);
                                slot = -1;
                                while (true) {
                                    uint32_t kd = keys[h];
                                    if (kd == d) { slot = (int)h; break; }
                                    if (kd == EMPTY) {
                                        if (filter) break;
                                        kd = atomicCAS(&keys[h], EMPTY, d);
                                        if (kd == EMPTY || kd == d) { slot = (int)h; break; }
                                    }
                                    h = (h + 1) & (HS - 1);
                                }
                                if (slot >= 0 && filter && (msk[slot] & L.req) != L.req) slot = -1;
                            }
                            if (slot >= 0) {
                                if (L.role == ROLE_NOT) {
                                    msk[slot] |= (uint8_t)BIT_NOT;
                                } else {
                                    const float tf = (float)(t[j] + 1u);
                                    float norm = L.cnorm;
                                    if (L.fn_field >= 0)
                                        norm = __ldg(p.ix.cache + L.fn_field * 256 +
                                                     __ldg(p.ix.fnorm[L.fn_field] + d));
                                    acc[slot] += L.weight * (tf / (tf + norm));
                                    if (L.bit) msk[slot] |= (uint8_t)L.bit;
                                    my_scored++;
                                }
                            }
                        }
                    }
                }
                if (!more) break;
                base_b += NT;
                __syncthreads();
            }
            __syncthreads();
            if (L.build_cb) {
                const uint32_t need = L.build_cb;
                if (DENSE) {
                    if (tid < CBW) {
                        const uint32_t* m32 = reinterpret_cast<const uint32_t*>(msk) + tid * 8;
                        uint32_t word = 0;
#pragma unroll
                        for (int w = 0; w < 8; w++) {
                            const uint32_t mm = m32[w];
#pragma unroll
                            for (int by = 0; by < 4; by++)
                                if ((((mm >> (8 * by)) & 0xFFu) & need) == need) word |= 1u << (w * 4 + by);
                        }
                        cb[tid] = word;
                    }
                } else {
                    if (tid < CBW) cb[tid] = 0;
                    __syncthreads();
                    for (int i = tid; i < HS; i += NT) {
                        const uint32_t kd = keys[i];
                        if (kd != EMPTY && (msk[i] & need) == need) {
                            const uint32_t bit = (kd - rlo) >> shift;
                            atomicOr(&cb[bit >> 5], 1u << (bit & 31));
                        }
                    }
                }
                __syncthreads();
            }
        }

        // ---- scan slots: clause mask test, top-k offer, reset ----
        {
            const int n4 = DENSE ? (int)((rhi - rlo + 3) >> 2) : HS / 4;
            uint32_t* m32 = reinterpret_cast<uint32_t*>(msk);
            for (int g0 = warp * 32; g0 < n4; g0 += NT) {
                const int g = g0 + lane;
                uint32_t m4 = 0;
                uint4 kv = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
                if (g < n4) {
                    m4 = m32[g];
                    if (!DENSE) kv = reinterpret_cast<uint4*>(keys)[g];
                }
                uint64_t cand[4];
                uint64_t best = 0;
                const uint32_t kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    cand[j] = 0;
                    const uint32_t m = (m4 >> (8 * j)) & 0xFFu;
                    const bool occ = DENSE ? (m != 0) : (kk[j] != EMPTY);
                    if (occ) {
                        const int slot = 4 * g + j;
                        const uint32_t doc = DENSE ? rlo + slot : kk[j];
                        const float sc = acc[slot];
                        acc[slot] = 0.f;
                        if (!DENSE) keys[slot] = EMPTY;
                        bool match = (m & 0x7Fu) == q.all_must && !(m & BIT_NOT);
                        if (match && p.ix.alive)
                            match = (__ldg(p.ix.alive + (doc >> 5)) >> (doc & 31)) & 1u;
                        if (match) {
                            my_matches++;
                            cand[j] = make_key(sc + q.const_score, doc);
                            if (cand[j] > best) best = cand[j];
                            if (p.match_bitmap)
                                atomicOr(p.match_bitmap + (size_t)it.query * p.bitmap_words + (doc >> 5),
                                         1u << (doc & 31));
                        }
                    }
                }
                if (m4) m32[g] = 0;
                if (__any_sync(FULL, best > tk.theta)) {
#pragma unroll
                    for (int j = 0; j < 4; j++) tk.offer(cand[j] != 0, cand[j], k, lane);
                }
            }
        }

        if (tid < nl) S.cur[tid] = max(S.cur[tid], S.cur_next[tid]);
        lo = rhi;
        __syncthreads();
    }

    // ---- per-item epilogue: merge the warp queues, write the partial list ----
    my_matches = warp_sum(my_matches);
    my_scored = warp_sum(my_scored);
    if (lane == 0) {
        atomicAdd(&S.match, my_matches);
        atomicAdd(&S.st_scored, (unsigned long long)my_scored);
        atomicAdd(&S.st_blocks, my_blocks);
        atomicAdd(&S.st_redecode, my_redecode);
    }
#pragma unroll
    for (int s = 0; s < KS; s++) scratch[(warp * KS + s) * 32 + lane] = tk.q[s];
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < NW; w++) {
#pragma unroll
            for (int s = 0; s < KS; s++) {
                const uint64_t c = scratch[(w * KS + s) * 32 + lane];
                tk.offer(c != 0, c, k, lane);
            }
        }
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int r = s * 32 + lane;
            if (r < (int)p.kcap) p.partial[(size_t)it.slot * p.kcap + r] = r < k ? tk.q[s] : 0;
        }
        if (lane == 0) {
            p.partial_count[it.slot] = S.match;
            if (p.stats) {
                atomicAdd(p.stats + 0, S.st_blocks);
                atomicAdd(p.stats + 1, S.st_redecode);
                atomicAdd(p.stats + 2, S.st_scored);
            }
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(NT) search_kernel(const SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Shared S;
    float* acc = reinterpret_cast<float*>(smem);
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem) + HS;  // aliases acc[HS..] (hash mode only)
    uint8_t* msk = smem + DW * 4;
    uint32_t* cb = reinterpret_cast<uint32_t*>(smem + DW * 4 + DW);
    uint32_t* wl = cb + CBW;
    uint64_t* scratch = reinterpret_cast<uint64_t*>(wl + NT);

    const DevItem it = p.items[blockIdx.x];
    const DevQuery q = p.queries[it.query];
    if (threadIdx.x < q.n_leaves) S.leaf[threadIdx.x] = p.leaves[q.leaf_begin + threadIdx.x];
    __syncthreads();
    if (it.mode == MODE_DENSE)
        run_item<KS, true>(p, it, q, S, acc, keys, msk, cb, wl, scratch);
    else
        run_item<KS, false>(p, it, q, S, acc, keys, msk, cb, wl, scratch);
}

// one warp per query: merge the per-item partial lists
template <int KS>
__global__ void __launch_bounds__(128) merge_kernel(const MergeParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= p.n_queries) return;
    const DevQuery q = p.queries[qi];
    const int k = (int)q.k;
    WarpTopK<KS> tk;
    tk.init();
    const uint64_t* src = p.partial + (size_t)q.item_begin * p.kcap;
    const uint32_t total = q.n_items * p.kcap;
    uint32_t cnt = 0;
    for (uint32_t i0 = 0; i0 < total; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint64_t c = i < total ? src[i] : 0;
        tk.offer(c != 0, c, k, lane);
    }
    for (uint32_t i = lane; i < q.n_items; i += 32) cnt += p.partial_count[q.item_begin + i];
    cnt = warp_sum(cnt);
    uint32_t nh = 0;
    uint2* out = reinterpret_cast<uint2*>(p.out_hits) + (size_t)qi * p.k_stride;
#pragma unroll
    for (int s = 0; s < KS; s++) {
        const int r = s * 32 + lane;
        const bool ok = r < k && tk.q[s] != 0;
        if (r < (int)p.k_stride) {
            uint2 h = make_uint2(0u, 0xFFFFFFFFu);
            if (ok) {
                h.x = __float_as_uint(unsortable((uint32_t)(tk.q[s] >> 32)));
                h.y = ~(uint32_t)(tk.q[s] & 0xFFFFFFFFu) + p.doc_base;
            }
            out[r] = h;
        }
        nh += __popc(__ballot_sync(FULL, ok));
    }
    if (lane == 0) {
        p.out_n[qi] = nh;
        if (p.out_count) p.out_count[qi] = cnt;
    }
}

// one warp per query: merge the all-gathered per-rank top-k lists
template <int KS>
__global__ void __launch_bounds__(128) merge_gathered_kernel(const uint2* hits, const uint32_t* n,
                                                             uint32_t n_ranks, uint32_t n_queries,
                                                             uint32_t k, uint32_t k_stride,
                                                             uint2* out_hits, uint32_t* out_n) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= n_queries) return;
    WarpTopK<KS> tk;
    tk.init();
    for (uint32_t r = 0; r < n_ranks; r++) {
        const uint32_t cnt = min(n[(size_t)r * n_queries + qi], k_stride);
        const uint2* src = hits + ((size_t)r * n_queries + qi) * k_stride;
        for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint64_t c = 0;
            if (i < cnt) {
                const uint2 h = src[i];
                c = make_key(__uint_as_float(h.x), h.y);
            }
            tk.offer(c != 0, c, (int)k, lane);
        }
    }
    uint32_t nh = 0;
#pragma unroll
    for (int s = 0; s < KS; s++) {
        const int r = s * 32 + lane;
        const bool ok = r < (int)k && tk.q[s] != 0;
        if (r < (int)k_stride) {
            uint2 h = make_uint2(0u, 0xFFFFFFFFu);
            if (ok) {
                h.x = __float_as_uint(unsortable((uint32_t)(tk.q[s] >> 32)));
                h.y = ~(uint32_t)(tk.q[s] & 0xFFFFFFFFu);
            }
            out_hits[(size_t)qi * k_stride + r] = h;
        }
        nh += __popc(__ballot_sync(FULL, ok));
    }
    if (lane == 0) out_n[qi] = nh;
}

}  // namespace

int search_smem_bytes(int ks) {
    return DW * 4 + DW + CBW * 4 + NT * 4 + NW * ks * 32 * 8;
}

template <int KS>
static void launch_search_t(const SearchParams& p, cudaStream_t st) {
    static bool configured = false;
    const int smem = search_smem_bytes(KS);
    if (!configured) {
        cudaFuncSetAttribute(search_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    search_kernel<KS><<<p.n_items, NT, smem, st>>>(p);
}

void launch_search(const SearchParams& p, int ks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p.n_items == 0) return;
    if (ks <= 1) launch_search_t<1>(p, st);
    else launch_search_t<4>(p, st);
}

void launch_merge(const MergeParams& p, int ks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p.n_queries == 0) return;
    const unsigned grid = (p.n_queries + 3) / 4;
    if (ks <= 1) merge_kernel<1><<<grid, 128, 0, st>>>(p);
    else merge_kernel<4><<<grid, 128, 0, st>>>(p);
}

void launch_merge_gathered(const void* hits, const uint32_t* n, uint32_t n_ranks, uint32_t n_queries,
                           uint32_t k, uint32_t k_stride, void* out_hits, uint32_t* out_n, int ks,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_queries == 0) return;
    const unsigned grid = (n_queries + 3) / 4;
    if (ks <= 1)
        merge_gathered_kernel<1><<<grid, 128, 0, st>>>((const uint2*)hits, n, n_ranks, n_queries, k,
                                                       k_stride, (uint2*)out_hits, out_n);
    else
        merge_gathered_kernel<4><<<grid, 128, 0, st>>>((const uint2*)hits, n, n_ranks, n_queries, k,
                                                       k_stride, (uint2*)out_hits, out_n);
}

}  // namespace fg
