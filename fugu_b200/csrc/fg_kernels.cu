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
// The leaves of one clause form a phase (barrier between phases); inside a phase all leaves are
// applied concurrently with shared-memory float atomics (a+b is exact-commutative; only docs with
// >= 3 contributions in one clause can differ in the last bit between runs, far inside the 1e-5
// parity tolerance). Each phase (a) scans the leaves' 16-byte skip entries, one warp per leaf and
// 32 entries per step, keeps the blocks that overlap the round and -- for filter clauses (non-lead
// Must, Should-under-Must, MustNot) -- whose doc range contains a candidate in the round's
// candidate bitmap, (b) decodes the surviving 128-posting blocks GRP at a time per warp: bit-unpack
// 4 gaps + 4 tfs per lane, warp prefix sum, all fieldnorm gathers issued together, BM25, slot update. After the last leaf every slot is tested against the
// query's clause mask and offered to a per-warp register top-k queue ordered like tantivy's
// TopDocs (score desc, doc asc). Per-item lists are merged per query by merge_kernel.
#include <fg_ptx.h>
#include <type_traits>
#include <stdint.h>
#include <stdlib.h>

#include "fg_device.h"
#include "fg_internal.h"

namespace fg {

using namespace dev;

namespace {


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

// Bucketised open addressing: the hash picks a 16-byte bucket of 4 keys (one LDS.128), an insert
// takes the first empty key of the bucket and moves to the next bucket only when it is full. Far
// fewer dependent probes than per-slot linear probing (a full bucket at load 0.6 is rare).
__device__ __forceinline__ uint32_t hash_bucket(uint32_t d) { return (d * 2654435761u) >> (32 - (HS_LOG2 - 2)); }

// claim (or find) the slot of doc d; the table never fills (rounds insert <= HBLK*128 + RES_CAP docs)
__device__ __forceinline__ int hash_insert(uint32_t* keys, uint32_t d) {
    uint32_t h = hash_bucket(d);
    while (true) {
        const uint4 k = reinterpret_cast<const uint4*>(keys)[h];
        if (k.x == d) return (int)(4 * h);
        if (k.y == d) return (int)(4 * h + 1);
        if (k.z == d) return (int)(4 * h + 2);
        if (k.w == d) return (int)(4 * h + 3);
        const int e = k.x == EMPTY ? 0 : k.y == EMPTY ? 1 : k.z == EMPTY ? 2 : k.w == EMPTY ? 3 : -1;
        if (e >= 0) {
            const uint32_t old = atomicCAS(&keys[4 * h + e], EMPTY, d);
            if (old == EMPTY || old == d) return (int)(4 * h + e);
            continue;  // somebody else took that key: look at the same bucket again
        }
        h = (h + 1) & (HS / 4 - 1);
    }
}
// slot of doc d or -1 (a bucket with an empty key ends the search: inserts fill buckets in order)
__device__ __forceinline__ int hash_find(const uint32_t* keys, uint32_t d) {
    uint32_t h = hash_bucket(d);
    while (true) {
        const uint4 k = reinterpret_cast<const uint4*>(keys)[h];
        if (k.x == d) return (int)(4 * h);
        if (k.y == d) return (int)(4 * h + 1);
        if (k.z == d) return (int)(4 * h + 2);
        if (k.w == d) return (int)(4 * h + 3);
        if (k.w == EMPTY) return -1;  // keys fill x, y, z, w in order
        h = (h + 1) & (HS / 4 - 1);
    }
}

constexpr int RES_CAP_ = 512;
static_assert(MAX_LEAVES * BLOCK + RES_CAP_ <= HS * 5 / 8, "hash rounds must never fill the table");
constexpr int SEG_CAP = 128;  // worklist entries per scanning warp
constexpr uint32_t LEAF_DONE = 0xFFFFFFFFu;
constexpr int RES_CAP = RES_CAP_;        // resident postings per work item (short lists decoded once per item)
constexpr int RES_MAX_BLOCKS = 3;   // a leaf is resident when it has at most this many blocks in the item's range

struct Shared {
    DevQuery q;
    DevLeaf leaf[MAX_LEAVES];
    uint32_t cur[MAX_LEAVES];
    uint32_t cur_next[MAX_LEAVES];
    uint32_t quota[MAX_LEAVES];
    uint32_t resume[MAX_LEAVES * NW];
    uint32_t phase_end[MAX_LEAVES];
    uint32_t res_cnt[MAX_LEAVES];   // blocks of the leaf inside the item's doc range (init only)
    uint32_t resident[MAX_LEAVES];
    uint32_t res_n;
    uint32_t segcnt[NW];
    uint32_t leafmask;  // leaves with at least one needed block in the current scan pass
    uint32_t rlo, rhi, shift, done, gtheta;
    uint32_t match;
    uint32_t theta_cta;  // sortable f32: best k-th score any warp of this CTA has reached (shared pre-test threshold)
    int32_t col_field;  // >= 0: every column leaf takes its norms from this fieldnorm field (ctab holds its cache)
    float ctab[256];    // BM25 norm cache K1*(1-B+B*dl/avg) of col_field, by fieldnorm id
    unsigned long long st_blocks, st_redecode, st_scored;
};

// byte j of w as a float (PRMT into the mantissa of 2^23, one FADD): column tfs without I2F
__device__ __forceinline__ float byte_f32(uint32_t w, int j) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)j)) - 8388608.0f;
}
// same with the 2^23 pattern held in a register, so that the selector is an immediate of PRMT
template <int J>
__device__ __forceinline__ float byte_f32_r(uint32_t w, uint32_t magic) {
    return __uint_as_float(prmt_byte<J>(w, magic)) - 8388608.0f;
}
__device__ __forceinline__ void col_term8(uint2 t, float w, uint32_t magic, const float n[8], float v[8]) {
    v[0] += w * tf_factor(byte_f32_r<0>(t.x, magic), n[0]);
    v[1] += w * tf_factor(byte_f32_r<1>(t.x, magic), n[1]);
    v[2] += w * tf_factor(byte_f32_r<2>(t.x, magic), n[2]);
    v[3] += w * tf_factor(byte_f32_r<3>(t.x, magic), n[3]);
    v[4] += w * tf_factor(byte_f32_r<0>(t.y, magic), n[4]);
    v[5] += w * tf_factor(byte_f32_r<1>(t.y, magic), n[5]);
    v[6] += w * tf_factor(byte_f32_r<2>(t.y, magic), n[6]);
    v[7] += w * tf_factor(byte_f32_r<3>(t.y, magic), n[7]);
}
// number of non-zero bytes of x
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t x) {
    return __popc((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u);
}

__device__ __forceinline__ void smem_add_f32(float* p, float v) { atomicAdd(p, v); }

// ---- streamed leaves ---------------------------------------------------------------------------
// A sparse insert leaf of a dense-window pure union (a few blocks per window at most) is walked by ONE
// warp, block after block, without skip scan / worklist / phase barriers: decode, score, float atomics.
__device__ __forceinline__ uint32_t stream_block(const SearchParams& p, const DevLeaf& L, const uint4 e,
                                                 uint32_t rlo, uint32_t rhi, float* acc, int lane) {
    const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
    const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e.z * 16u);
    uint32_t g[4], t[4];
    unpack4(wd, lane, bd, g);
    unpack4(wd + 4 * bd, lane, bt, t);
    g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
    const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
    uint32_t d[4];
    bool ok[4];
    float nrm[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        d[j] = off + g[j] + j;
        ok[j] = 4u * lane + j < n && d[j] >= rlo && d[j] < rhi;
    }
    const int ff = L.fn_field;
    const uint8_t* fnp = p.ix.fnorm[ff < 0 ? 0 : ff];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        nrm[j] = L.cnorm;
        if (ok[j] && ff >= 0) nrm[j] = __ldg(p.ix.cache + ff * 256 + __ldg(fnp + d[j]));
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (ok[j]) {
            smem_add_f32(&acc[d[j] - rlo], L.weight * tf_factor((float)(t[j] + 1u), nrm[j]));
            cnt++;
        }
    return cnt;
}

// ---- column leaves (dense tf columns, fg_api.cu: fg_index_upload) ---------------------------------
// Pure-union dense window: add the scores of every column leaf to the 4 docs rlo + 4g .. +3.
// A column byte of 0 (term absent) contributes exactly 0 (0 / (0 + norm), norm > 0).
__device__ __forceinline__ void col_add_dense(const Shared& S, const SearchParams& p, int nl, int ncol,
                                              uint32_t doc0, float v[4]) {
    if (S.col_field >= 0) {
        const uint32_t fn4 = __ldg(reinterpret_cast<const uint32_t*>(p.ix.fnorm[S.col_field] + doc0));
        const DevLeaf* L = &S.leaf[nl];
        int c = 0;
        uint32_t ta = __ldg(reinterpret_cast<const uint32_t*>(L[0].col + doc0)), tb = 0;
        if (ncol > 1) tb = __ldg(reinterpret_cast<const uint32_t*>(L[1].col + doc0));
        float n[4];
#pragma unroll
        for (int j = 0; j < 4; j++) n[j] = S.ctab[(fn4 >> (8 * j)) & 255u];
        while (true) {
            {
                const float w = L[c].weight;
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] += w * tf_factor(byte_f32(ta, j), n[j]);
            }
            if (c + 1 >= ncol) break;
            {
                const float w = L[c + 1].weight;
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] += w * tf_factor(byte_f32(tb, j), n[j]);
            }
            c += 2;
            if (c >= ncol) break;
            ta = __ldg(reinterpret_cast<const uint32_t*>(L[c].col + doc0));
            if (c + 1 < ncol) tb = __ldg(reinterpret_cast<const uint32_t*>(L[c + 1].col + doc0));
        }
    } else {  // column leaves of different fields (facet columns have a constant norm)
        for (int c = 0; c < ncol; c++) {
            const DevLeaf& L = S.leaf[nl + c];
            const uint32_t t4 = __ldg(reinterpret_cast<const uint32_t*>(L.col + doc0));
            uint32_t fn4 = 0;
            if (L.fn_field >= 0) fn4 = __ldg(reinterpret_cast<const uint32_t*>(p.ix.fnorm[L.fn_field] + doc0));
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float n = L.fn_field >= 0 ? __ldg(p.ix.cache + L.fn_field * 256 + ((fn4 >> (8 * j)) & 255u)) : L.cnorm;
                const float t = byte_f32(t4, j);
                v[j] += L.weight * tf_factor(t, n);
            }
        }
    }
}

// Masked plans (and pure unions that filter deleted docs): apply every column leaf to 4 slots.
// DENSE: slots are the docs doc0 .. doc0+3; hash: slot j holds doc kk[j]. A slot takes part when
// ok[j]; a present term adds its score and sets its clause bit (or BIT_NOT).
template <bool DENSE>
__device__ __forceinline__ void col_apply(const Shared& S, const SearchParams& p, int nl, int ncol,
                                          uint32_t doc0, const uint32_t kk[4], const bool ok[4],
                                          float v[4], uint32_t m[4]) {
    if (!(ok[0] || ok[1] || ok[2] || ok[3])) return;
    float n[4] = {0.f, 0.f, 0.f, 0.f};
    const int cf = S.col_field;
    if (cf >= 0) {
        if (DENSE) {
            const uint32_t fn4 = __ldg(reinterpret_cast<const uint32_t*>(p.ix.fnorm[cf] + doc0));
#pragma unroll
            for (int j = 0; j < 4; j++) n[j] = S.ctab[(fn4 >> (8 * j)) & 255u];
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (ok[j]) n[j] = S.ctab[__ldg(p.ix.fnorm[cf] + kk[j])];
        }
    }
    for (int c = 0; c < ncol; c++) {
        const DevLeaf& L = S.leaf[nl + c];
        uint32_t t4 = 0;
        if (DENSE) t4 = __ldg(reinterpret_cast<const uint32_t*>(L.col + doc0));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (!ok[j]) continue;
            const uint32_t tb = DENSE ? (t4 >> (8 * j)) & 255u : (uint32_t)__ldg(L.col + kk[j]);
            if (!tb) continue;
            if (L.role == ROLE_NOT) { m[j] |= BIT_NOT; continue; }
            float nn = n[j];
            if (cf < 0) {
                const uint32_t d = DENSE ? doc0 + j : kk[j];
                nn = L.fn_field >= 0 ? __ldg(p.ix.cache + L.fn_field * 256 + __ldg(p.ix.fnorm[L.fn_field] + d)) : L.cnorm;
            }
            const float t = (float)tb;
            v[j] += L.weight * tf_factor(t, nn);
            m[j] |= L.bit;
        }
    }
}

// One work item. DENSE: slot = doc - round_lo; else hash slots. PURE: the plan is a plain union
// (only Should clauses, positive weights): no clause masks, every touched slot matches.
template <int KS, int GRP, bool DENSE, bool PURE>
__device__ __forceinline__ void run_item(const SearchParams& p, const DevItem& it, const DevQuery& q,
                                         Shared& S, float* acc, uint32_t* keys, uint8_t* msk,
                                         uint32_t* cb, uint32_t* wl, uint64_t* scratch, uint32_t* res_doc,
                                         float* res_val) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nl = (int)q.n_leaves;   // block leaves; the q.n_col column leaves follow them in S.leaf
    const int ncol = (int)q.n_col;
    const int nlp = nl - (int)q.n_stream;  // leaves evaluated in clause phases; [nlp, nl) are streamed
    const uint4* __restrict__ skip = p.ix.skip;
    const int k = (int)q.k;
    const unsigned lt_mask = (1u << lane) - 1u;

    // ---- per-item init ----
    if (tid < nl) {
        const DevLeaf& L = S.leaf[tid];
        uint32_t a = 0, b = L.n_blocks;  // first block whose last_doc >= doc_lo
        while (a < b) {
            uint32_t m = (a + b) >> 1;
            if (__ldg(&skip[L.blk_begin + m]).x >= it.doc_lo) b = m; else a = m + 1;
        }
        S.cur[tid] = a;
        S.cur_next[tid] = a;
        int l1 = tid + 1;  // leaves [tid, phase_end) share (role, bit) = one clause phase
        while (!p.deterministic && !L.solo && l1 < nlp && S.leaf[l1].role == L.role && S.leaf[l1].bit == L.bit &&
               !S.leaf[l1].solo) l1++;
        S.phase_end[tid] = (uint32_t)l1;
        // blocks of an insert leaf that overlap the item's doc range (few => resident)
        uint32_t cnt = 0xFFFFu;
        if (L.role == ROLE_INSERT && !p.deterministic) {
            uint32_t a2 = a, b2 = L.n_blocks;  // first block whose first_base >= doc_hi
            while (a2 < b2) {
                uint32_t m = (a2 + b2) >> 1;
                if (__ldg(&skip[L.blk_begin + m]).y >= it.doc_hi) b2 = m; else a2 = m + 1;
            }
            cnt = a2 - a;
        }
        S.res_cnt[tid] = cnt;
    }
    if (tid == 0) {
        S.res_n = 0; S.match = 0; S.theta_cta = 0; S.st_blocks = 0; S.st_redecode = 0; S.st_scored = 0;
        int cf = ncol ? S.leaf[nl].fn_field : -1;
        for (int c = 1; c < ncol; c++) if (S.leaf[nl + c].fn_field != cf) cf = -1;
        S.col_field = cf;
    }
    __syncthreads();
    if (ncol && S.col_field >= 0) S.ctab[tid] = __ldg(p.ix.cache + S.col_field * 256 + tid);
    if (warp == 0) {
        // resident leaves: short lists are decoded ONCE per item into (doc, score) pairs and applied to
        // every round from shared memory, instead of re-decoding their straddling blocks each round
        const bool ins = lane < (int)q.n_insert;
        const uint32_t c = ins ? S.res_cnt[lane] : 0xFFFFu;
        const bool cand = c <= (uint32_t)RES_MAX_BLOCKS;
        uint32_t pre = cand ? c * BLOCK : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, pre, o);
            if (lane >= o) pre += y;
        }
        const bool res = cand && pre <= (uint32_t)RES_CAP;
        if (lane < nl) S.resident[lane] = res ? 1u : 0u;
        // hash mode: share HBLK insert blocks per round among the STREAMING insert leaves, by list length
        const uint32_t nb = (ins && !res) ? S.leaf[lane].n_blocks : 0u;
        const uint32_t tot = warp_sum(nb);
        const uint32_t n_stream = __popc(__ballot_sync(FULL, ins && !res));
        // every streaming leaf gets at least one block; more than HBLK leaves still fit the table
        // (MAX_LEAVES * 128 + RES_CAP <= HS * 5 / 8)
        const uint32_t extra = n_stream < (uint32_t)HBLK ? (uint32_t)HBLK - n_stream : 0u;
        if (ins) S.quota[lane] = 1 + (uint32_t)(((unsigned long long)extra * nb) / (tot ? tot : 1));
    }
    __syncthreads();
    for (int l = 0; l < (int)q.n_insert; l++) {
        if (!S.resident[l]) continue;
        const DevLeaf& L = S.leaf[l];
        const uint32_t b0 = S.cur[l], b1 = b0 + S.res_cnt[l];
        for (uint32_t bb = b0 + warp; bb < b1; bb += NW) {
            const uint4 e = __ldg(&skip[L.blk_begin + bb]);
            const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
            const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e.z * 16u);
            uint32_t g[4], t[4];
            unpack4(wd, lane, bd, g);
            unpack4(wd + 4 * bd, lane, bt, t);
            g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
            const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
            uint32_t scored_here = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t d = off + g[j] + j;
                const bool ok = 4u * lane + j < n && d >= it.doc_lo && d < it.doc_hi;
                const unsigned m = __ballot_sync(FULL, ok);
                uint32_t pos = 0;
                if (lane == 0 && m) pos = atomicAdd(&S.res_n, (uint32_t)__popc(m));
                pos = __shfl_sync(FULL, pos, 0) + __popc(m & ((1u << lane) - 1u));
                if (ok) {
                    float norm = L.cnorm;
                    if (L.fn_field >= 0) norm = __ldg(p.ix.cache + L.fn_field * 256 + __ldg(p.ix.fnorm[L.fn_field] + d));
                    const float tf = (float)(t[j] + 1u);
                    res_doc[pos] = d;
                    res_val[pos] = L.weight * tf_factor(tf, norm);
                }
                scored_here += __popc(m);
            }
            if (p.acct && lane == 0) {
                const unsigned long long by = ((n * bd + 7) >> 3) + ((n * bt + 7) >> 3) + 16;
                if (e.y >= it.doc_lo) atomicAdd(&S.st_blocks, by); else atomicAdd(&S.st_redecode, by);
                atomicAdd(&S.st_scored, (unsigned long long)scored_here);
            }
        }
    }
    __syncthreads();
    if (tid < (int)q.n_insert && S.resident[tid]) { S.cur[tid] = S.leaf[tid].n_blocks; S.cur_next[tid] = S.leaf[tid].n_blocks; }
    {
        float4* a4 = reinterpret_cast<float4*>(acc);
        for (int i = tid; i < (DENSE ? DW : HS) / 4; i += NT) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!DENSE) {
            uint4* k4 = reinterpret_cast<uint4*>(keys);
            for (int i = tid; i < HS / 4; i += NT) k4[i] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
        }
        if (!PURE)
            for (int i = tid; i < (DENSE ? DW : HS) / 4; i += NT) reinterpret_cast<uint32_t*>(msk)[i] = 0;
    }
    __syncthreads();

#ifdef FG_PROFILE_PHASES  // dev tool: per-phase cycle counters (costs registers; off in product builds)
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tp = clock64();
#define PROF(i) do { if (p.prof && tid == 0) { const long long now_ = clock64(); pt[i] += now_ - tp; tp = now_; } } while (0)
#else
#define PROF(i) do { } while (0)
#endif
    WarpTopK<KS> tk;
    tk.init();
    float theta_s = -INFINITY;  // score part of tk.theta (a candidate needs score >= theta_s)
    uint32_t lo = it.doc_lo;
    const uint32_t end = it.doc_hi;
    uint32_t my_matches = 0, my_scored = 0;
    unsigned long long my_blocks = 0, my_redecode = 0;
    PROF(0);  // item init

    // Round setup, run by warp 0 only: next round's doc range + the scan resume points of phase 0.
    // For rounds after the first it runs inside the previous round's slot scan (its global loads
    // overlap the scan), so a round needs no setup barrier of its own.
    auto round_setup = [&](uint32_t lo_) {
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
            for (uint32_t i = lane; i < S.res_n; i += 32) {  // next resident posting at or after lo
                const uint32_t d = res_doc[i];
                if (d >= lo_) fb = min(fb, d);
            }
            fb = warp_min(fb);
            chi = warp_min(chi);
            if (DENSE && (q.flags & QF_COL_INSERT)) fb = lo_;  // a column insert leaf: every doc is a candidate
            if (lane == 0) {
                uint32_t rlo = fb == EMPTY ? end : max(lo_, fb);
                const bool fin = rlo >= end;
                // dense windows start 16-aligned (lo_ is: work items are cut at multiples of 16 and a full
                // window spans DW docs): columns and fieldnorms are read 8 docs per 64-bit load
                if (DENSE) rlo &= ~15u;
                uint32_t rhi = DENSE ? (uint32_t)min((unsigned long long)rlo + DW, (unsigned long long)end) : chi;
                // a plan of column leaves only keeps nothing in the slots: one round covers the whole range
                if (DENSE && PURE && nl == 0 && !p.ix.alive && !p.match_bitmap) rhi = end;
                S.done = fin;
                if (rhi < rlo) rhi = rlo;
                S.rlo = rlo;
                S.rhi = rhi;
                const uint32_t span = rhi - rlo;
                int sh = span > 1 ? (32 - __clz(span - 1)) - CB_LOG2 : 0;
                S.shift = sh > 0 ? sh : 0;
                S.gtheta = p.qtheta ? __ldcg(p.qtheta + it.query) : 0u;
            }
            // scan resume points of phase 0 (insert leaves), see the phase loop
            {
                const int pl1 = nlp ? (int)S.phase_end[0] : 0;
                for (int i = lane; i < pl1 * NW; i += 32) {
                    const int l = i / NW, w = i % NW;
                    const uint32_t c0 = (uint32_t)((w - l) & (NW - 1));
                    const uint32_t maxb = S.leaf[l].role != ROLE_INSERT ? 0xFFFFFFFFu
                                          : (DENSE ? (uint32_t)(DW / BLOCK + 2) : S.quota[l]);
                    S.resume[l * NW + w] = (c0 * 32u < maxb && S.cur[l] + c0 * 32u < S.leaf[l].n_blocks)
                                               ? S.cur[l] + c0 * 32u : LEAF_DONE;
                }
                if (lane == 0) S.leafmask = 0;
            }
    };
    if (warp == 0) round_setup(lo);

    while (true) {
        __syncthreads();
        if (S.done) break;
        const uint32_t rlo = S.rlo, rhi = S.rhi, shift = S.shift;
        PROF(1);  // round setup
        if (S.res_n) {  // resident postings of this round (uniform branch)
            const uint32_t rbit = S.leaf[0].bit;
            for (uint32_t i = tid; i < S.res_n; i += NT) {
                const uint32_t d = res_doc[i];
                if (d >= rlo && d < rhi) {
                    const int sl = DENSE ? (int)(d - rlo) : hash_insert(keys, d);
                    smem_add_f32(&acc[sl], res_val[i]);
                    if (!PURE && rbit) msk[sl] = (uint8_t)(msk[sl] | rbit);
                }
            }
            __syncthreads();
        }
#ifdef FG_PROFILE_PHASES
        if (p.prof && tid == 0) pt[7]++;
#endif

        // ---- streamed leaves: warp w walks the leaves nlp + w, nlp + w + NW, ... ----
        if (DENSE && PURE && q.n_stream) {
            for (int l = nlp + (p.deterministic ? 0 : warp); l < nl; l += (p.deterministic ? 1 : NW)) {
                if (p.deterministic) {  // bit-reproducible sums: one leaf at a time, in leaf order
                    __syncthreads();
                    if (warp != (l - nlp) % NW) continue;
                }
                if (S.resident[l]) continue;
                const DevLeaf& L = S.leaf[l];
                uint32_t b = S.cur[l];
                while (b < L.n_blocks) {
                    const uint4 e = __ldg(&skip[L.blk_begin + b]);
                    if (e.y >= rhi) break;
                    if (e.x >= rlo) {
                        const uint32_t c = stream_block(p, L, e, rlo, rhi, acc, lane);
                        if (p.acct) {
                            my_scored += c;
                            if (lane == 0) {
                                const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
                                const unsigned long long by = ((n * bd + 7) >> 3) + ((n * bt + 7) >> 3) + 16;
                                if (e.y >= lo) my_blocks += by; else my_redecode += by;
                            }
                        }
                    }
                    if (e.x >= rhi) break;  // continues into the next window
                    b++;
                }
                if (lane == 0) { S.cur[l] = b; S.cur_next[l] = b; }
            }
            if (nlp == 0 || p.deterministic) __syncthreads();
        }

        // ---- clause phases: leaves [l0, l1) share (role, bit) ----
        int l0 = 0;
        while (l0 < nlp) {
            const int l1 = (int)S.phase_end[l0];
            const uint32_t role = S.leaf[l0].role, bit = S.leaf[l0].bit, req = S.leaf[l0].req;
            // LF_NOFILT: a filter-role leaf without usable precondition bits (an earlier clause has a
            // column leaf): it is applied to every doc of the dense window like an insert leaf
            const bool filter = role != ROLE_INSERT && !(S.leaf[l0].lflags & LF_NOFILT);
            bool prepared = l0 == 0;  // phase 0's first pass was prepared by round_setup (no barrier needed)
            if (!prepared) {
                for (int i = tid; i < (l1 - l0) * NW; i += NT) {
                    const int l = l0 + i / NW, w = i % NW;
                    const uint32_t c0 = (uint32_t)((w - l) & (NW - 1));  // this warp's first chunk of leaf l
                    // insert leaves touch a bounded number of blocks per round: warps beyond it sit out
                    const uint32_t maxb = filter ? 0xFFFFFFFFu : (DENSE ? (uint32_t)(DW / BLOCK + 2) : S.quota[l]);
                    S.resume[l * NW + w] = (c0 * 32u < maxb && S.cur[l] + c0 * 32u < S.leaf[l].n_blocks)
                                               ? S.cur[l] + c0 * 32u : LEAF_DONE;
                }
            }
            while (true) {
                if (!prepared) {
                    if (tid == 0) S.leafmask = 0;
                    __syncthreads();
                }
                prepared = false;
                int my_pending = 0;
                // (a) skip-entry scan: every warp takes 32-entry chunks warp, warp+NW, ... of each leaf
                //     and appends the needed blocks to its private worklist segment
                uint32_t cnt = 0;
                for (int l = l0; l < l1; l++) {
                    const DevLeaf& L = S.leaf[l];
                    uint32_t b = S.resume[l * NW + warp];
                    if (b == LEAF_DONE) continue;
                    while (true) {
                        const uint32_t bi = b + lane;
                        bool in_range = false, needed = false, consumed = false;
                        if (bi < L.n_blocks) {
                            const uint4 e = __ldg(&skip[L.blk_begin + bi]);
                            in_range = e.y < rhi;
                            consumed = in_range && e.x < rhi;
                            if (in_range && e.x >= rlo) {
                                needed = true;
                                if (filter) {
                                    const uint32_t a = max(e.y, rlo) - rlo, z = min(e.x, rhi - 1) - rlo;
                                    needed = cb_any(cb, a >> shift, z >> shift);
                                    if (needed && p.exact_filter && !DENSE) {
                                        bool any = false;  // exact accounting: a candidate inside [e.y, e.x]?
                                        for (int i = 0; i < HS && !any; i++) {
                                            const uint32_t kd = keys[i];
                                            any = kd != EMPTY && kd >= e.y && kd <= e.x && (msk[i] & req) == req;
                                        }
                                        needed = any;
                                    }
                                }
                            }
                        }
                        const unsigned nm = __ballot_sync(FULL, needed);
                        const int nin = __popc(__ballot_sync(FULL, in_range));
                        const int ncons = __popc(__ballot_sync(FULL, consumed));
                        if (cnt + __popc(nm) > SEG_CAP) {  // segment full: decode what we have, come back
                            if (lane == 0) S.resume[l * NW + warp] = b;
                            my_pending = 1;
                            break;
                        }
                        if (needed) wl[warp * SEG_CAP + cnt + __popc(nm & lt_mask)] = ((uint32_t)l << 24) | bi;
                        if (lane == 0 && nm) atomicOr(&S.leafmask, 1u << l);
                        cnt += __popc(nm);
                        if (lane == 0 && ncons) atomicMax(&S.cur_next[l], b + ncons);
                        if (nin < 32) { if (lane == 0) S.resume[l * NW + warp] = LEAF_DONE; break; }
                        b += 32 * NW;
                    }
                }
                if (lane == 0) S.segcnt[warp] = cnt;
                __syncthreads();
                PROF(2);  // skip scan
                // only one leaf has blocks in this pass (typical: the other leaves of the clause are short or
                // resident): a slot is then touched by one thread only -> plain adds instead of CAS-loop atomics
                const bool solo = __popc(S.leafmask) <= 1 && !q.n_stream;  // streamed leaves add concurrently
                // (b) decode GRP blocks per warp step
                for (int sg = 0; sg < NW; sg++) {
                const uint32_t total = S.segcnt[sg];
                const uint32_t* seg = wl + sg * SEG_CAP;
                for (uint32_t i0 = (uint32_t)((warp - sg) & (NW - 1)) * GRP; i0 < total; i0 += NW * GRP) {
                    uint4 e[GRP];
                    uint32_t lf[GRP];
                    bool val[GRP];
#pragma unroll
                    for (int g = 0; g < GRP; g++) {
                        const uint32_t idx = i0 + g;
                        val[g] = idx < total;
                        const uint32_t ent = val[g] ? seg[idx] : 0u;
                        lf[g] = ent >> 24;
                        e[g] = val[g] ? __ldg(&skip[S.leaf[lf[g]].blk_begin + (ent & 0xFFFFFFu)])
                                      : make_uint4(0, 0, 0, 0);
                    }
                    // ---- fast path (the bulk of dense-mode work): one full block that lies entirely inside
                    // the window, both widths in 1..8, a TEXT leaf of a pure union. Straight-line code, no
                    // per-posting predicates, no masks; plain adds when the phase has a single leaf.
                    if (GRP == 1 && DENSE && PURE && !p.acct && val[0]) {
                        const uint4 e0 = e[0];
                        const uint32_t bd = e0.w & 63u, bt = (e0.w >> 6) & 63u;
                        const DevLeaf& L0 = S.leaf[lf[0]];
                        if (((e0.w >> 12) & 127u) == 127u && e0.y >= rlo && e0.x < rhi && bd - 1u < 8u && bt < 9u &&
                            L0.fn_field >= 0) {
                            const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e0.z * 16u);
                            const uint32_t bit0 = (uint32_t)lane * 4u * bd, bitt = (uint32_t)lane * 4u * bt;
                            const uint32_t* wt = wd + 4 * bd;
                            const uint32_t a0 = __ldg(wd + (bit0 >> 5)), a1 = __ldg(wd + (bit0 >> 5) + 1);
                            uint32_t t0 = 0, t1 = 0;
                            if (bt) { t0 = __ldg(wt + (bitt >> 5)); t1 = __ldg(wt + (bitt >> 5) + 1); }
                            const uint32_t xd = __funnelshift_r(a0, a1, bit0 & 31), md = (1u << bd) - 1u;
                            const uint32_t xt = __funnelshift_r(t0, t1, bitt & 31), mt_ = (1u << bt) - 1u;
                            const uint32_t g0 = xd & md, g1 = g0 + ((xd >> bd) & md), g2 = g1 + ((xd >> (2 * bd)) & md),
                                           g3 = g2 + ((xd >> (3 * bd)) & md);
                            const uint32_t base = warp_excl_scan(g3, lane) + e0.y + 4u * lane - rlo;  // slot of posting 0 minus its gap sum
                            const uint8_t* fnp = p.ix.fnorm[L0.fn_field] + rlo;
                            const float* cache = p.ix.cache + L0.fn_field * 256;
                            const float wgt = L0.weight;
                            const uint32_t s0 = base + g0, s1 = base + g1 + 1u, s2 = base + g2 + 2u, s3 = base + g3 + 3u;
                            const float n0 = __ldg(cache + __ldg(fnp + s0)), n1 = __ldg(cache + __ldg(fnp + s1)),
                                        n2 = __ldg(cache + __ldg(fnp + s2)), n3 = __ldg(cache + __ldg(fnp + s3));
                            const float f0 = (float)((xt & mt_) + 1u), f1 = (float)(((xt >> bt) & mt_) + 1u),
                                        f2 = (float)(((xt >> (2 * bt)) & mt_) + 1u), f3 = (float)(((xt >> (3 * bt)) & mt_) + 1u);
                            const float v0 = wgt * tf_factor(f0, n0), v1 = wgt * tf_factor(f1, n1),
                                        v2 = wgt * tf_factor(f2, n2), v3 = wgt * tf_factor(f3, n3);
                            if (solo) { acc[s0] += v0; acc[s1] += v1; acc[s2] += v2; acc[s3] += v3; }
                            else { smem_add_f32(&acc[s0], v0); smem_add_f32(&acc[s1], v1); smem_add_f32(&acc[s2], v2); smem_add_f32(&acc[s3], v3); }
                            continue;
                        }
                    }
                    // ---- filter fast path (non-lead Must / Should-under-Must / MustNot blocks): only doc ids
                    // are unpacked for the whole block; a posting survives only if it hits a candidate slot
                    // (mask test / candidate bitmap + bucket lookup), and only survivors fetch their tf,
                    // fieldnorm and score. In an intersection ~1 posting in 128 survives.
                    if (GRP == 1 && !PURE && filter && val[0]) {
                        const uint4 e0 = e[0];
                        const uint32_t bd = e0.w & 63u, bt = (e0.w >> 6) & 63u, n = ((e0.w >> 12) & 127u) + 1u;
                        const DevLeaf& L0 = S.leaf[lf[0]];
                        const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e0.z * 16u);
                        const uint32_t* wt = wd + 4 * bd;
                        uint32_t gg[4];
                        unpack4(wd, lane, bd, gg);
                        gg[1] += gg[0]; gg[2] += gg[1]; gg[3] += gg[2];
                        const uint32_t off = warp_excl_scan(gg[3], lane) + e0.y + 4u * lane;
                        if (p.acct && lane == 0) {
                            const unsigned long long by = ((n * bd + 7) >> 3) + ((n * bt + 7) >> 3) + 16;
                            if (e0.y >= lo) my_blocks += by; else my_redecode += by;
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t d = off + gg[j] + j, idx = 4u * lane + j;
                            int sl = -1;
                            if (idx < n && d >= rlo && d < rhi) {
                                if (DENSE) {
                                    sl = (int)(d - rlo);
                                    const uint32_t m = msk[sl];
                                    if ((m & req) != req || (role == ROLE_NOT && m == 0)) sl = -1;
                                } else if ((cb[((d - rlo) >> shift) >> 5] >> (((d - rlo) >> shift) & 31)) & 1u) {
                                    sl = hash_find(keys, d);
                                    if (sl >= 0 && (msk[sl] & req) != req) sl = -1;
                                }
                            }
                            if (sl >= 0) {
                                if (role == ROLE_NOT) {
                                    msk[sl] = (uint8_t)(msk[sl] | BIT_NOT);
                                } else {
                                    uint32_t tfv = 0;
                                    if (bt) {
                                        const uint32_t bp = idx * bt;
                                        const uint32_t w0 = __ldg(wt + (bp >> 5)), w1 = __ldg(wt + (bp >> 5) + 1);
                                        tfv = __funnelshift_r(w0, w1, bp & 31) & (bt >= 32 ? 0xFFFFFFFFu : ((1u << bt) - 1u));
                                    }
                                    float norm = L0.cnorm;
                                    if (L0.fn_field >= 0)
                                        norm = __ldg(p.ix.cache + L0.fn_field * 256 + __ldg(p.ix.fnorm[L0.fn_field] + d));
                                    const float t = (float)(tfv + 1u);
                                    const float v = L0.weight * tf_factor(t, norm);
                                    if (solo) acc[sl] += v; else smem_add_f32(&acc[sl], v);
                                    if (bit) msk[sl] = (uint8_t)(msk[sl] | bit);
                                    if (p.acct) my_scored++;
                                }
                            }
                        }
                        continue;
                    }
                    uint32_t gp[GRP][4], tf[GRP][4], nn[GRP];
#pragma unroll
                    for (int g = 0; g < GRP; g++) {
                        const uint32_t bd = e[g].w & 63u, bt = (e[g].w >> 6) & 63u;
                        nn[g] = val[g] ? ((e[g].w >> 12) & 127u) + 1u : 0u;
                        const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e[g].z * 16u);
                        unpack4(wd, lane, bd, gp[g]);
                        unpack4(wd + 4 * bd, lane, bt, tf[g]);
                        if (p.acct && lane == 0 && val[g]) {
                            const unsigned long long by = ((nn[g] * bd + 7) >> 3) + ((nn[g] * bt + 7) >> 3) + 16;
                            if (e[g].y >= lo) my_blocks += by; else my_redecode += by;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < GRP; g++) {
                        gp[g][1] += gp[g][0]; gp[g][2] += gp[g][1]; gp[g][3] += gp[g][2];
                        const uint32_t off = warp_excl_scan(gp[g][3], lane) + e[g].y + 4u * lane;
#pragma unroll
                        for (int j = 0; j < 4; j++) gp[g][j] += off + j;  // now doc ids
                    }
                    // slot lookup / insertion first (filter clauses touch only existing candidates) ...
                    int slot[GRP][4];
                    bool inside[GRP];  // block entirely inside the round: no per-posting range test
#pragma unroll
                    for (int g = 0; g < GRP; g++) inside[g] = e[g].y >= rlo && e[g].x < rhi;
#pragma unroll
                    for (int g = 0; g < GRP; g++) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t d = gp[g][j];
                            int sl = -1;
                            if (4u * lane + j < nn[g] && (inside[g] || (d >= rlo && d < rhi))) {
                                if (DENSE) {
                                    sl = (int)(d - rlo);
                                    if (!PURE && filter) {
                                        const uint32_t m = msk[sl];
                                        if ((m & req) != req || (role == ROLE_NOT && m == 0)) sl = -1;
                                    }
                                } else if (!filter || ((cb[((d - rlo) >> shift) >> 5] >> (((d - rlo) >> shift) & 31)) & 1u)) {
                                    sl = filter ? hash_find(keys, d) : hash_insert(keys, d);
                                    if (!PURE && sl >= 0 && filter && (msk[sl] & req) != req) sl = -1;
                                }
                            }
                            slot[g][j] = sl;
                        }
                    }
                    if (!PURE && role == ROLE_NOT) {
#pragma unroll
                        for (int g = 0; g < GRP; g++)
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (slot[g][j] >= 0) msk[slot[g][j]] = (uint8_t)(msk[slot[g][j]] | BIT_NOT);
                        continue;
                    }
                    // ... then the fieldnorm gathers of the surviving postings, issued together ...
                    float norm[GRP][4];
#pragma unroll
                    for (int g = 0; g < GRP; g++) {
                        const DevLeaf& L = S.leaf[lf[g]];
                        const int ff = L.fn_field;
                        const uint8_t* fnp = p.ix.fnorm[ff < 0 ? 0 : ff];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            norm[g][j] = L.cnorm;
                            if (slot[g][j] >= 0 && ff >= 0)
                                norm[g][j] = __ldg(p.ix.cache + ff * 256 + __ldg(fnp + gp[g][j]));
                        }
                    }
                    // ... then BM25 and the slot update
#pragma unroll
                    for (int g = 0; g < GRP; g++) {
                        const float wgt = S.leaf[lf[g]].weight;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int sl = slot[g][j];
                            if (sl < 0) continue;
                            const float t = (float)(tf[g][j] + 1u);
                            const float v = wgt * tf_factor(t, norm[g][j]);
                            if (solo) acc[sl] += v; else smem_add_f32(&acc[sl], v);
                            if (!PURE && bit) msk[sl] = (uint8_t)(msk[sl] | bit);
                            if (p.acct) my_scored++;
                        }
                    }
                }
                }
                const int more_ = __syncthreads_or(my_pending);
                PROF(3);  // decode + apply
                if (!more_) break;
            }
            // candidate bitmap for the following filter clause
            const uint32_t need = S.leaf[l1 - 1].build_cb;
            if (!PURE && need) {
                if (DENSE) {
                    if (tid < DW / 32) {
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
                    for (int i = tid; i < CBW; i += NT) cb[i] = 0;
                    __syncthreads();
                    for (int i = tid; i < HS; i += NT) {
                        const uint32_t kd = keys[i];
                        if (kd != EMPTY && (msk[i] & need) == need) {
                            const uint32_t bitpos = (kd - rlo) >> shift;
                            atomicOr(&cb[bitpos >> 5], 1u << (bitpos & 31));
                        }
                    }
                }
                __syncthreads();
            }
            l0 = l1;
        }

        // warp 0 advances the leaf cursors and prepares the next round while the others already scan
        if (warp == 0) {
            if (lane < nl) S.cur[lane] = max(S.cur[lane], S.cur_next[lane]);
            __syncwarp();
            round_setup(rhi);
        }
        // ---- slot scan: match test, cheap f32 pre-test against the k-th score, reset ----
        {
            // score threshold shared by all work items of the query: >= k docs are known to score at
            // least this much, so anything strictly below can never reach the final top-k
            {
                const uint32_t gt = max(S.gtheta, S.theta_cta);  // other items of the query / other warps of this CTA
                if (gt) theta_s = fmaxf(theta_s, unsortable(gt));
            }
            const int n4 = DENSE ? (int)((rhi - rlo + 3) >> 2) : HS / 4;
            uint32_t* m32 = reinterpret_cast<uint32_t*>(msk);
            float4* a4 = reinterpret_cast<float4*>(acc);
            uint4* k4 = reinterpret_cast<uint4*>(keys);
            const bool generic = !PURE || p.ix.alive != nullptr || p.match_bitmap != nullptr;
            if (DENSE && PURE && !generic) {
                // tight loop for the bulk case (pure union, dense window, no deletes): a touched slot has
                // a positive score and matches; only slots that can enter the top-k take the slow path
                const bool count = p.want_counts != 0;
                const bool use_acc = nl > 0;  // without block leaves the slots stay zero: scores come from columns
                if (ncol && S.col_field >= 0) {
                    // column windows: 8 docs per thread and step. The fieldnorm ids and every column's tf
                    // bytes arrive as one 64-bit load each (windows start 16-aligned), the norms come
                    // from the 1 KB table in shared memory, a term costs PRMT, 2 FADD, MUFU.RCP, FMUL, FFMA.
                    uint32_t magic;
                    FG_MAGIC_2P23(magic);
                    const int n8 = (int)((rhi - rlo + 7) >> 3);
                    const DevLeaf* CL = &S.leaf[nl];
                    const uint8_t* fnb = p.ix.fnorm[S.col_field] + rlo;
                    uint32_t seen_t = 0;
                    for (int g0 = warp * 32; g0 < n8; g0 += NT) {
                        const int g = g0 + lane;
                        {   // threshold reached by any warp of the CTA
                            const uint32_t tc = S.theta_cta;
                            if (tc != seen_t) { seen_t = tc; theta_s = fmaxf(theta_s, unsortable(tc)); }
                        }
                        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        uint32_t px = 0, py = 0;  // OR of the columns' tf bytes: non-zero byte = doc matches
                        if (g < n8) {
                            const uint32_t o8 = 8u * (uint32_t)g;
                            const uint2 fn8 = __ldg(reinterpret_cast<const uint2*>(fnb + o8));
                            uint2 ta = __ldg(reinterpret_cast<const uint2*>(CL[0].col + rlo + o8)), tb = make_uint2(0u, 0u);
                            if (ncol > 1) tb = __ldg(reinterpret_cast<const uint2*>(CL[1].col + rlo + o8));
                            if (use_acc) {
                                const float4 a0 = a4[2 * g], a1 = a4[2 * g + 1];
                                a4[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
                                a4[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                                v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
                                v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                            }
                            float n[8];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                n[j] = S.ctab[(fn8.x >> (8 * j)) & 255u];
                                n[4 + j] = S.ctab[(fn8.y >> (8 * j)) & 255u];
                            }
                            int c = 0;
                            while (true) {
                                col_term8(ta, CL[c].weight, magic, n, v);
                                px |= ta.x; py |= ta.y;
                                if (c + 1 >= ncol) break;
                                col_term8(tb, CL[c + 1].weight, magic, n, v);
                                px |= tb.x; py |= tb.y;
                                c += 2;
                                if (c >= ncol) break;
                                ta = __ldg(reinterpret_cast<const uint2*>(CL[c].col + rlo + o8));
                                if (c + 1 < ncol) tb = __ldg(reinterpret_cast<const uint2*>(CL[c + 1].col + rlo + o8));
                            }
                        }
                        if (count) {
                            if (use_acc) {
#pragma unroll
                                for (int j = 0; j < 8; j++) my_matches += v[j] > 0.f;
                            } else {
                                my_matches += nonzero_bytes(px) + nonzero_bytes(py);
                            }
                        }
                        const float mx = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                        if (__any_sync(FULL, mx > 0.f && mx + q.const_score >= theta_s)) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float sc = v[j] + q.const_score;
                                const bool c = v[j] > 0.f && sc >= theta_s;
                                tk.offer(c, c ? make_key(sc, rlo + 8 * g + j) : 0ull, k, lane);
                            }
                            const float mine = tk.theta ? unsortable((uint32_t)(tk.theta >> 32)) : -INFINITY;
                            if (mine > theta_s) {
                                theta_s = mine;
                                if (lane == 0) {
                                    atomicMax(&S.theta_cta, sortable(mine));
                                    if (p.qtheta) atomicMax(p.qtheta + it.query, sortable(mine));
                                }
                            } else if (p.qtheta) {
                                // a long round: pick up what the other work items of the query have reached
                                const uint32_t gq = __ldcg(p.qtheta + it.query);
                                if (gq > seen_t) { if (lane == 0) atomicMax(&S.theta_cta, gq); theta_s = fmaxf(theta_s, unsortable(gq)); }
                            }
                        }
                    }
                } else
                for (int g0 = warp * 32; g0 < n4; g0 += NT) {
                    const int g = g0 + lane;
                    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g < n4) {
                        if (use_acc) { av = a4[g]; a4[g] = make_float4(0.f, 0.f, 0.f, 0.f); }
                        if (ncol) {
                            float v[4] = {av.x, av.y, av.z, av.w};
                            col_add_dense(S, p, nl, ncol, rlo + 4u * (uint32_t)g, v);
                            av = make_float4(v[0], v[1], v[2], v[3]);
                        }
                    }
                    if (count) my_matches += (av.x > 0.f) + (av.y > 0.f) + (av.z > 0.f) + (av.w > 0.f);
                    const float mx = fmaxf(fmaxf(av.x, av.y), fmaxf(av.z, av.w));
                    if (__any_sync(FULL, mx > 0.f && mx + q.const_score >= theta_s)) {
                        const float raw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float sc = raw[j] + q.const_score;
                            const bool c = raw[j] > 0.f && sc >= theta_s;
                            tk.offer(c, c ? make_key(sc, rlo + 4 * g + j) : 0ull, k, lane);
                        }
                        const float mine = tk.theta ? unsortable((uint32_t)(tk.theta >> 32)) : -INFINITY;
                        if (mine > theta_s) {
                            theta_s = mine;
                            if (lane == 0) {
                                atomicMax(&S.theta_cta, sortable(mine));
                                if (p.qtheta) atomicMax(p.qtheta + it.query, sortable(mine));
                            }
                        }
                    }
                }
            } else
            for (int g0 = warp * 32; g0 < n4; g0 += NT) {
                const int g = g0 + lane;
                float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                uint4 kv = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
                uint32_t m4 = 0;
                if (g < n4) {
                    av = a4[g];
                    a4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!DENSE) { kv = k4[g]; k4[g] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY); }
                    if (!PURE) { m4 = m32[g]; if (m4) m32[g] = 0; }
                }
                float raw[4] = {av.x, av.y, av.z, av.w};
                const uint32_t kk[4] = {kv.x, kv.y, kv.z, kv.w};
                uint32_t mm[4] = {m4 & 0xFFu, (m4 >> 8) & 0xFFu, (m4 >> 16) & 0xFFu, m4 >> 24};
                if (ncol && g < n4) {
                    // column leaves: only slots that still can match (col_req = the Must bits block phases decide)
                    bool ok[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        ok[j] = (DENSE || kk[j] != EMPTY) && (PURE || (mm[j] & q.col_req) == q.col_req);
                    col_apply<DENSE>(S, p, nl, ncol, rlo + 4u * (uint32_t)g, kk, ok, raw, mm);
                }
                bool mt[4];
                bool hot;
                if (!generic) {
                    // pure union, no deletes: every touched slot matches and its score is > 0
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        mt[j] = DENSE ? raw[j] > 0.f : kk[j] != EMPTY;
                        my_matches += mt[j];
                    }
                    const float mx = fmaxf(fmaxf(raw[0], raw[1]), fmaxf(raw[2], raw[3]));
                    hot = mx + q.const_score >= theta_s && (DENSE ? mx > 0.f : true);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (PURE) mt[j] = DENSE ? raw[j] > 0.f : kk[j] != EMPTY;
                        else {
                            const uint32_t m = mm[j];
                            mt[j] = (m & 0x7Fu) == q.all_must && !(m & BIT_NOT);
                        }
                    }
                    if (p.ix.alive || p.match_bitmap) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (!mt[j]) continue;
                            const uint32_t doc = DENSE ? rlo + 4 * g + j : kk[j];
                            if (p.ix.alive && !((__ldg(p.ix.alive + (doc >> 5)) >> (doc & 31)) & 1u)) { mt[j] = false; continue; }
                            if (p.match_bitmap)
                                atomicOr(p.match_bitmap + (size_t)it.query * p.bitmap_words + (doc >> 5), 1u << (doc & 31));
                        }
                    }
                    hot = false;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        my_matches += mt[j];
                        hot |= mt[j] && raw[j] + q.const_score >= theta_s;
                    }
                }
                if (__any_sync(FULL, hot)) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t doc = DENSE ? rlo + 4 * g + j : kk[j];
                        const float sc = raw[j] + q.const_score;
                        const bool c = mt[j] && sc >= theta_s;
                        tk.offer(c, c ? make_key(sc, doc) : 0ull, k, lane);
                    }
                    const float mine = tk.theta ? unsortable((uint32_t)(tk.theta >> 32)) : -INFINITY;
                    if (mine > theta_s) {
                        theta_s = mine;
                        if (lane == 0) {
                            atomicMax(&S.theta_cta, sortable(mine));
                            if (p.qtheta) atomicMax(p.qtheta + it.query, sortable(mine));
                        }
                    }
                }
            }
        }

        lo = rhi;  // (the barrier that ends the round is the one at the top of the loop)
        PROF(4);  // slot scan
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
#ifdef FG_PROFILE_PHASES
            if (p.prof) {
                PROF(5);  // epilogue
                for (int i = 0; i < 8; i++) atomicAdd(p.prof + i, (unsigned long long)pt[i]);
            }
#endif
            if (p.stats) {
                atomicAdd(p.stats + 0, S.st_blocks);
                atomicAdd(p.stats + 1, S.st_redecode);
                atomicAdd(p.stats + 2, S.st_scored);
            }
        }
    }
}

template <int KS, int GRP, int MINB, bool DENSE, bool PURE>
__global__ void __launch_bounds__(NT, MINB) search_kernel(const SearchParams p) {
    FG_DYN_SMEM(smem);
    __shared__ Shared S;
    // [0, 32K)   dense: acc[DW]            | hash: acc[HS] + keys[HS]
    // masked plans only (pure unions need neither masks nor candidate bitmaps):
    // [32K, 40K) dense: msk[DW]            | hash: msk[HS] + candidate bitmap (CBW words)
    // [40K, 41K) dense: candidate bitmap (DW bits)
    static_assert(2 * HS <= DW && HS + CBW * 4 <= DW, "hash layout must fit the dense regions");
    float* acc = reinterpret_cast<float*>(smem);
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem) + HS;
    uint8_t* msk = smem + DW * 4;
    uint32_t* cb_dense = reinterpret_cast<uint32_t*>(smem + DW * 4 + DW);
    uint32_t* cb_hash = reinterpret_cast<uint32_t*>(msk + HS);
    uint32_t* wl = PURE ? reinterpret_cast<uint32_t*>(smem + DW * 4) : cb_dense + DW / 32;
    uint64_t* scratch = reinterpret_cast<uint64_t*>(wl + NW * SEG_CAP);

    const DevItem it = p.items[p.item_begin + blockIdx.x];
    {
        const DevQuery q0 = p.queries[it.query];
        if (threadIdx.x == 0) S.q = q0;  // the plan header lives in shared memory: it is read in every round
        if (threadIdx.x < q0.n_leaves + q0.n_col) S.leaf[threadIdx.x] = p.leaves[q0.leaf_begin + threadIdx.x];
    }
    __syncthreads();
    const DevQuery& q = S.q;
    uint32_t* res_doc = reinterpret_cast<uint32_t*>(scratch + NW * KS * 32);
    float* res_val = reinterpret_cast<float*>(res_doc + RES_CAP);
    run_item<KS, GRP, DENSE, PURE>(p, it, q, S, acc, keys, msk, DENSE ? cb_dense : cb_hash, wl, scratch, res_doc, res_val);
}


// ---------------------------------------------------------------------------------------------------
// Column-scan kernel (class 4): pure unions with at least one column insert leaf whose block leaves are
// all streamed. Every doc of the item's range is a candidate, so the item is walked in fixed windows
// of CW docs with TWO accumulator buffers: while all warps scan window r (columns + buffer r&1), the
// warps that own streamed leaves first decode those leaves' blocks of window r+1 into the other
// buffer, so the decode latency chain (skip entry -> payload -> fieldnorm -> norm) hides behind the
// scan. Scan chunks (256 docs per warp step) are claimed dynamically, so a warp that streamed takes
// fewer of them. One barrier per window, no skip scan, no worklists, no round setup.
// ---------------------------------------------------------------------------------------------------
struct CShared {
    DevQuery q;
    DevLeaf leaf[MAX_LEAVES];
    float ctab[256];
    uint32_t chunk[2];
    uint32_t theta_cta;
    uint32_t match;
    unsigned long long st_blocks, st_redecode, st_scored;
    uint32_t st_chunks, st_skipped;  // counters mode only
    uint32_t nhit[3];    // first-touch count of the hit list of window w (list w % 3), see CHIT
    uint32_t wtheta[2];  // sortable f32: the query's threshold as published before round r (slot r & 1)
};
// Hit lists: while a window's streamed postings are added to its accumulators, the slot of every doc
// touched for the first time is appended to the window's list (three lists in rotation: filled in round
// w-1, read in round w, reset in round w+1). A window whose list did not overflow and whose docs
// without a streamed posting can no longer reach the top-k is evaluated from the list alone.
constexpr int CHIT = 512;
constexpr int colscan_smem_bytes(int ks) { return 2 * CW * 4 + NW * ks * 32 * 8 + 3 * CHIT * 2; }

template <int KS, int MINB>
__global__ void __launch_bounds__(NT, MINB) colscan_kernel(const SearchParams p) {
    FG_DYN_SMEM(smem);
    __shared__ CShared S;
    float* acc = reinterpret_cast<float*>(smem);                       // [2][CW]
    uint64_t* scratch = reinterpret_cast<uint64_t*>(smem + 2 * CW * 4);  // [NW][KS][32]
    uint16_t* hitlist = reinterpret_cast<uint16_t*>(smem + 2 * CW * 4 + NW * KS * 32 * 8);  // [3][CHIT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevItem it = p.items[p.item_begin + blockIdx.x];
    {
        const DevQuery q0 = p.queries[it.query];
        if (tid == 0) {
            S.q = q0;
            S.chunk[0] = 0; S.chunk[1] = 0; S.match = 0;
            S.theta_cta = p.qtheta ? __ldcg(p.qtheta + it.query) : 0u;
            S.st_blocks = 0; S.st_redecode = 0; S.st_scored = 0;
            S.st_chunks = 0; S.st_skipped = 0;
            S.nhit[0] = 0; S.nhit[1] = 0; S.nhit[2] = 0;
            S.wtheta[0] = S.theta_cta; S.wtheta[1] = 0;
        }
        if (tid < q0.n_leaves + q0.n_col) S.leaf[tid] = p.leaves[q0.leaf_begin + tid];
    }
    __syncthreads();
    const DevQuery& q = S.q;
    const int nl = (int)q.n_leaves, ncol = (int)q.n_col, k = (int)q.k;
    const DevLeaf* CL = &S.leaf[nl];
    const int cf = CL[0].fn_field;
    S.ctab[tid] = __ldg(p.ix.cache + cf * 256 + tid);
    if (nl) {
        float4* a4 = reinterpret_cast<float4*>(acc);
        for (int i = tid; i < 2 * CW / 4; i += NT) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint4* __restrict__ skip = p.ix.skip;
    const uint32_t lo0 = it.doc_lo, end = it.doc_hi;
    // Streamed leaf of this warp (leaf index = warp; the planner keeps nl <= NW for this class). A block is
    // decoded and scored ONCE: its 128 (doc, score) pairs stay in registers (4 per lane, ascending docs)
    // and are applied to whichever window they fall into, round after round -- no re-decode of blocks
    // that straddle windows, however sparse the list.
    uint32_t cur = 0;
    uint4 e_cur = make_uint4(0u, 0u, 0u, 0u);  // skip entry of block `cur`, fetched one block ahead
    uint32_t cd[4] = {EMPTY, EMPTY, EMPTY, EMPTY};  // carried postings, EMPTY = consumed
    float cv[4] = {0.f, 0.f, 0.f, 0.f};
    // The NW warps share the plan's nl streamed leaves: warp w walks every nsub-th block of leaf w % nl
    // (sub-stream w / nl), so a plan with one or two streamed leaves still decodes on all warps.
    // (deterministic mode: one warp per leaf, leaves add in leaf order)
    const int nsub = (nl && !p.deterministic) ? NW / nl : 1;
    const int my_leaf = nl ? warp % nl : 0, my_sub = nl ? warp / nl : 0;
    const bool streams = nl && my_sub < nsub;
    if (streams) {
        const DevLeaf& L = S.leaf[my_leaf];
        uint32_t a = 0, b = L.n_blocks;  // first block whose last_doc >= doc_lo
        while (a < b) {
            const uint32_t m = (a + b) >> 1;
            if (__ldg(&skip[L.blk_begin + m]).x >= lo0) b = m; else a = m + 1;
        }
        cur = a + (uint32_t)((my_sub + nsub - (int)(a % (uint32_t)nsub)) % nsub);  // first block >= a of this sub-stream
        if (cur < L.n_blocks) e_cur = __ldg(&skip[L.blk_begin + cur]);
    }
    uint32_t my_matches = 0, my_scored = 0;
    unsigned long long my_blocks = 0, my_redecode = 0;
    // (see "sparse-hit gating" below) every doc has to be visited when the caller wants match counts or
    // the matched doc-id set
    const bool may_prune = nl && !p.no_prune && p.want_counts == 0 && p.match_bitmap == nullptr;
    const bool track = may_prune;
    auto stream_leaf = [&](uint32_t wlo, uint32_t whi, float* buf, uint32_t li) {
        const DevLeaf& L = S.leaf[my_leaf];
        while (true) {
            bool left = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (cd[j] < whi) {
                    if (cd[j] >= wlo) {
                        const uint32_t slot = cd[j] - wlo;
                        const float before = atomicAdd(&buf[slot], cv[j]);
                        my_scored++;
                        if (track && before == 0.f) {  // first posting of this doc in the window
                            const uint32_t hi = atomicAdd(&S.nhit[li], 1u);
                            if (hi < (uint32_t)CHIT) hitlist[li * CHIT + hi] = (uint16_t)slot;
                        }
                    }
                    cd[j] = EMPTY;
                }
                left = left || cd[j] != EMPTY;
            }
            if (__any_sync(FULL, left)) break;  // the block continues behind this window
            if (cur >= L.n_blocks) break;
            const uint4 e = e_cur;              // (loaded when the previous block was taken)
            if (e.y >= whi) break;              // the next block starts behind this window
            cur += (uint32_t)nsub;
            if (cur < L.n_blocks) e_cur = __ldg(&skip[L.blk_begin + cur]);
            if (e.x < wlo) continue;            // (only before the item's first window)
            const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
            const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e.z * 16u);
            uint32_t g[4], t[4];
            unpack4(wd, lane, bd, g);
            unpack4(wd + 4 * bd, lane, bt, t);
            g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
            const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
            const int ff = L.fn_field;
            const uint8_t* fnp = p.ix.fnorm[ff < 0 ? 0 : ff];
            float nrm[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t d = off + g[j] + j;
                const bool ok = 4u * lane + j < n && d >= lo0 && d < end;
                cd[j] = ok ? d : EMPTY;
                nrm[j] = L.cnorm;
                if (ok && ff >= 0) {
                    const uint32_t id = __ldg(fnp + d);
                    nrm[j] = ff == cf ? S.ctab[id] : __ldg(p.ix.cache + ff * 256 + id);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) cv[j] = L.weight * tf_factor((float)(t[j] + 1u), nrm[j]);
            if (p.acct && lane == 0) my_blocks += ((n * bd + 7) >> 3) + ((n * bt + 7) >> 3) + 16;
        }
    };
    auto stream_window = [&](uint32_t wlo, uint32_t whi, float* buf, uint32_t li) {
        if (p.deterministic) {  // bit-reproducible sums: one leaf at a time, in leaf order
            for (int l = 0; l < nl; l++) {
                if (warp == l) stream_leaf(wlo, whi, buf, li);
                __syncthreads();
            }
        } else if (streams) {
            stream_leaf(wlo, whi, buf, li);
        }
    };
    // without streamed leaves nothing lives in the accumulators: one window covers the whole range
    const uint32_t step = nl ? (uint32_t)CW : (end - lo0 + 16u);
    __syncthreads();
    if (nl) {
        stream_window(lo0, (uint32_t)min((unsigned long long)lo0 + step, (unsigned long long)end), acc, 0u);
        __syncthreads();
    }

    WarpTopK<KS> tk;
    tk.init();
    float theta_s = -INFINITY;
    uint32_t seen_t = 0;
    uint32_t magic;
    FG_MAGIC_2P23(magic);
    const bool count = p.want_counts != 0;
    const uint8_t* fnb = p.ix.fnorm[cf];
    const uint8_t* alive8 = reinterpret_cast<const uint8_t*>(p.ix.alive);
    // Sparse-hit gating (MaxScore on the columns): tf/(tf+norm) < 1, so a doc that no streamed leaf
    // touched scores below the sum of the column weights. Once the k-th best score known for the query
    // (theta_s: this warp's queue, the CTA's and the other work items' through qtheta) exceeds that
    // bound, only docs with a streamed posting can still enter the top-k, and a 256-doc chunk whose
    // accumulators are all zero is skipped before its fieldnorm / column bytes are even loaded. Exactly
    // the docs the `score >= theta_s` pre-test below would reject anyway: results do not change.
    // Off when the caller wants match counts or the matched doc-id set (every doc must be visited).
    // Two grains: a whole window is evaluated from its hit list (the docs streamed postings touched)
    // when the gate already held before the round and the list did not overflow; otherwise the window
    // is scanned and single 256-doc chunks without a hit are skipped.
    float col_bound = q.const_score;
    for (int c = 0; c < ncol; c++) col_bound += CL[c].weight * 1.000001f;  // rcp.approx: 1 ulp above 1 at most

    for (uint32_t wlo = lo0, r = 0; wlo < end; wlo += step, r++) {
        const uint32_t whi = (uint32_t)min((unsigned long long)wlo + step, (unsigned long long)end);
        float* buf = acc + (r & 1) * CW;
        if (nl && whi < end)
            stream_window(whi, (uint32_t)min((unsigned long long)whi + step, (unsigned long long)end), acc + ((r + 1) & 1) * CW, (r + 1) % 3u);
        const uint32_t n8 = (whi - wlo + 7) >> 3;
        // publishes the item's and the query's threshold; returns after a candidate entered this warp's queue
        auto publish_theta = [&]() {
            const float mine = tk.theta ? unsortable((uint32_t)(tk.theta >> 32)) : -INFINITY;
            if (mine > theta_s) {
                theta_s = mine;
                if (lane == 0) {
                    atomicMax(&S.theta_cta, sortable(mine));
                    if (p.qtheta) atomicMax(p.qtheta + it.query, sortable(mine));
                }
            } else if (p.qtheta) {  // pick up what the other work items of the query have reached
                const uint32_t gq = __ldcg(p.qtheta + it.query);
                if (gq > seen_t) { if (lane == 0) atomicMax(&S.theta_cta, gq); theta_s = fmaxf(theta_s, unsortable(gq)); }
            }
        };
        bool from_list = false;
        if (may_prune) {
            // the same for every thread of the CTA: the snapshot was written before the last barrier and the
            // list of this window is complete
            const uint32_t wt = S.wtheta[r & 1], nh = S.nhit[r % 3u];
            if (wt) theta_s = fmaxf(theta_s, unsortable(wt));
            from_list = wt != 0u && col_bound < unsortable(wt) && nh <= (uint32_t)CHIT;
            if (from_list) {
                const uint16_t* hl = hitlist + (r % 3u) * CHIT;
                for (uint32_t base = (uint32_t)warp * 32u; base < nh; base += (uint32_t)NT) {
                    const bool valid = base + lane < nh;
                    const uint32_t slot = valid ? hl[base + lane] : 0u;
                    const uint32_t d = wlo + slot;
                    float v = 0.f;
                    if (valid) {
                        v = buf[slot];
                        buf[slot] = 0.f;
                        const float nrm = S.ctab[__ldg(fnb + d)];
                        for (int c = 0; c < ncol; c++) v += CL[c].weight * tf_factor((float)__ldg(CL[c].col + d), nrm);
                        if (alive8 && !((__ldg(alive8 + (d >> 3)) >> (d & 7u)) & 1u)) v = 0.f;
                    }
                    const float sc = v + q.const_score;
                    const bool c = valid && v > 0.f && sc >= theta_s;
                    if (__any_sync(FULL, c)) {
                        tk.offer(c, c ? make_key(sc, d) : 0ull, k, lane);
                        publish_theta();
                    }
                }
                if (p.acct && tid == 0) { S.st_chunks += (n8 + 31) >> 5; S.st_skipped += (n8 + 31) >> 5; }
            }
        }
        // The scan loop is instantiated per (number of column leaves: 1, 2, any) x (streamed leaves yes/no):
        // column pointers and weights live in registers, no loop over columns, no accumulator traffic for
        // plans without streamed leaves. Lanes past the end of the last window are clamped, not branched.
        auto chunk_loop = [&](auto nc_tag, auto acc_tag) {
            constexpr int NC = decltype(nc_tag)::value;
            constexpr bool ACC = decltype(acc_tag)::value;
            const uint8_t* fnw = fnb + wlo;
            const uint8_t* cp0 = CL[0].col + wlo;
            const float w0 = CL[0].weight;
            const uint8_t* cp1 = NC == 1 ? cp0 : CL[1].col + wlo;
            const float w1 = NC == 1 ? 0.f : CL[1].weight;
            const bool slow = count || alive8 != nullptr || p.match_bitmap != nullptr;
            while (true) {
                uint32_t c0 = 0;
                if (lane == 0) {
                    c0 = smem_atomic_inc(&S.chunk[r & 1]);
                    // lane 0 decides for the warp whether this chunk may be gated (its theta_s; the lanes'
                    // copies can differ for a moment while another warp raises S.theta_cta) and sends the
                    // decision along with the chunk number
                    if (ACC && may_prune && col_bound < theta_s) c0 |= 0x80000000u;
                }
                c0 = __shfl_sync(FULL, c0, 0);
                const bool gate = ACC && (c0 >> 31) != 0u;
                c0 = (c0 & 0x7FFFFFFFu) * 32u;
                if (c0 >= n8) break;
                const bool active = c0 + lane < n8;
                const uint32_t g = active ? c0 + lane : n8 - 1;
                const uint32_t o8 = 8u * g;
                {   // threshold reached by any warp of the CTA
                    const uint32_t tc = S.theta_cta;
                    if (tc != seen_t) { seen_t = tc; theta_s = fmaxf(theta_s, unsortable(tc)); }
                }
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                float4* b4 = reinterpret_cast<float4*>(buf) + 2 * g;
                if (ACC) {
                    a0 = b4[0];
                    a1 = b4[1];
                    if (p.acct && lane == 0) atomicAdd(&S.st_chunks, 1u);
                    if (gate) {
                        const float am = fmaxf(fmaxf(fmaxf(a0.x, a0.y), fmaxf(a0.z, a0.w)), fmaxf(fmaxf(a1.x, a1.y), fmaxf(a1.z, a1.w)));
                        if (!__any_sync(FULL, active && am > 0.f)) {  // no streamed posting in these 256 docs (slots are already zero)
                            if (p.acct && lane == 0) atomicAdd(&S.st_skipped, 1u);
                            continue;
                        }
                    }
                }
                const uint2 fn8 = __ldg(reinterpret_cast<const uint2*>(fnw + o8));
                const uint2 ta = __ldg(reinterpret_cast<const uint2*>(cp0 + o8));
                uint2 tb = make_uint2(0u, 0u);
                if (NC != 1) tb = __ldg(reinterpret_cast<const uint2*>(cp1 + o8));
                float v[8];
                float accmax = 0.f;
                if (ACC) {
                    if (active) {  // a clamped lane reads the last group's slots too, but only their owner may clear them
                        b4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                        b4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
                    v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                    if (slow) accmax = fmaxf(fmaxf(fmaxf(a0.x, a0.y), fmaxf(a0.z, a0.w)), fmaxf(fmaxf(a1.x, a1.y), fmaxf(a1.z, a1.w)));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = 0.f;
                }
                float n[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    n[j] = S.ctab[(fn8.x >> (8 * j)) & 255u];
                    n[4 + j] = S.ctab[(fn8.y >> (8 * j)) & 255u];
                }
                col_term8(ta, w0, magic, n, v);
                if (NC != 1) col_term8(tb, w1, magic, n, v);
                uint32_t px = ta.x | tb.x, py = ta.y | tb.y;  // OR of the columns' tf bytes: non-zero byte = doc matches
                if (NC == 0) {  // three or more column leaves: in pairs, the next pair's loads in flight
                    uint2 t2 = make_uint2(0u, 0u), t3 = t2;
                    int c = 2;
                    t2 = __ldg(reinterpret_cast<const uint2*>(CL[2].col + wlo + o8));
                    if (ncol > 3) t3 = __ldg(reinterpret_cast<const uint2*>(CL[3].col + wlo + o8));
                    while (true) {
                        col_term8(t2, CL[c].weight, magic, n, v);
                        px |= t2.x; py |= t2.y;
                        if (c + 1 >= ncol) break;
                        col_term8(t3, CL[c + 1].weight, magic, n, v);
                        px |= t3.x; py |= t3.y;
                        c += 2;
                        if (c >= ncol) break;
                        t2 = __ldg(reinterpret_cast<const uint2*>(CL[c].col + wlo + o8));
                        if (c + 1 < ncol) t3 = __ldg(reinterpret_cast<const uint2*>(CL[c + 1].col + wlo + o8));
                    }
                }
                if (slow) {
                    const uint32_t d8 = wlo + o8;
                    if (alive8) {  // deleted docs: windows are 16-aligned, so docs d8..d8+7 are one byte of the bitset
                        const uint32_t al = __ldg(alive8 + (d8 >> 3));
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = ((al >> j) & 1u) ? v[j] : 0.f;
                    }
                    if (p.match_bitmap && active) {
                        uint32_t bits = 0;
#pragma unroll
                        for (int j = 0; j < 8; j++) bits |= (v[j] > 0.f ? 1u : 0u) << j;
                        if (bits) atomicOr(p.match_bitmap + (size_t)it.query * p.bitmap_words + (d8 >> 5), bits << (d8 & 31));
                    }
                    if (count && active) {
                        // matches = docs with a non-zero column byte, + docs that only streamed leaves touched (rare)
                        if (alive8 || accmax > 0.f) {
#pragma unroll
                            for (int j = 0; j < 8; j++) my_matches += v[j] > 0.f;
                        } else {
                            my_matches += nonzero_bytes(px) + nonzero_bytes(py);
                        }
                    }
                }
                const float mx = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
                if (__any_sync(FULL, active && mx > 0.f && mx + q.const_score >= theta_s)) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float sc = v[j] + q.const_score;
                        const bool c = active && v[j] > 0.f && sc >= theta_s;
                        if (!__any_sync(FULL, c)) continue;
                        tk.offer(c, c ? make_key(sc, wlo + o8 + j) : 0ull, k, lane);
                    }
                    publish_theta();
                }
            }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        if (from_list) {
            // done above
        } else if (nl) {
            if (ncol == 1) chunk_loop(I1{}, std::true_type{});
            else if (ncol == 2) chunk_loop(I2{}, std::true_type{});
            else chunk_loop(I0{}, std::true_type{});
        } else {
            if (ncol == 1) chunk_loop(I1{}, std::false_type{});
            else if (ncol == 2) chunk_loop(I2{}, std::false_type{});
            else chunk_loop(I0{}, std::false_type{});
        }
        if (tid == 0) {
            S.chunk[(r + 1) & 1] = 0;
            if (may_prune) {
                S.nhit[(r + 2) % 3u] = 0;  // the list window r+2 will fill in the next round (last read in round r-1)
                uint32_t t = S.theta_cta;
                if (p.qtheta) t = max(t, __ldcg(p.qtheta + it.query));
                S.wtheta[(r + 1) & 1] = t;  // what round r+1 may rely on
            }
        }
        __syncthreads();
    }

    // ---- epilogue: merge the warp queues, write the partial list ----
    my_matches = warp_sum(my_matches);
    my_scored = warp_sum(my_scored);
    if (lane == 0) {
        atomicAdd(&S.match, my_matches);
        if (p.acct) {
            atomicAdd(&S.st_scored, (unsigned long long)my_scored);
            atomicAdd(&S.st_blocks, my_blocks);
            atomicAdd(&S.st_redecode, my_redecode);
        }
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
            const int rr = s * 32 + lane;
            if (rr < (int)p.kcap) p.partial[(size_t)it.slot * p.kcap + rr] = rr < k ? tk.q[s] : 0;
        }
        if (lane == 0) {
            p.partial_count[it.slot] = S.match;
            if (p.stats && p.acct) {
                atomicAdd(p.stats + 0, S.st_blocks);
                atomicAdd(p.stats + 1, S.st_redecode);
                atomicAdd(p.stats + 2, S.st_scored);
                atomicAdd(p.stats + 3, (unsigned long long)S.st_chunks);   // chunks of windowed plans reached ...
                atomicAdd(p.stats + 4, (unsigned long long)S.st_skipped);  // ... and skipped by the sparse-hit gate
            }
        }
    }
}

// one warp per query: merge the per-item partial lists
template <int KS>
__global__ void __launch_bounds__(128) merge_kernel(const MergeParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= p.n_queries) return;
    const DevQuery q = p.queries[qi];
    const int k = (int)q.k;
    if (q.flags & QF_ALL) {  // AllQuery: the first k alive docs (every score equals const_score)
        uint2* outp = reinterpret_cast<uint2*>(p.out_hits) + (size_t)qi * p.k_stride;
        uint32_t found = 0;
        for (uint32_t base = 0; base < p.n_docs && found < (uint32_t)k; base += 32) {
            const uint32_t d = base + lane;
            const bool al = d < p.n_docs && (!p.alive || ((p.alive[d >> 5] >> (d & 31)) & 1u));
            const unsigned m = __ballot_sync(FULL, al);
            const uint32_t r = found + __popc(m & ((1u << lane) - 1u));
            if (al && r < (uint32_t)k && r < p.k_stride) outp[r] = make_uint2(__float_as_uint(q.const_score), d + p.doc_base);
            found += __popc(m);
        }
        const uint32_t nh_all = min(min(found, (uint32_t)k), p.k_stride);
        for (uint32_t r = nh_all + lane; r < p.k_stride; r += 32) outp[r] = make_uint2(0u, 0xFFFFFFFFu);
        if (lane == 0) {
            p.out_n[qi] = nh_all;
            if (p.out_count) p.out_count[qi] = p.n_alive;
        }
        return;
    }
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

int search_smem_bytes(int ks, bool pure) {
    return DW * 4 + (pure ? 0 : DW + DW / 8) + NW * SEG_CAP * 4 + NW * ks * 32 * 8 + RES_CAP * 8;
}

#ifndef PURE_MINB
#define PURE_MINB 4
#endif
template <int KS, int GRP, int MINB, bool DENSE, bool PURE>
static void launch_one(SearchParams p, uint32_t begin, uint32_t count, cudaStream_t st) {
    if (!count) return;
    static bool configured = false;
    const int smem = search_smem_bytes(KS, PURE);
    if (!configured) {
        cudaFuncSetAttribute(search_kernel<KS, GRP, MINB, DENSE, PURE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    p.item_begin = begin;
    auto kernel = search_kernel<KS, GRP, MINB, DENSE, PURE>;
    FG_LAUNCH(kernel, count, NT, smem, st, p);
}

#ifndef COLSCAN_MINB
#define COLSCAN_MINB 4
#endif
template <int KS, int MINB>
static void launch_colscan(SearchParams p, uint32_t begin, uint32_t count, cudaStream_t st) {
    if (!count) return;
    static bool configured = false;
    const int smem = colscan_smem_bytes(KS);
    if (!configured) {
        cudaFuncSetAttribute(colscan_kernel<KS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    p.item_begin = begin;
    auto kernel = colscan_kernel<KS, MINB>;
    FG_LAUNCH(kernel, count, NT, smem, st, p);
}

// items are grouped by kernel class: [dense pure | dense masked | hash pure | hash masked | column scan]
void launch_search(const SearchParams& p, int ks, const uint32_t class_count[NCLS], void* const streams[NCLS]) {
    cudaStream_t s0 = (cudaStream_t)streams[0], s1 = (cudaStream_t)streams[1], s2 = (cudaStream_t)streams[2],
                 s3 = (cudaStream_t)streams[3], s4 = (cudaStream_t)streams[4];
    const uint32_t b1 = class_count[0], b2 = b1 + class_count[1], b3 = b2 + class_count[2], b4 = b3 + class_count[3];
    if (ks <= 1) launch_colscan<1, COLSCAN_MINB>(p, b4, class_count[4], s4);
    else if (ks <= 4) launch_colscan<4, 3>(p, b4, class_count[4], s4);
    if (ks <= 1) {
        launch_one<1, 1, PURE_MINB, true, true>(p, 0, class_count[0], s0);
        launch_one<1, 1, 4, true, false>(p, b1, class_count[1], s1);
        launch_one<1, 1, PURE_MINB, false, true>(p, b2, class_count[2], s2);
        launch_one<1, 1, 4, false, false>(p, b3, class_count[3], s3);
    } else if (ks <= 4) {
        launch_one<4, 1, 3, true, true>(p, 0, class_count[0], s0);
        launch_one<4, 1, 3, true, false>(p, b1, class_count[1], s1);
        launch_one<4, 1, 3, false, true>(p, b2, class_count[2], s2);
        launch_one<4, 1, 3, false, false>(p, b3, class_count[3], s3);
    } else {  // k up to 1024: 32 queue rows per lane (slow path: deep pagination only)
        launch_one<32, 1, 1, true, true>(p, 0, class_count[0], s0);
        launch_one<32, 1, 1, true, false>(p, b1, class_count[1], s1);
        launch_one<32, 1, 1, false, true>(p, b2, class_count[2], s2);
        launch_one<32, 1, 1, false, false>(p, b3, class_count[3], s3);
    }
}

void launch_merge(const MergeParams& p, int ks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p.n_queries == 0) return;
    const unsigned grid = (p.n_queries + 3) / 4;
    if (ks <= 1) FG_LAUNCH(merge_kernel<1>, grid, 128, 0, st, p);
    else if (ks <= 4) FG_LAUNCH(merge_kernel<4>, grid, 128, 0, st, p);
    else FG_LAUNCH(merge_kernel<32>, grid, 128, 0, st, p);
}

void launch_merge_gathered(const void* hits, const uint32_t* n, uint32_t n_ranks, uint32_t n_queries,
                           uint32_t k, uint32_t k_stride, void* out_hits, uint32_t* out_n, int ks,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_queries == 0) return;
    const unsigned grid = (n_queries + 3) / 4;
    if (ks <= 1)
        FG_LAUNCH(merge_gathered_kernel<1>, grid, 128, 0, st, (const uint2*)hits, n, n_ranks, n_queries, k, k_stride,
                  (uint2*)out_hits, out_n);
    else if (ks <= 4)
        FG_LAUNCH(merge_gathered_kernel<4>, grid, 128, 0, st, (const uint2*)hits, n, n_ranks, n_queries, k, k_stride,
                  (uint2*)out_hits, out_n);
    else
        FG_LAUNCH(merge_gathered_kernel<32>, grid, 128, 0, st, (const uint2*)hits, n, n_ranks, n_queries, k, k_stride,
                  (uint2*)out_hits, out_n);
}

}  // namespace fg
