// Lead-driven evaluation of the fugu query hot path on sm_100a: what tantivy does inside
// `searcher.search(&q, &TopDocs::with_limit(k))` (/root/reference/src/db/search.rs:162) -- posting
// decode, Intersection / Union / RequiredOptional / Exclude, Bm25Weight, TopDocs with block-max pruning
//  -- restructured for a GPU:
//
//   * every warp is autonomous: it takes work items from a global queue, walks the 128-posting blocks
//     of ONE lead leaf of ONE query, and keeps its own top-k queue in registers. No CTA barriers, no
//     shared accumulators, no atomics on scores: a document's score is summed by one lane, in leaf
//     order, so results are bit-reproducible.
//   * a lead block is decoded by the whole warp (4 postings per lane, warp prefix sum), its candidates
//     are compacted and looked up in the query's other leaves: one byte load from a dense tf column, or
//     a warp-cooperative gallop over the 16-byte skip entries (32 entries per step, ballot) followed by
//     a decode of the target block into shared memory and a binary search per candidate.
//   * MaxScore / block-max pruning (fg_internal.h): whole leads, blocks and candidates are skipped once
//     the query's k-th best score (shared by all warps of the query through one global word) exceeds
//     what they can still reach. Pruning never changes the result; it is off when the caller wants match
//     counts or the matched doc-id set.
#include <fg_ptx.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "fg_device.h"
#include "fg_internal.h"

namespace fg {

using namespace dev;

namespace {

constexpr int LNT = 256;      // threads per CTA
constexpr int LNW = LNT / 32; // warps per CTA
#ifndef FG_LEAD_LC
#define FG_LEAD_LC 4
#endif
constexpr int LC = FG_LEAD_LC;  // decoded lookup blocks cached per warp (slot = LLeaf::slot, assigned by the lowering)
static_assert(LC % 4 == 0, "WarpShared::cdocs must stay 16-byte aligned");
constexpr int ROUND = 128;            // candidates evaluated together: 4 per lane (4 independent gathers in flight per lookup)
constexpr int CAND_CAP = 2 * BLOCK;   // candidates wait here until a full round is available
constexpr int STAGE_BYTES = 512;      // staged payload capacity per slot: bd + bt <= 32 bits per posting

struct WarpShared {
    LLeaf leaf[LMAX_LEAVES];
    uint4 cure[LMAX_LEAVES];     // skip entry of block cur[j] (valid while cur[j] != EMPTY)
    uint32_t cur[LMAX_LEAVES];   // per leaf: the block the last lookup ended in (lookups ascend), EMPTY = none yet
    uint32_t ctag[LC];           // global block index of the decoded block in cdocs[slot], EMPTY = none
    uint32_t cdocs[LC][BLOCK];   // doc ids of a decoded lookup block, padded with 0xFFFFFFFF
    uint32_t cand_doc[CAND_CAP]; // candidates that survived their lead block's bound test, ascending
    uint32_t cand_val[CAND_CAP]; // their lead score (f32 bits), or the lead tf while required clauses come first
};
// Only in the kernel variant that stages lead-block payloads with 1-D bulk copies (TMA engine): two slots per warp, the
// copy of the next block to decode is in flight while the current one is processed; payloads above STAGE_BYTES are
// read with plain loads instead. (Kept out of WarpShared: shared memory the default variant does not use would only
// shrink its L1.)
struct StageShared {
    uint32_t stage[2][STAGE_BYTES / 4 + 4];
    uint64_t bar[2];
    uint32_t bar_parity;  // bit s: parity of the phase the next wait on bar[s] needs (carried from item to item)
    uint32_t pad_[3];
};
struct LeadShared {
    WarpShared w[LNW];
};
// The BM25 norm caches K1*(1-B+B*dl/avg) (1 KB per field, by fieldnorm id) are read from global memory: they live in
// L1, and 8 KB less shared memory per CTA is 24 KB more L1 per SM for the lookups' gathers (measured: -4 % step time).
#define NORM_CACHE(S, p, i) __ldg((p).ix.cache + (i))
#ifndef FG_LEAD_MINB
#define FG_LEAD_MINB 3
#endif

// per-warp byte / posting counters (FG_EXEC_COUNTERS)
struct Acct {
    unsigned long long block_bytes = 0;   // payload + 16 B skip entry of every block decoded (lead or lookup)
    unsigned long long meta_bytes = 0;    // block-max words and skip entries read while skipping
    unsigned long long gathers = 0;       // 1-byte gathers: fieldnorm ids, column bytes, alive bits
    unsigned long long gathers_l = 0, block_bytes_l = 0;  // the same, counted per lane (summed over the warp at the end)
    unsigned long long lead_blocks = 0, lead_blocks_seen = 0;
};

// dev tool (make PROFILE=1): time of a warp per phase of run_item, summed into stats[16 + phase] (ns) by lane 0
#ifdef FG_PROFILE_PHASES
#define PH_DECL unsigned long long ph_t = global_timer_ns(), ph_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PH_MARK(i) do { const unsigned long long t_ = global_timer_ns(); ph_acc[i] += t_ - ph_t; ph_t = t_; } while (0)
#define PH_FLUSH() do { if (lane == 0 && p.stats) { for (int i_ = 0; i_ < 10; i_++) atomicAdd(p.stats + 16 + i_, ph_acc[i_]); } } while (0)
#else
#define PH_DECL
#define PH_MARK(i)
#define PH_FLUSH()
#endif
enum { PH_PLAN = 0, PH_WALK = 1, PH_DECODE = 2, PH_LOOKUP = 3, PH_SCORE = 4, PH_TOPK = 5, PH_APPEND = 6, PH_OTHER = 7, PH_LOOKUP_BM = 8, PH_LOOKUP_PROBE = 9 };

// First block index in [from, n) whose last_doc >= target; n if there is none. Warp-collective
// (uniform arguments): the next 32 skip entries first (the common case while candidates and blocks
// advance together), then a 32-ary search over the rest of the list.
__device__ __forceinline__ uint32_t seek(const uint4* __restrict__ sk, uint32_t n, uint32_t from, uint32_t target,
                                         int lane, bool acct, Acct& A) {
    if (from >= n) return n;
    {
        const uint32_t i = from + lane;
        const bool ge = i >= n || __ldg(&sk[i].x) >= target;
        const unsigned m = __ballot_sync(FULL, ge);
        if (acct) A.meta_bytes += 4u * 32u;
        if (m) return min(from + (uint32_t)__ffs(m) - 1u, n);
    }
    uint32_t lo = from + 32u, hi = n;  // blocks below lo end before target; block hi (if < n) does not
    while (hi - lo > 32u) {
        const uint32_t step = (hi - lo + 31u) / 32u;
        const uint32_t i = lo + ((uint32_t)lane + 1u) * step - 1u;
        const bool ge = i >= hi || __ldg(&sk[i].x) >= target;
        const unsigned m = __ballot_sync(FULL, ge);
        if (acct) A.meta_bytes += 4u * 32u;
        if (!m) return hi;  // lane 31 probed block hi - 1 or beyond
        const uint32_t f = (uint32_t)__ffs(m) - 1u;
        const uint32_t nhi = min(lo + (f + 1u) * step - 1u, hi);
        lo = lo + f * step;
        hi = nhi;
    }
    const uint32_t i = lo + lane;
    const bool ge = i >= hi || __ldg(&sk[i].x) >= target;
    const unsigned m = __ballot_sync(FULL, ge);
    if (acct) A.meta_bytes += 4u * 32u;
    return m ? min(lo + (uint32_t)__ffs(m) - 1u, hi) : hi;
}

// tf of posting `pos` of the block described by skip entry e
__device__ __forceinline__ uint32_t extract_tf(const uint8_t* __restrict__ blk, const uint4 e, uint32_t pos) {
    const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u;
    if (bt == 0) return 1u;
    const uint32_t* wt = reinterpret_cast<const uint32_t*>(blk + (size_t)e.z * 16u) + 4u * bd;
    const uint32_t bit = pos * bt, wi = bit >> 5, sh = bit & 31u;
    const uint32_t lo = __ldg(wt + wi), hi = __ldg(wt + wi + 1);
    const uint32_t m = bt >= 32u ? 0xFFFFFFFFu : ((1u << bt) - 1u);
    return (__funnelshift_r(lo, hi, sh) & m) + 1u;
}

// doc ids of block (L, b) in the warp's lookup cache (decoded on a miss); 128 entries, ascending,
// padded with 0xFFFFFFFF
__device__ __forceinline__ const uint32_t* cached_block(const uint8_t* __restrict__ blk, WarpShared& W, int j,
                                                        uint32_t gblock, const uint4 e, int lane, bool acct, Acct& A) {
    const int slot = (int)W.leaf[j].slot & (LC - 1);
    uint32_t* cd = W.cdocs[slot];
    if (W.ctag[slot] != gblock) {  // uniform
        const uint32_t bd = e.w & 63u, n = ((e.w >> 12) & 127u) + 1u;
        const uint32_t* wd = reinterpret_cast<const uint32_t*>(blk + (size_t)e.z * 16u);
        uint32_t g[4];
        unpack4(wd, lane, bd, g);
        g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
        const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
        uint4 d;
        d.x = 4u * lane + 0u < n ? off + g[0] + 0u : 0xFFFFFFFFu;
        d.y = 4u * lane + 1u < n ? off + g[1] + 1u : 0xFFFFFFFFu;
        d.z = 4u * lane + 2u < n ? off + g[2] + 2u : 0xFFFFFFFFu;
        d.w = 4u * lane + 3u < n ? off + g[3] + 3u : 0xFFFFFFFFu;
        __syncwarp();  // earlier searches in this slot are done
        reinterpret_cast<uint4*>(cd)[lane] = d;
        if (lane == 0) W.ctag[slot] = gblock;
        __syncwarp();
        if (acct) A.block_bytes += ((n * bd + 7u) >> 3) + 16u;
    }
    return cd;
}

// Term frequency of doc c in leaf L (0 = absent) for the lanes with `live`; candidates ascend with the
// lane index, and from one call to the next (per leaf). Warp-collective. need_tf = false: presence only.
// The leaf's cursor (block index + its skip entry, in shared memory) makes the common cases cheap: the
// candidates still fall into the block the previous call ended in (sparse list: no global load at all), or
// into one of the next 32 blocks (one 4-byte load per lane + ballot).
__device__ __forceinline__ uint32_t probe(const LeadParams& p, WarpShared& W, int j, const LLeaf& L, uint32_t c,
                                          bool live, bool need_tf, int lane, Acct& A) {
    const unsigned m = __ballot_sync(FULL, live);
    if (!m) return 0u;
    const bool acct = p.acct != 0;
    const uint4* __restrict__ sk = p.ix.skip + L.blk_begin;
    const uint32_t cmin = __shfl_sync(FULL, c, __ffs(m) - 1), cmax = __shfl_sync(FULL, c, 31 - __clz((int)m));
    uint32_t b = W.cur[j];
    uint4 e = W.cure[j];
    if (b == EMPTY || e.x < cmin) {  // (uniform) the cursor's block ends before the first candidate
        b = seek(sk, L.n_blocks, b == EMPTY ? 0u : b + 1u, cmin, lane, acct, A);
        if (b < L.n_blocks) {
            e = __ldg(&sk[b]);
            if (acct) A.meta_bytes += 16u;
        }
    }
    uint32_t tf = 0u;
    bool pend = live;
    while (b < L.n_blocks) {
        if (e.y > cmax) break;  // the block starts behind the last candidate
        const bool inb = pend && c >= e.y && c <= e.x;
        if (__any_sync(FULL, inb)) {
            const uint32_t* cd = cached_block(p.ix.blk, W, j, L.blk_begin + b, e, lane, acct, A);
            // sparse lists: most of the time no posting of the block lies between the first and the last candidate
            const uint4 d4 = reinterpret_cast<const uint4*>(cd)[lane];
            const bool inr = (d4.x >= cmin && d4.x <= cmax) || (d4.y >= cmin && d4.y <= cmax) ||
                             (d4.z >= cmin && d4.z <= cmax) || (d4.w >= cmin && d4.w <= cmax);
            if (__any_sync(FULL, inr)) {
                uint32_t pos = 0u;  // entries < c among the first 127
#pragma unroll
                for (uint32_t s = 64u; s; s >>= 1)
                    if (inb && cd[pos + s - 1u] < c) pos += s;
                if (inb && cd[pos] == c) {
                    tf = need_tf ? extract_tf(p.ix.blk, e, pos) : 1u;
                    if (acct && need_tf) A.block_bytes += 8u;
                }
            }
        }
        pend = pend && c > e.x;
        const unsigned mp = __ballot_sync(FULL, pend);
        if (!mp) break;
        const uint32_t cn = __shfl_sync(FULL, c, __ffs(mp) - 1);
        b = seek(sk, L.n_blocks, b + 1u, cn, lane, acct, A);
        if (b < L.n_blocks) {
            e = __ldg(&sk[b]);
            if (acct) A.meta_bytes += 16u;
        }
    }
    __syncwarp();
    if (lane == 0) {
        W.cur[j] = b;
        W.cure[j] = e;
    }
    __syncwarp();
    return tf;
}

template <int KS, bool TMA>
__device__ __forceinline__ void run_item(const LeadParams& p, const LeadShared& S, WarpShared& W, StageShared& G, const LItem item, const uint32_t theta_seen, int lane) {
    PH_DECL;
    const LQuery q = p.queries[item.query];
    __syncwarp();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.leaves + q.leaf_begin);
        uint32_t* dst = reinterpret_cast<uint32_t*>(W.leaf);
        for (uint32_t i = lane; i < q.n_leaves * (uint32_t)(sizeof(LLeaf) / 4); i += 32) dst[i] = __ldg(src + i);
        W.cur[lane] = EMPTY;
        if (lane < LC) W.ctag[lane] = EMPTY;
    }
    __syncwarp();
    PH_MARK(PH_PLAN);
    const int k = (int)q.k;
    const bool prune = !p.exhaustive && (q.flags & LQ_PRUNE);
    const bool acct = p.acct != 0;
    const int n_lead = (int)q.n_lead, n_req = (int)q.n_req, n_leaves = (int)q.n_leaves;
    const bool has_req = n_req != 0;
    const float slack = q.slack + q.const_score;  // every bound below is compared with final scores, which include the constant
    const unsigned lt = (1u << lane) - 1u;
    Acct A;

    // KS == 0 (deep pages, k > 1024): no register queue -- every accepted hit is appended to the query's partial region
    // right away (the histogram threshold still prunes) and lead_select_kernel picks the k best afterwards
    WarpTopK<(KS > 0 ? KS : 1)> tk;
    tk.init();
    float theta = -INFINITY;  // a candidate needs score >= theta (ties are decided by the doc id in the queue)
    uint32_t pub = 0u;        // sortable threshold this warp has seen or published
    if (prune && theta_seen) {  // the value the caller read for its item-level test
        pub = theta_seen;
        theta = unsortable(theta_seen);
    }
    uint32_t n_match = 0u;
    uint32_t ncand = 0u;      // candidates waiting in W.cand_* (fewer than a round between blocks)

    // BM25 norms of the candidates selected by `mask` (bit r = row r) in leaf L's field: the four fieldnorm gathers are
    // issued together (rows outside the mask read doc 0: no branch between the loads), then the cache lookups
    auto norms4 = [&](const LLeaf& L, const uint32_t (&cc)[4], uint32_t mask, float (&out)[4]) {
        if (L.fn_field >= 0) {
            const uint8_t* __restrict__ fn = p.ix.fnorm[L.fn_field];
            uint32_t id[4];
#pragma unroll
            for (int r = 0; r < 4; r++) id[r] = __ldg(fn + (((mask >> r) & 1u) ? cc[r] : 0u));
#pragma unroll
            for (int r = 0; r < 4; r++) out[r] = NORM_CACHE(S, p, L.fn_field * 256 + (int)id[r]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) out[r] = L.cnorm;
        }
    };

    // ---- payload staging (p.tma): slot state is uniform across the warp ----
    uint32_t st_blk[2] = {EMPTY, EMPTY};  // payload offset (off16) staged or in flight in each slot
    uint32_t st_par = TMA ? G.bar_parity : 0u;  // bit s: parity of the phase the next wait on slot s needs
    uint32_t st_pend = 0u;                // bit s: a copy into slot s was issued and not waited for yet
    auto stage_wait = [&](int sl) {
        if ((st_pend >> sl) & 1u) {
            mbar_wait(&G.bar[sl], (st_par >> sl) & 1u);
            st_par ^= 1u << sl;
            st_pend &= ~(1u << sl);
        }
    };
    auto stage_issue = [&](int sl, const uint4 e2) {
        const uint32_t bytes = 16u * ((e2.w & 63u) + ((e2.w >> 6) & 63u));
        stage_wait(sl);  // (a slot is re-armed only after its previous copy has landed)
        __syncwarp();
        if (lane == 0) bulk_g2s(G.stage[sl], p.ix.blk + (size_t)e2.z * 16u, bytes, &G.bar[sl]);
        st_blk[sl] = e2.z;
        st_pend |= 1u << sl;
    };
    auto stageable = [&](const uint4 e2) -> bool {
        const uint32_t wsum = (e2.w & 63u) + ((e2.w >> 6) & 63u);
        return TMA && wsum != 0u && wsum * 16u <= (uint32_t)STAGE_BYTES;
    };
    int st_cur = 0;  // slot the next decode reads from

    // An item walks the leads [lead0, lead1) of its query one after the other (short lists share an item: the per-item
    // cost -- plan fetch, threshold, result append -- is paid once, and the threshold the early leads establish is
    // already in this warp's queue when the later ones start); a long lead has an item of its own, in several copies.
    const int lead0 = (int)(item.lead & 0xFFFFu), lead1 = (int)(item.lead >> 16);
    // blocks claimed per step: a plan with required clauses evaluates a full round of lookups per lead block (every
    // posting is a candidate), so its leads are shared out in small steps; a pruned union mostly skips blocks
    // (an unpruned union block costs a round of lookups in each of the other leaves: the step shrinks with their number)
    const uint32_t chunk = has_req ? min(p.chunk, max(2u, p.chunk_req / (uint32_t)n_req)) : min(p.chunk, max(2u, p.union_work / (uint32_t)max(1, n_leaves - 1)));
    for (int lead = lead0; lead < lead1; lead++) {
    const LLeaf LD = W.leaf[lead];
    const uint4* __restrict__ sk = p.ix.skip + LD.blk_begin;
    const float* __restrict__ bmx = p.ix.bmax + LD.blk_begin;
    const float rest_s = LD.rest + slack;
    uint32_t* const cursor = p.cursors + item.cursor + (uint32_t)(lead - lead0);
    if (lead != lead0) {  // the lookups of the next lead start from the front of every list again
        __syncwarp();
        W.cur[lane] = EMPTY;
        __syncwarp();
    }

    // ---- state of the block walk: chunk [g0 .. b1) claimed from the lead's cursor, current group of 32 blocks ----
    uint32_t g0 = 0u, b1 = 0u;
    unsigned m = 0u;     // blocks of the current group still to decode
    uint4 eg = make_uint4(0u, 0u, 0u, 0u);  // this lane's skip entry of the group
    float bound = INFINITY;                 // ... and its block's upper bound
    bool fin = false;    // nothing left to decode (list exhausted, or the lead cannot contribute a hit any more)
    {
        if (prune && lead != lead0) {  // (the first lead starts from the value read with the item)
            // (one lane's view of the shared threshold for the whole warp: two lanes may see different values)
            const uint32_t g = __shfl_sync(FULL, __ldcg(p.qtheta + item.query), 0);
            if (g > pub) { pub = g; theta = fmaxf(theta, unsortable(g)); }
        }
        fin = prune && LD.ub + rest_s < theta;
    }


    while (true) {
        // ================= next block to decode, if any =================
        bool have = false;
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        while (!fin && !have) {
            if (prune && m) m &= __ballot_sync(FULL, bound >= theta);  // the threshold may have risen
            if (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                e.x = __shfl_sync(FULL, eg.x, src);
                e.y = __shfl_sync(FULL, eg.y, src);
                e.z = __shfl_sync(FULL, eg.z, src);
                e.w = __shfl_sync(FULL, eg.w, src);
                have = true;
                if (TMA && stageable(e)) {
                    if (st_blk[st_cur] != e.z) {
                        if (st_blk[st_cur ^ 1] == e.z) st_cur ^= 1;  // staged ahead by the previous block
                        else stage_issue(st_cur, e);
                    }
                    if (m) {  // the block after this one (unless the threshold drops it): its copy runs under this block's work
                        const int nsrc = __ffs(m) - 1;
                        uint4 ne;
                        ne.x = __shfl_sync(FULL, eg.x, nsrc);
                        ne.y = __shfl_sync(FULL, eg.y, nsrc);
                        ne.z = __shfl_sync(FULL, eg.z, nsrc);
                        ne.w = __shfl_sync(FULL, eg.w, nsrc);
                        if (stageable(ne) && st_blk[st_cur ^ 1] != ne.z) stage_issue(st_cur ^ 1, ne);
                    }
                }
                break;
            }
            if (g0 >= b1) {  // claim the next chunk of the lead
                if (b1 >= LD.n_blocks && b1 != 0u) { fin = true; break; }  // (this warp took the list's last chunk)
                uint32_t c0 = 0u;
                if (lane == 0) c0 = atomicAdd(cursor, chunk);
                c0 = __shfl_sync(FULL, c0, 0);
                if (c0 >= LD.n_blocks) { fin = true; break; }
                g0 = c0;
                b1 = min(c0 + chunk, LD.n_blocks);
            }
            if (prune) {
                const uint32_t g = __shfl_sync(FULL, __ldcg(p.qtheta + item.query), 0);
                if (g > pub) { pub = g; theta = fmaxf(theta, unsortable(g)); }
                if (LD.ub + rest_s < theta) { fin = true; break; }
            }
            const uint32_t b = g0 + lane;
            bound = INFINITY;
            if (b < b1) eg = __ldg(&sk[b]);  // the group's skip entries: one coalesced request, independent of the bound load
            if (prune && b < b1) bound = LD.weight * __ldg(bmx + b) + rest_s;
            if (acct) {
                A.lead_blocks_seen += min(32u, b1 - g0);
                A.meta_bytes += (prune ? 20u : 16u) * min(32u, b1 - g0);
            }
            m = __ballot_sync(FULL, b < b1 && bound >= theta);
            if ((m >> lane) & 1u) {  // payloads of the blocks that will be decoded: on their way while earlier ones are processed
                const uint8_t* pp = p.ix.blk + (size_t)eg.z * 16u;
                prefetch_l2(pp);
                if (((eg.w & 63u) + ((eg.w >> 6) & 63u)) > 8u) prefetch_l2(pp + 128);
            }
            g0 += 32u;
        }

        PH_MARK(PH_WALK);
        // ================= decode the lead block, keep the postings that can still reach the top-k =================
        if (have) {
            const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
            uint32_t g[4], t[4];
            if (TMA && stageable(e)) {  // (uniform) the payload was staged in shared memory by a bulk copy
                stage_wait(st_cur);
                const uint32_t* wd = G.stage[st_cur];
                unpack4_smem(wd, lane, bd, g);
                unpack4_smem(wd + 4 * bd, lane, bt, t);
                st_blk[st_cur] = EMPTY;  // (the slot may be re-armed as soon as every lane has read it: __syncwarp in stage_issue)
                st_cur ^= 1;
            } else {
                const uint32_t* wd = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e.z * 16u);
                unpack4(wd, lane, bd, g);
                unpack4(wd + 4 * bd, lane, bt, t);
            }
            g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
            const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
            uint32_t d[4], val[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                d[j] = off + g[j] + (uint32_t)j;
                ok[j] = 4u * lane + (uint32_t)j < n;
            }
            if (acct) {
                A.block_bytes += ((n * bd + 7u) >> 3) + ((n * bt + 7u) >> 3) + 16u;
                A.lead_blocks++;
            }
            if (!has_req) {
                uint32_t id[4] = {0u, 0u, 0u, 0u};
                if (LD.fn_field >= 0) {
                    const uint8_t* __restrict__ fn = p.ix.fnorm[LD.fn_field];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (ok[j]) id[j] = __ldg(fn + d[j]);
                    if (acct) A.gathers += n;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float nrm = LD.fn_field >= 0 ? NORM_CACHE(S, p, LD.fn_field * 256 + (int)id[j]) : LD.cnorm;
                    const float s = LD.weight * tf_factor((float)(t[j] + 1u), nrm);
                    if (prune) ok[j] = ok[j] && (s + rest_s >= theta);
                    val[j] = __float_as_uint(s);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) val[j] = t[j] + 1u;  // required clauses first: the lead is scored for the survivors only
            }
            const unsigned m0 = __ballot_sync(FULL, ok[0]), m1 = __ballot_sync(FULL, ok[1]), m2 = __ballot_sync(FULL, ok[2]),
                           m3 = __ballot_sync(FULL, ok[3]);
            const uint32_t total = (uint32_t)(__popc(m0) + __popc(m1) + __popc(m2) + __popc(m3));
            if (total) {
                uint32_t pos = ncand + (uint32_t)(__popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt));
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (ok[j]) {
                        W.cand_doc[pos] = d[j];
                        W.cand_val[pos] = val[j];
                        pos++;
                    }
                ncand += total;
                __syncwarp();
            }
        }

        PH_MARK(PH_DECODE);
        // ================= evaluate a round of candidates =================
        // A full round (128: 4 per lane, so every lookup has 4 independent gathers in flight) as soon as it is there;
        // whatever is left when the walk ends. Candidate 32*r + lane of the round belongs to this lane: every r is an
        // ascending run over the lanes (what probe() needs).
        if (ncand >= (uint32_t)ROUND || (fin && ncand)) {
            const uint32_t cnt = min(ncand, (uint32_t)ROUND);
            const int nr = (int)((cnt + 31u) >> 5);  // rows in use (uniform): rows >= nr are skipped
            uint32_t c[4], v[4], tf[4], lv = 0u;
            float sc[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t idx = 32u * r + lane;
                c[r] = 0xFFFFFFFFu;
                v[r] = 0u;
                sc[r] = 0.f;
                tf[r] = 0u;
                if (idx < cnt) {
                    lv |= 1u << r;
                    c[r] = W.cand_doc[idx];
                    v[r] = W.cand_val[idx];
                }
            }
            float rem = LD.rest;
            if (!has_req) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    sc[r] = __uint_as_float(v[r]);
                    // (the threshold may have risen while the candidate waited)
                    if (prune && !(sc[r] + rest_s >= theta)) lv &= ~(1u << r);
                }
            }
            // The other leaves in evaluation order: required clauses (a candidate must occur in every one), the earlier
            // leads (a candidate that occurs in one is scored by that lead's items: dropped here), the later leads and
            // the optional leaves (add their scores), the excluded leaves (drop on a hit).
            uint32_t found = 0u, clause = has_req ? (W.leaf[n_lead].role >> 8) : 0u;
            const int n_steps = n_leaves - 1;
            for (int s = 0; s <= n_steps; s++) {
                if (has_req && s == n_req) {  // all required clauses seen: close the last one, score the lead
                    lv &= found;
                    float nl[4];
                    norms4(LD, c, lv, nl);
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if ((lv >> r) & 1u) {
                            sc[r] += LD.weight * tf_factor((float)v[r], nl[r]);
                            if (prune && !(sc[r] + rem + slack >= theta)) lv &= ~(1u << r);
                        }
                    if (acct) A.gathers_l += (unsigned long long)__popc(lv);
                }
                if (s == n_steps || !__any_sync(FULL, lv)) break;
                int j, mode;  // mode 0 required, 1 drop on hit, 2 add
                if (s < n_req) { j = n_lead + s; mode = 0; }
                else {
                    const int s2 = s - n_req;
                    if (s2 < lead) { j = s2; mode = 1; }
                    else if (s2 < n_lead - 1) { j = s2 + 1; mode = 2; }
                    else { j = s2 + 1 + n_req; mode = j < n_lead + n_req + (int)q.n_opt ? 2 : 1; }
                }
                const LLeaf& L = W.leaf[j];
                if (mode == 0) {
                    const uint32_t cl = L.role >> 8;
                    if (cl != clause) {
                        lv &= found;
                        found = 0u;
                        clause = cl;
                        if (!__any_sync(FULL, lv)) break;
                    }
                }
                const bool need_tf = mode != 1;
                // ---- term frequencies of the candidates in leaf j (0 = absent) ----
                if (L.col) {  // dense tf column: one byte per doc
#pragma unroll
                    for (int r = 0; r < 4; r++) tf[r] = ((lv >> r) & 1u) ? (uint32_t)__ldg(L.col + c[r]) : 0u;
                    if (acct) A.gathers_l += (unsigned long long)__popc(lv);
                } else if (L.bits) {
                    // membership bit first (one 4-byte gather per candidate, all in flight together; most lookups miss); a hit
                    // finds its posting through the rank directory: position = postings before the doc, block = position / 128
                    uint32_t wv[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) wv[r] = ((lv >> r) & 1u) ? __ldg(L.bits + (c[r] >> 5)) : 0u;
#pragma unroll
                    for (int r = 0; r < 4; r++) tf[r] = (wv[r] >> (c[r] & 31u)) & 1u;
                    if (need_tf && __any_sync(FULL, (tf[0] | tf[1] | tf[2] | tf[3]) != 0u)) {
                        // The hits' postings, all rows in step (each stage's loads of the four rows are in flight together;
                        // rows without a hit read element 0 of everything: no branch between the loads):
                        // rank directory + the words of the doc's 256-doc chunk before its own (same 32-byte sector as wv)
                        uint32_t rk[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) rk[r] = __ldg(L.rank + (tf[r] ? (c[r] >> 8) : 0u));
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint32_t ch = tf[r] ? (c[r] >> 8) : 0u, wi = tf[r] ? ((c[r] >> 5) & 7u) : 0u;
                            const uint4 a = __ldg(reinterpret_cast<const uint4*>(L.bits) + 2u * ch), b = __ldg(reinterpret_cast<const uint4*>(L.bits) + 2u * ch + 1u);
                            uint32_t n = (uint32_t)__popc(wv[r] & ((1u << (c[r] & 31u)) - 1u));
                            n += wi > 0u ? (uint32_t)__popc(a.x) : 0u;
                            n += wi > 1u ? (uint32_t)__popc(a.y) : 0u;
                            n += wi > 2u ? (uint32_t)__popc(a.z) : 0u;
                            n += wi > 3u ? (uint32_t)__popc(a.w) : 0u;
                            n += wi > 4u ? (uint32_t)__popc(b.x) : 0u;
                            n += wi > 5u ? (uint32_t)__popc(b.y) : 0u;
                            n += wi > 6u ? (uint32_t)__popc(b.z) : 0u;
                            rk[r] = tf[r] ? rk[r] + n : 0u;
                        }
                        // position -> block (rank / 128) -> its skip entry -> the two words of the tf stream holding the value
                        uint4 e2[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) e2[r] = __ldg(p.ix.skip + L.blk_begin + (rk[r] >> 7));
                        uint32_t lo[4], hi[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint32_t bd = e2[r].w & 63u, bt = (e2[r].w >> 6) & 63u;
                            const uint32_t* wt = reinterpret_cast<const uint32_t*>(p.ix.blk + (size_t)e2[r].z * 16u) + 4u * bd;
                            const uint32_t bit = (rk[r] & 127u) * bt;
                            lo[r] = __ldg(wt + (bit >> 5));
                            hi[r] = __ldg(wt + (bit >> 5) + 1);
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint32_t bt = (e2[r].w >> 6) & 63u, sh = ((rk[r] & 127u) * bt) & 31u;
                            const uint32_t m = bt >= 32u ? 0xFFFFFFFFu : ((1u << bt) - 1u);
                            if (tf[r]) {
                                tf[r] = (__funnelshift_r(lo[r], hi[r], sh) & m) + 1u;
                                if (acct) A.block_bytes_l += 56ull;
                            }
                        }
                    }
                    if (acct) A.gathers_l += 4ull * (unsigned long long)__popc(lv);
                } else {
#pragma unroll 1
                    for (int r = 0; r < nr; r++) {
                        const uint32_t cr = r == 0 ? c[0] : r == 1 ? c[1] : r == 2 ? c[2] : c[3];
                        const uint32_t t1 = probe(p, W, j, L, cr, ((lv >> r) & 1u) != 0u, need_tf, lane, A);
                        if (r == 0) tf[0] = t1; else if (r == 1) tf[1] = t1; else if (r == 2) tf[2] = t1; else tf[3] = t1;
                    }
                }
                PH_MARK(L.col ? PH_LOOKUP : (L.bits ? PH_LOOKUP_BM : PH_LOOKUP_PROBE));
                // ---- what a hit means ----
                if (mode == 1) {
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (tf[r]) lv &= ~(1u << r);
                } else {
                    rem -= L.ub;
                    const uint32_t hitm = (tf[0] ? 1u : 0u) | (tf[1] ? 2u : 0u) | (tf[2] ? 4u : 0u) | (tf[3] ? 8u : 0u);
                    if (__any_sync(FULL, hitm != 0u)) {
                        float nl[4];
                        norms4(L, c, hitm, nl);
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            if (tf[r]) sc[r] += L.weight * tf_factor((float)tf[r], nl[r]);
                        found |= hitm;
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (mode == 2 && prune && !(sc[r] + rem + slack >= theta)) lv &= ~(1u << r);
                }
#pragma unroll
                for (int r = 0; r < 4; r++) tf[r] = 0u;
                PH_MARK(PH_SCORE);
            }
            PH_MARK(PH_SCORE);
            if (p.ix.alive) {
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (((lv >> r) & 1u) && !((__ldg(p.ix.alive + (c[r] >> 5)) >> (c[r] & 31u)) & 1u)) lv &= ~(1u << r);
            }
            if (__any_sync(FULL, lv)) {
                uint32_t* hq = p.qhist + (size_t)item.query * LHIST_B;
#pragma unroll 1
                for (int r = 0; r < nr; r++) {
                    const uint32_t cr = r == 0 ? c[0] : r == 1 ? c[1] : r == 2 ? c[2] : c[3];
                    const float sr = (r == 0 ? sc[0] : r == 1 ? sc[1] : r == 2 ? sc[2] : sc[3]) + q.const_score;
                    const bool live = ((lv >> r) & 1u) != 0u;
                    if (p.exhaustive) {
                        n_match += (uint32_t)__popc(__ballot_sync(FULL, live));
                        if (p.match_bitmap && live)
                            atomicOr(p.match_bitmap + (size_t)item.query * p.bitmap_words + (cr >> 5), 1u << (cr & 31u));
                    }
                    if (KS == 0) {
                        const unsigned mk = __ballot_sync(FULL, live);
                        if (mk) {
                            uint32_t base = 0u;
                            if (lane == 0) base = atomicAdd(p.qcount + item.query, (uint32_t)__popc(mk));
                            base = __shfl_sync(FULL, base, 0) + (uint32_t)__popc(mk & lt);
                            if (live && base < q.part_cap) p.partial[q.part_begin + base] = make_key(sr, cr);
                        }
                    } else {
                        tk.offer(live, make_key(sr, cr), k, lane);
                    }
                    if (prune && live) {  // the query-wide histogram of accepted scores (fg_internal.h)
                        int idx = (int)(__float_as_uint(sr) >> LHIST_SHIFT) - (int)q.hist_base + (LHIST_B - 1);
                        idx = sr > 0.f ? min(max(idx, 0), LHIST_B - 1) : 0;
                        atomicAdd(hq + idx, 1u);
                    }
                }
                if (prune) {
                    // this warp's own k-th best, and the k-th best the histogram proves for the whole query
                    uint32_t best = (KS > 0 && tk.theta) ? (uint32_t)(tk.theta >> 32) : 0u;
                    __syncwarp();
                    const uint4 h4 = __ldcg(reinterpret_cast<const uint4*>(hq) + lane);  // buckets 4*lane .. 4*lane+3
                    const uint32_t mine = h4.x + h4.y + h4.z + h4.w;
                    uint32_t above = mine;  // inclusive suffix sum over the lanes
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_down_sync(FULL, above, o);
                        if (lane + o < 32) above += y;
                    }
                    above -= mine;  // counts in the buckets above this lane's
                    int hit = -1;   // highest bucket of this lane at which the count from the top reaches k
                    uint32_t acc = above + h4.w;
                    if (acc >= (uint32_t)k) hit = 4 * lane + 3;
                    else if ((acc += h4.z) >= (uint32_t)k) hit = 4 * lane + 2;
                    else if ((acc += h4.y) >= (uint32_t)k) hit = 4 * lane + 1;
                    else if ((acc += h4.x) >= (uint32_t)k) hit = 4 * lane;
#pragma unroll
                    for (int o = 16; o; o >>= 1) hit = max(hit, __shfl_xor_sync(FULL, hit, o));
                    if (hit > 0) {  // bucket 0 collects everything below the window: no bound
                        const float edge = __uint_as_float((q.hist_base - (uint32_t)(LHIST_B - 1) + (uint32_t)hit) << LHIST_SHIFT);
                        best = max(best, sortable(edge));
                    }
                    if (best > pub) {
                        pub = best;
                        theta = fmaxf(theta, unsortable(best));
                        if (lane == 0) atomicMax(p.qtheta + item.query, best);
                    }
                }
            }
            PH_MARK(PH_TOPK);
            // candidates behind the round move to the front
            const uint32_t left = ncand - cnt;
            if (left) {
                uint32_t cd0[4], cv0[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t idx = 32u * r + lane;
                    cd0[r] = idx < left ? W.cand_doc[cnt + idx] : 0u;
                    cv0[r] = idx < left ? W.cand_val[cnt + idx] : 0u;
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t idx = 32u * r + lane;
                    if (idx < left) {
                        W.cand_doc[idx] = cd0[r];
                        W.cand_val[idx] = cv0[r];
                    }
                }
            }
            ncand = left;
            __syncwarp();
        }
        if (fin && !ncand) break;
    }
    }  // leads of the item

    if (TMA) {
        stage_wait(0);
        stage_wait(1);
        __syncwarp();
        if (lane == 0) G.bar_parity = st_par;
    }

    PH_MARK(PH_OTHER);
    // append this warp's queue to the query's region of the partial array
    {
        // (the queue is sorted best first; ranks >= k are scratch)
        uint32_t nz = 0u;
#pragma unroll
        for (int s = 0; s < KS; s++) nz += (uint32_t)__popc(__ballot_sync(FULL, tk.q[s] != 0 && s * 32 + lane < k));
        if (KS > 0 && nz) {
            uint32_t base = 0u;
            if (lane == 0) base = atomicAdd(p.qcount + item.query, nz);
            base = __shfl_sync(FULL, base, 0) + q.part_begin;
#pragma unroll
            for (int s = 0; s < KS; s++) {
                const bool wr = tk.q[s] != 0 && s * 32 + lane < k;
                const unsigned mk = __ballot_sync(FULL, wr);
                if (wr) p.partial[base + (uint32_t)__popc(mk & lt)] = tk.q[s];
                base += (uint32_t)__popc(mk);
            }
        }
        if (acct) {  // lane-local counters: sum over the warp
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                A.gathers_l += __shfl_xor_sync(FULL, A.gathers_l, o);
                A.block_bytes_l += __shfl_xor_sync(FULL, A.block_bytes_l, o);
            }
        }
        if (lane == 0) {
            if (n_match) atomicAdd(p.qmatch + item.query, n_match);
            if (acct && p.stats) {
                atomicAdd(p.stats + 0, A.block_bytes + A.block_bytes_l);
                atomicAdd(p.stats + 2, A.gathers + A.gathers_l);
                atomicAdd(p.stats + 5, A.meta_bytes);
                atomicAdd(p.stats + 6, A.lead_blocks);
                atomicAdd(p.stats + 7, A.lead_blocks_seen);
            }
        }
    }
    PH_MARK(PH_APPEND);
    PH_FLUSH();
}

template <int KS, bool TMA>
__global__ void __launch_bounds__(LNT, KS <= 1 ? FG_LEAD_MINB : (KS <= 4 ? 3 : 1)) lead_kernel(const LeadParams p) {
    static_assert(KS > 0 || !TMA, "the deep-page variant is built without payload staging");
    FG_DYN_SMEM(smem);
    LeadShared& S = *reinterpret_cast<LeadShared*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpShared& W = S.w[warp];
    StageShared& G = reinterpret_cast<StageShared*>(smem + sizeof(LeadShared))[TMA ? warp : 0];  // (present only when TMA)
    if (TMA) {
        if (lane == 0) {
            mbar_init(&G.bar[0], 1u);
            mbar_init(&G.bar[1], 1u);
            fence_mbar_init();
            G.bar_parity = 0u;
        }
        __syncwarp();
    }
    // The work queue hands out ONE item per atomic, in array order: items are sorted by lead index, so a query's short
    // leads have run -- and set its threshold -- before its long leads start. (Claiming 4 items per atomic was measured:
    // the first wave then reaches 4x deeper into the array, later leads start cold, 80 % more blocks get decoded.)
    // dev tool (FG_PROF + FG_EXEC_COUNTERS): busy time of the warps, longest item, when the first warp ran out of work
    const bool prof = p.prof != 0u;
    unsigned long long t_busy = 0ull, t_longest = 0ull;
    if (prof && lane == 0) atomicMax(p.stats + 13, ~global_timer_ns());  // (complement: the earliest start wins)
    while (true) {
        uint32_t it = 0u;
        if (lane == 0) it = atomicAdd(p.work, 1u);
        it = __shfl_sync(FULL, it, 0);
        if (it >= p.n_items) break;
        const LItem item = p.items[it];
        // pruned form: an item none of whose documents can reach the query's current threshold costs two loads
        // (lane 0's view for the whole warp: the threshold may move between two lanes' loads)
        const uint32_t th = p.exhaustive ? 0u : __shfl_sync(FULL, __ldcg(p.qtheta + item.query), 0);
        if (!p.exhaustive && item.bound < unsortable(th)) continue;
        const unsigned long long t0 = prof ? global_timer_ns() : 0ull;
        run_item<KS, TMA>(p, S, W, G, item, th, lane);
        if (prof) {
            const unsigned long long dt = global_timer_ns() - t0;
            t_busy += dt;
            t_longest = max(t_longest, dt);
        }
    }
    if (prof && lane == 0) {
        const unsigned long long t = global_timer_ns();
        atomicAdd(p.stats + 8, t_busy);
        atomicMax(p.stats + 9, t_longest);
        atomicMax(p.stats + 10, ~t);  // first warp out of work
        atomicMax(p.stats + 11, t);   // last warp done
        atomicAdd(p.stats + 12, 1ull);
    }
}

// resets the per-execution device state of a prepared batch
__global__ void lead_init_kernel(const LeadParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *p.work = 0u;
    if (i < p.n_cursors) p.cursors[i] = 0u;
    if (i < p.n_queries) {
        p.qtheta[i] = p.exhaustive ? 0u : p.queries[i].theta0;
        p.qcount[i] = 0u;
        p.qmatch[i] = 0u;
    }
}

// one warp per query: the k best of the entries the lead warps appended
template <int KS>
__global__ void __launch_bounds__(128) lead_merge_kernel(const LeadMergeParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= p.n_queries) return;
    const LQuery q = p.queries[qi];
    const int k = (int)q.k;
    uint2* out = reinterpret_cast<uint2*>(p.out_hits) + (size_t)qi * p.k_stride;
    if (q.flags & LQ_ALL) {  // AllQuery: the first k alive docs (every score equals const_score)
        uint32_t found = 0;
        for (uint32_t base = 0; base < p.n_docs && found < (uint32_t)k; base += 32) {
            const uint32_t d = base + lane;
            const bool al = d < p.n_docs && (!p.alive || ((p.alive[d >> 5] >> (d & 31)) & 1u));
            const unsigned m = __ballot_sync(FULL, al);
            const uint32_t r = found + __popc(m & ((1u << lane) - 1u));
            if (al && r < (uint32_t)k && r < p.k_stride) out[r] = make_uint2(__float_as_uint(q.const_score), d + p.doc_base);
            found += __popc(m);
        }
        const uint32_t nh_all = min(min(found, (uint32_t)k), p.k_stride);
        for (uint32_t r = nh_all + lane; r < p.k_stride; r += 32) out[r] = make_uint2(0u, 0xFFFFFFFFu);
        if (lane == 0) {
            p.out_n[qi] = nh_all;
            if (p.out_count) p.out_count[qi] = p.n_alive;
        }
        return;
    }
    WarpTopK<KS> tk;
    tk.init();
    const uint64_t* src = p.partial + q.part_begin;
    const uint32_t total = min(p.qcount[qi], q.part_cap);
    for (uint32_t i0 = 0; i0 < total; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint64_t c = i < total ? src[i] : 0;
        tk.offer(c != 0, c, k, lane);
    }
    uint32_t nh = 0;
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
    for (uint32_t r = KS * 32 + lane; r < p.k_stride; r += 32) out[r] = make_uint2(0u, 0xFFFFFFFFu);
    if (lane == 0) {
        p.out_n[qi] = nh;
        if (p.out_count) p.out_count[qi] = p.qmatch[qi];
    }
}

// CTA-wide (256 threads): the k best of `total` distinct 64-bit keys at src (0 = empty slot), sorted, written as hits to
// out[0 .. k_stride) (padded); returns the number of hits. sel = scratch of the next power of two >= k keys.
__device__ uint32_t select_page(const uint64_t* src, uint32_t total, uint32_t k_want, uint64_t* sel, uint2* out, uint32_t k_stride,
                                uint32_t doc_base, uint32_t* hist, uint64_t* s_prefix_p, uint32_t* s_remaining_p, uint32_t* s_cnt_p) {
    const int tid = threadIdx.x;
    uint64_t& s_prefix = *s_prefix_p;
    uint32_t &s_remaining = *s_remaining_p, &s_cnt = *s_cnt_p;
    const uint32_t k = min(k_want, k_stride);
    uint32_t cap2 = 1u;
    while (cap2 < k) cap2 <<= 1;
    uint64_t T = 0;  // the k-th largest key
    if (total > k) {
        uint64_t prefix = 0;
        uint32_t remaining = k;
        for (int shift = 56; shift >= 0; shift -= 8) {
            hist[tid] = 0u;
            __syncthreads();
            for (uint32_t i = tid; i < total; i += 256u) {
                const uint64_t key = src[i];
                if (shift == 56 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t acc = 0u;
                int d = 255;
                for (; d > 0; d--) {
                    if (acc + hist[d] >= remaining) break;
                    acc += hist[d];
                }
                s_prefix = (prefix << 8) | (uint64_t)d;
                s_remaining = remaining - acc;
            }
            __syncthreads();
            prefix = s_prefix;
            remaining = s_remaining;
            __syncthreads();
        }
        T = prefix;
    }
    if (tid == 0) s_cnt = 0u;
    __syncthreads();
    for (uint32_t i = tid; i < total; i += 256u) {
        const uint64_t key = src[i];
        if (key >= T && key != 0) {
            const uint32_t pos = atomicAdd(&s_cnt, 1u);
            if (pos < cap2) sel[pos] = key;
        }
    }
    __syncthreads();
    const uint32_t n_sel = min(s_cnt, cap2);
    for (uint32_t i = n_sel + tid; i < cap2; i += 256u) sel[i] = 0;
    __syncthreads();
    for (uint32_t size = 2u; size <= cap2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride; stride >>= 1) {
            for (uint32_t i = tid; i < cap2 / 2u; i += 256u) {
                const uint32_t pos = 2u * i - (i & (stride - 1u));
                const bool desc = (pos & size) == 0u;
                const uint64_t a = sel[pos], b = sel[pos + stride];
                if ((a < b) == desc) {
                    sel[pos] = b;
                    sel[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
    const uint32_t nh = min(n_sel, k);
    for (uint32_t r = tid; r < k_stride; r += 256u) {
        uint2 h = make_uint2(0u, 0xFFFFFFFFu);
        if (r < nh) {
            const uint64_t key = sel[r];
            h.x = __float_as_uint(unsortable((uint32_t)(key >> 32)));
            h.y = ~(uint32_t)(key & 0xFFFFFFFFu) + doc_base;
        }
        out[r] = h;
    }
    return nh;
}

// Deep pages (k > 1024, any page * per_page + per_page the handler lets through: /root/reference/src/db/search.rs:154-160):
// the lead warps appended every accepted hit to the query's partial region. One CTA per query: radix select of the
// k-th largest 64-bit key (most significant byte first; keys are distinct, the doc id is part of them), compaction of
// the k keys at or above it into the query's scratch region, bitonic sort there (padded to a power of two), page out.
__global__ void __launch_bounds__(256) lead_select_kernel(const LeadMergeParams p) {
    __shared__ uint32_t hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_remaining, s_cnt;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t qi = blockIdx.x;
    const LQuery q = p.queries[qi];
    uint2* out = reinterpret_cast<uint2*>(p.out_hits) + (size_t)qi * p.k_stride;
    if (q.flags & LQ_ALL) {  // (uniform) AllQuery: the first k alive docs, every score equals const_score
        if (tid >= 32) return;
        uint32_t found = 0;
        for (uint32_t base = 0; base < p.n_docs && found < q.k; base += 32) {
            const uint32_t d = base + lane;
            const bool al = d < p.n_docs && (!p.alive || ((p.alive[d >> 5] >> (d & 31)) & 1u));
            const unsigned m = __ballot_sync(FULL, al);
            const uint32_t r = found + __popc(m & ((1u << lane) - 1u));
            if (al && r < q.k && r < p.k_stride) out[r] = make_uint2(__float_as_uint(q.const_score), d + p.doc_base);
            found += __popc(m);
        }
        const uint32_t nh_all = min(min(found, q.k), p.k_stride);
        for (uint32_t r = nh_all + lane; r < p.k_stride; r += 32) out[r] = make_uint2(0u, 0xFFFFFFFFu);
        if (lane == 0) {
            p.out_n[qi] = nh_all;
            if (p.out_count) p.out_count[qi] = p.n_alive;
        }
        return;
    }
    const uint64_t* src = p.partial + q.part_begin;
    const uint32_t total = min(p.qcount[qi], q.part_cap);
    const uint32_t nh = select_page(src, total, min(q.k, total), p.sel + q.sel_begin, out, p.k_stride, p.doc_base, hist, &s_prefix, &s_remaining, &s_cnt);
    if (tid == 0) {
        p.out_n[qi] = nh;
        if (p.out_count) p.out_count[qi] = p.qmatch[qi];
    }
}

// fg_search_union_of: the queries of the batch are the DISJUNCTS of one query (tantivy: a BooleanQuery whose Should
// children are themselves boolean queries, e.g. `(a AND b) OR (c AND d)`; a document's score is the sum of the scores of
// the disjuncts it matches). The lead warps appended every match of every disjunct (deep-page form, no pruning: k is
// unbounded). One CTA: gather the m lists as (~doc, score) entries, bitonic sort (equal docs become neighbours, in a
// value-determined order), one combined key per run of equal docs (scores summed in run order), then the ordinary
// page selection over the combined keys. a, b: scratch of cap2 keys each (cap2 = power of two >= all matches).
// n_filter: the LAST n_filter queries are FILTER children (tantivy: Bool[Must(the union of the other children), Must(f)..],
// what Dataset::search builds from a nested text query and a facet query, src/db/search.rs:140-144): a document counts
// only when an ordinary child AND every filter child hold it; the filter scores are added after the ordinary sum. Their
// entries are tagged by clearing the top bit of sortable(score) (set for every score >= 0), which also sorts them to the
// end of their document's run.
__global__ void __launch_bounds__(256) lead_combine_kernel(const LeadMergeParams p, uint32_t k, uint64_t* a, uint64_t* b, uint32_t cap2,
                                                           uint32_t n_filter) {
    __shared__ uint32_t hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_remaining, s_cnt, s_fill, s_runs;
    const int tid = threadIdx.x;
    if (tid == 0) s_fill = 0u, s_runs = 0u;
    __syncthreads();
    for (uint32_t j = 0; j < p.n_queries; j++) {
        const LQuery q = p.queries[j];
        const uint64_t* src = p.partial + q.part_begin;
        const uint32_t cnt = min(p.qcount[j], q.part_cap), base = s_fill;
        const uint64_t tag = j + n_filter >= p.n_queries ? ~0x80000000ull : ~0ull;
        for (uint32_t i = tid; i < cnt; i += 256u) {
            const uint64_t key = src[i];
            if (base + i < cap2) a[base + i] = ((key << 32) | (key >> 32)) & tag;  // (~doc) above sortable(score)
        }
        __syncthreads();
        if (tid == 0) s_fill = min(base + cnt, cap2);
        __syncthreads();
    }
    const uint32_t n = s_fill;
    for (uint32_t i = n + tid; i < cap2; i += 256u) a[i] = 0;
    __syncthreads();
    for (uint32_t size = 2u; size <= cap2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride; stride >>= 1) {
            for (uint32_t i = tid; i < cap2 / 2u; i += 256u) {
                const uint32_t pos = 2u * i - (i & (stride - 1u));
                const bool desc = (pos & size) == 0u;
                const uint64_t x = a[pos], y = a[pos + stride];
                if ((x < y) == desc) {
                    a[pos] = y;
                    a[pos + stride] = x;
                }
            }
            __syncthreads();
        }
    }
    // runs of equal docs -> one key each (at the run's first position; 0 elsewhere)
    for (uint32_t i = tid; i < cap2; i += 256u) {
        uint64_t outk = 0;
        const uint64_t e = i < n ? a[i] : 0;
        if (e != 0 && (i == 0 || (a[i - 1] >> 32) != (e >> 32))) {
            float sum = 0.f, fsum = 0.f;
            uint32_t nt = 0, nf = 0;
            for (uint32_t t = i; t < n && (a[t] >> 32) == (e >> 32); t++) {
                const uint32_t bits = (uint32_t)(a[t] & 0xFFFFFFFFu);
                if (n_filter && !(bits & 0x80000000u)) fsum += unsortable(bits | 0x80000000u), nf++;
                else sum += unsortable(bits), nt++;
            }
            if (n_filter == 0u || (nt && nf == n_filter)) {
                outk = ((uint64_t)sortable(sum + fsum) << 32) | (e >> 32);
                atomicAdd(&s_runs, 1u);
            }
        }
        b[i] = outk;
    }
    __syncthreads();
    const uint32_t n_runs = s_runs;
    const uint32_t nh = select_page(b, cap2, min(k, n_runs), a, reinterpret_cast<uint2*>(p.out_hits), p.k_stride, p.doc_base, hist, &s_prefix,
                                    &s_remaining, &s_cnt);
    if (tid == 0) {
        p.out_n[0] = nh;
        if (p.out_count) p.out_count[0] = n_runs;
    }
}

// one warp per query: merge the all-gathered per-rank top-k lists; the query's k is read from the
// batch's query array (word k_word of a record of q_words words)
template <int KS>
__global__ void __launch_bounds__(128) merge_ranks_kernel(const uint2* __restrict__ hits, const uint32_t* __restrict__ n, uint32_t n_ranks,
                                                          uint32_t n_queries, const uint32_t* __restrict__ qrec, uint32_t q_words,
                                                          uint32_t k_word, uint32_t k_stride, uint2* out_hits, uint32_t* out_n) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= n_queries) return;
    const int k = (int)min(__ldg(qrec + (size_t)qi * q_words + k_word), k_stride);
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
            tk.offer(c != 0, c, k, lane);
        }
    }
    uint32_t nh = 0;
#pragma unroll
    for (int s = 0; s < KS; s++) {
        const int r = s * 32 + lane;
        const bool ok = r < k && tk.q[s] != 0;
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
    for (uint32_t r = KS * 32 + lane; r < k_stride; r += 32) out_hits[(size_t)qi * k_stride + r] = make_uint2(0u, 0xFFFFFFFFu);
    if (lane == 0) out_n[qi] = nh;
}

// block-max metadata: one warp per block, the same arithmetic as the scoring path
__global__ void __launch_bounds__(256) blockmax_kernel(const DevIndex ix, uint32_t b0, uint32_t b1, int fn_field,
                                                       float cnorm, float* bmax) {
    const int lane = threadIdx.x & 31;
    const uint32_t b = b0 + blockIdx.x * 8u + (threadIdx.x >> 5);
    if (b >= b1) return;
    const uint4 e = __ldg(&ix.skip[b]);
    const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u, n = ((e.w >> 12) & 127u) + 1u;
    const uint32_t* wd = reinterpret_cast<const uint32_t*>(ix.blk + (size_t)e.z * 16u);
    uint32_t g[4], t[4];
    unpack4(wd, lane, bd, g);
    unpack4(wd + 4 * bd, lane, bt, t);
    g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
    const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
    float best = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (4u * lane + (uint32_t)j < n) {
            const uint32_t d = off + g[j] + (uint32_t)j;
            const float nrm = fn_field >= 0 ? __ldg(ix.cache + fn_field * 256 + (int)__ldg(ix.fnorm[fn_field] + d)) : cnorm;
            best = fmaxf(best, tf_factor((float)(t[j] + 1u), nrm));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(FULL, best, o));
    if (lane == 0) bmax[b] = best;
}

// membership bitmaps of the mid-frequency terms: one warp per selected block sets the bits of its 128 docs
__global__ void __launch_bounds__(256) bitmap_build_kernel(const DevIndex ix, const uint2* __restrict__ sel, uint32_t n_sel,
                                                           uint32_t* bits, uint64_t stride_words) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (i >= n_sel) return;
    const uint2 sb = __ldg(&sel[i]);
    const uint4 e = __ldg(&ix.skip[sb.x]);
    const uint32_t bd = e.w & 63u, n = ((e.w >> 12) & 127u) + 1u;
    const uint32_t* wd = reinterpret_cast<const uint32_t*>(ix.blk + (size_t)e.z * 16u);
    uint32_t g[4];
    unpack4(wd, lane, bd, g);
    g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
    const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
    uint32_t* dst = bits + (size_t)sb.y * stride_words;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (4u * lane + (uint32_t)j < n) {
            const uint32_t d = off + g[j] + (uint32_t)j;
            atomicOr(dst + (d >> 5), 1u << (d & 31u));
        }
}
// rank directory: postings before each 256-doc chunk; one CTA per bitmap, tiles of 256 chunks
__global__ void __launch_bounds__(256) bitmap_rank_kernel(const uint32_t* __restrict__ bits, uint32_t* rank, uint64_t stride_words) {
    __shared__ uint32_t part[8];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* src = bits + (size_t)blockIdx.x * stride_words;
    uint32_t* dst = rank + (size_t)blockIdx.x * (stride_words / 8);
    const uint32_t n_chunks = (uint32_t)(stride_words / 8);
    if (tid == 0) carry = 0u;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 256u) {
        const uint32_t c = c0 + (uint32_t)tid;
        uint32_t cnt = 0u;
        if (c < n_chunks) {
            const uint4 a = reinterpret_cast<const uint4*>(src)[2 * (size_t)c], b = reinterpret_cast<const uint4*>(src)[2 * (size_t)c + 1];
            cnt = (uint32_t)(__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w));
        }
        const uint32_t ex = warp_excl_scan(cnt, lane);
        if (lane == 31) part[warp] = ex + cnt;
        __syncthreads();
        uint32_t base = carry;
        for (int w = 0; w < warp; w++) base += part[w];
        if (c < n_chunks) dst[c] = base + ex;
        __syncthreads();
        if (tid == 255) carry = base + ex + cnt;
        __syncthreads();
    }
}

// one warp per item record: the copies of the item, at their place in the queue
__global__ void __launch_bounds__(256) expand_items_kernel(const LItemRec* __restrict__ recs, uint32_t n_recs, LItem* items) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (i >= n_recs) return;
    const LItemRec r = recs[i];
    const uint4 v = make_uint4(r.item.query, r.item.lead, r.item.cursor, __float_as_uint(r.item.bound));
    for (uint32_t c = lane; c < r.copies; c += 32u) reinterpret_cast<uint4*>(items)[r.dst + c] = v;
}

// ---- snapshot append (fg_index_append) ----
// one warp per listed block: its skip entry's (last_doc, first_base) and its decoded postings (doc ids and tfs,
// 128 slots each; slots >= n undefined) go to a staging buffer the host reads back
__global__ void __launch_bounds__(256) tail_decode_kernel(const DevIndex ix, const uint32_t* __restrict__ blocks, uint32_t n_list,
                                                          uint32_t* out_meta, uint32_t* out_docs, uint32_t* out_tfs) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (i >= n_list) return;
    const uint4 e = __ldg(&ix.skip[__ldg(blocks + i)]);
    if (lane == 0) {
        out_meta[2 * i] = e.x;
        out_meta[2 * i + 1] = e.y;
    }
    if (!out_docs) return;  // (uniform) only the skip entry was asked for
    const uint32_t bd = e.w & 63u, bt = (e.w >> 6) & 63u;
    const uint32_t* wd = reinterpret_cast<const uint32_t*>(ix.blk + (size_t)e.z * 16u);
    uint32_t g[4], t[4];
    unpack4(wd, lane, bd, g);
    unpack4(wd + 4 * bd, lane, bt, t);
    g[1] += g[0]; g[2] += g[1]; g[3] += g[2];
    const uint32_t off = warp_excl_scan(g[3], lane) + e.y + 4u * lane;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        out_docs[(size_t)i * BLOCK + 4u * lane + j] = off + g[j] + (uint32_t)j;
        out_tfs[(size_t)i * BLOCK + 4u * lane + j] = t[j] + 1u;
    }
}
// one warp per range {src_begin, dst_begin, count}: dst[dst_begin + j] = src[src_begin + j] (16-byte entries)
__global__ void __launch_bounds__(256) copy_ranges_kernel(const uint4* __restrict__ src, uint4* dst, const uint3* __restrict__ ranges, uint32_t n_ranges) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (i >= n_ranges) return;
    const uint3 r = ranges[i];
    for (uint32_t j = lane; j < r.z; j += 32u) dst[r.y + j] = __ldg(&src[r.x + j]);
}

}  // namespace

void launch_expand_items(const LItemRec* recs, uint32_t n_recs, LItem* items, void* stream) {
    if (!n_recs) return;
    FG_LAUNCH(expand_items_kernel, (n_recs + 7) / 8, 256, 0, (cudaStream_t)stream, recs, n_recs, items);
}
void launch_tail_decode(const DevIndex& ix, const uint32_t* blocks, uint32_t n_list, uint32_t* out_meta, uint32_t* out_docs, uint32_t* out_tfs, void* stream) {
    if (!n_list) return;
    FG_LAUNCH(tail_decode_kernel, (n_list + 7) / 8, 256, 0, (cudaStream_t)stream, ix, blocks, n_list, out_meta, out_docs, out_tfs);
}
void launch_copy_ranges(const void* src, void* dst, const void* ranges, uint32_t n_ranges, void* stream) {
    if (!n_ranges) return;
    FG_LAUNCH(copy_ranges_kernel, (n_ranges + 7) / 8, 256, 0, (cudaStream_t)stream, (const uint4*)src, (uint4*)dst, (const uint3*)ranges, n_ranges);
}

void launch_bitmap_build(const DevIndex& ix, const uint2* sel, uint32_t n_sel, uint32_t* bits, uint64_t stride_words, void* stream) {
    if (!n_sel) return;
    FG_LAUNCH(bitmap_build_kernel, (n_sel + 7) / 8, 256, 0, (cudaStream_t)stream, ix, sel, n_sel, bits, stride_words);
}
void launch_bitmap_rank(const uint32_t* bits, uint32_t* rank, uint32_t n_slots, uint64_t stride_words, void* stream) {
    if (!n_slots) return;
    FG_LAUNCH(bitmap_rank_kernel, n_slots, 256, 0, (cudaStream_t)stream, bits, rank, stride_words);
}

void launch_lead(const LeadParams& p, int ks, int n_sms, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p.n_queries == 0) return;
    const uint32_t n_init = max(max(p.n_cursors, p.n_queries), 1u);
    FG_LAUNCH(lead_init_kernel, (n_init + 255) / 256, 256, 0, st, p);
    if (p.n_items == 0) return;
    const int smem = (int)sizeof(LeadShared) + (p.tma ? (int)(LNW * sizeof(StageShared)) : 0);
    static bool configured = false;
    if (!configured) {
        const int s0 = (int)sizeof(LeadShared), s1 = s0 + (int)(LNW * sizeof(StageShared));
        cudaFuncSetAttribute(lead_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s0);
        cudaFuncSetAttribute(lead_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s0);
        cudaFuncSetAttribute(lead_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s0);
        cudaFuncSetAttribute(lead_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s0);
        cudaFuncSetAttribute(lead_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
        cudaFuncSetAttribute(lead_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
        configured = true;
    }
    const unsigned per_sm = ks <= 1 ? (unsigned)FG_LEAD_MINB : (ks <= 4 ? 3u : 1u);
    const unsigned grid = min((p.n_items + LNW - 1) / LNW, (unsigned)n_sms * per_sm);
    // p.tma (FG_LEAD_TMA=1): the variant that stages lead-block payloads with 1-D bulk copies -- measured 9 % slower
    // than plain loads of L2-prefetched payloads on the C2 mix (profiles/r02_tma_ab.txt), so it is not the default
    if (ks == 0) {
        FG_LAUNCH((lead_kernel<0, false>), grid, LNT, (int)sizeof(LeadShared), st, p);
    } else if (ks <= 1) {
        if (p.tma) FG_LAUNCH((lead_kernel<1, true>), grid, LNT, smem, st, p);
        else FG_LAUNCH((lead_kernel<1, false>), grid, LNT, smem, st, p);
    } else if (ks <= 4) {
        if (p.tma) FG_LAUNCH((lead_kernel<4, true>), grid, LNT, smem, st, p);
        else FG_LAUNCH((lead_kernel<4, false>), grid, LNT, smem, st, p);
    } else {
        FG_LAUNCH((lead_kernel<32, false>), grid, LNT, (int)sizeof(LeadShared), st, p);
    }
}

void launch_lead_merge(const LeadMergeParams& p, int ks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p.n_queries == 0) return;
    const unsigned grid = (p.n_queries + 3) / 4;
    if (ks == 0) FG_LAUNCH(lead_select_kernel, p.n_queries, 256, 0, st, p);
    else if (ks <= 1) FG_LAUNCH(lead_merge_kernel<1>, grid, 128, 0, st, p);
    else if (ks <= 4) FG_LAUNCH(lead_merge_kernel<4>, grid, 128, 0, st, p);
    else FG_LAUNCH(lead_merge_kernel<32>, grid, 128, 0, st, p);
}

void launch_lead_combine(const LeadMergeParams& p, uint32_t k, uint64_t* a, uint64_t* b, uint32_t cap2, uint32_t n_filter, void* stream) {
    FG_LAUNCH(lead_combine_kernel, 1, 256, 0, (cudaStream_t)stream, p, k, a, b, cap2, n_filter);
}

void launch_merge_ranks(const void* hits, const uint32_t* n, uint32_t n_ranks, uint32_t n_queries, const void* qrec, uint32_t q_words,
                        uint32_t k_word, uint32_t k_stride, void* out_hits, uint32_t* out_n, int ks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_queries == 0) return;
    const unsigned grid = (n_queries + 3) / 4;
    if (ks <= 1)
        FG_LAUNCH(merge_ranks_kernel<1>, grid, 128, 0, st, (const uint2*)hits, n, n_ranks, n_queries, (const uint32_t*)qrec, q_words, k_word,
                  k_stride, (uint2*)out_hits, out_n);
    else if (ks <= 4)
        FG_LAUNCH(merge_ranks_kernel<4>, grid, 128, 0, st, (const uint2*)hits, n, n_ranks, n_queries, (const uint32_t*)qrec, q_words, k_word,
                  k_stride, (uint2*)out_hits, out_n);
    else
        FG_LAUNCH(merge_ranks_kernel<32>, grid, 128, 0, st, (const uint2*)hits, n, n_ranks, n_queries, (const uint32_t*)qrec, q_words, k_word,
                  k_stride, (uint2*)out_hits, out_n);
}

void launch_blockmax(const DevIndex& ix, uint32_t b0, uint32_t b1, int fn_field, float cnorm, float* bmax, void* stream) {
    if (b1 <= b0) return;
    FG_LAUNCH(blockmax_kernel, (b1 - b0 + 7) / 8, 256, 0, (cudaStream_t)stream, ix, b0, b1, fn_field, cnorm, bmax);
}

}  // namespace fg
