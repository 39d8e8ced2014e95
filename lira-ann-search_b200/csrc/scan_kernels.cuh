// K2: grouped IVF-Flat list scan with fused per-(query, list) top-k, and K4: per-query merge
// with id de-duplication. Replaces the arithmetic of
//   - get_cmp_recall (LIRA_smallscale.py:145-174): one exact top-k per (query, list) pair,
//   - search.cpp:468-514: scan of the probed lists + global top-k,
//   - faiss IndexFlat{L2,IP}.search as called from those sites.
//
// Work decomposition: the probe sets are inverted into per-list query groups, so one work item is
// (list b, up to TM queries of b's group). The CTA streams list b ONCE through the TMA ring
// (128 vectors x 32 floats per stage) while the item's query rows are re-fed from L2 by cp.async;
// 8 consumer warps keep a TM x 128 tile of distances in registers, spill it to a shared staging
// tile, and every query row is then screened by one warp against that row's running k-th best
// (a 64-bit (score, position) key), survivors being inserted into a warp-distributed sorted list.
#pragma once
#include "tile_engine.cuh"

namespace lira {

struct ScanItem {
    int list;     // list (partition) id
    int q_begin;  // first slot (position in the grouped probe order) of this item
    int q_count;  // valid query rows (<= tm)
    int tm;       // tile height class: 64, 32, 16 or 8
};

struct ScanParams {
    const float* q;              // [Q, ldq] queries (device)
    long ldq;
    int d;                       // padded dimension (multiple of 4) == K extent
    const int* group_queries;    // [P] query id per slot
    const long long* list_offsets;  // [B+1]
    const int* list_ids;         // [E] global id per list entry
    const ScanItem* items;
    const int* n_items;          // device scalar written by build_items_kernel
    unsigned long long* part_key;  // [P, k] output: sorted keys (score, id-or-position), KEY_INF padded
    int k;
    int store_local;             // 1: low word = position in list (IndexFlat label), 0: global id
    int max_rows;                // > 0: scan only the first max_rows entries of each list (seed pass)
};

static constexpr int SCAN_TM_MAX = 64;
static constexpr int SCAN_NSTAGE = 5;
static constexpr int SCAN_STAGE_BYTES = SCAN_TM_MAX * ROW_BYTES + B_STAGE_BYTES;  // uniform ring layout, 24 KiB
static constexpr int SCAN_A_BYTES = SCAN_TM_MAX * ROW_BYTES;

template <int S>
__host__ __device__ constexpr size_t scan_smem_bytes() {
    return 1024 /*align slack*/ + (size_t)SCAN_NSTAGE * SCAN_STAGE_BYTES + (size_t)SCAN_TM_MAX * DT_LD * 4 +
           (size_t)SCAN_TM_MAX * 32 * S * 8 + SCAN_TM_MAX * 4 + 2 * SCAN_NSTAGE * 8 + 64;
}

// uniform-layout variants of produce/consume (A region is SCAN_TM_MAX rows regardless of TM)
template <int TM>
__device__ __forceinline__ void scan_produce_kstep(uint8_t* stages, uint64_t* full_bar, uint64_t* empty_bar,
                                                   PipeState& ps, const CUtensorMap* tmap_b, int b_row0,
                                                   const float* __restrict__ a_base, long lda, int kdim,
                                                   const int (&arow)[TM / 4], int kc, int lane) {
    mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1u);
    uint8_t* sA = stages + (size_t)ps.stage * SCAN_STAGE_BYTES;
    uint8_t* sB = sA + SCAN_A_BYTES;
    if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[ps.stage], B_STAGE_BYTES);
        tma_load_2d(sB, tmap_b, kc * KC, b_row0, &full_bar[ps.stage]);
    }
    const int c = lane & 7;
    const int col = kc * KC + c * 4;
    const bool col_ok = col < kdim;
    const uint32_t sA_u32 = smem_u32(sA);
#pragma unroll
    for (int t = 0; t < TM / 4; ++t) {
        const int r = (lane >> 3) + 4 * t;
        const int g = arow[t];
        const bool ok = col_ok && g >= 0;
        const float* src = ok ? (a_base + (long)g * lda + col) : a_base;
        cp_async_16(sA_u32 + r * ROW_BYTES + ((c ^ (r & 7)) << 4), src, ok ? 16u : 0u);
    }
    cp_async_mbar_arrive_noinc(&full_bar[ps.stage]);
    ps.advance(SCAN_NSTAGE);
}

template <int TM, int OP>
__device__ __forceinline__ void scan_consume_tile(uint8_t* stages, uint64_t* full_bar, uint64_t* empty_bar,
                                                  PipeState& ps, int nk,
                                                  float (&acc)[TileCfg<TM>::RQ][TileCfg<TM>::RV], int q0, int v0,
                                                  int lane) {
    using C = TileCfg<TM>;
#pragma unroll
    for (int i = 0; i < C::RQ; ++i)
#pragma unroll
        for (int j = 0; j < C::RV; ++j) acc[i][j] = 0.0f;
    for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full_bar[ps.stage], ps.phase);
        const uint8_t* sA = stages + (size_t)ps.stage * SCAN_STAGE_BYTES;
        consume_kstep<TM, OP>(sA, sA + SCAN_A_BYTES, acc, q0, v0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[ps.stage]);
        ps.advance(SCAN_NSTAGE);
    }
}

// ---- producer: one work item ---------------------------------------------------------------
template <int TM>
__device__ __forceinline__ void scan_item_producer(const ScanParams& p, const ScanItem& it, const CUtensorMap* tmap,
                                                   uint8_t* stages, uint64_t* full_bar, uint64_t* empty_bar,
                                                   PipeState& ps, int lane) {
    int arow[TM / 4];
#pragma unroll
    for (int t = 0; t < TM / 4; ++t) {
        const int r = (lane >> 3) + 4 * t;
        arow[t] = (r < it.q_count) ? __ldg(p.group_queries + it.q_begin + r) : -1;
    }
    const long long lo = p.list_offsets[it.list];
    long long hi = p.list_offsets[it.list + 1];
    if (p.max_rows > 0 && hi - lo > p.max_rows) hi = lo + p.max_rows;
    const int nk = (p.d + KC - 1) / KC;
    for (long long row0 = lo; row0 < hi; row0 += TN)
        for (int kc = 0; kc < nk; ++kc)
            scan_produce_kstep<TM>(stages, full_bar, empty_bar, ps, tmap, (int)row0, p.q, p.ldq, p.d, arow, kc, lane);
}

// ---- consumers: one work item --------------------------------------------------------------
template <int TM, int OP, int S>
__device__ __forceinline__ void scan_item_consumer(const ScanParams& p, const ScanItem& it, uint8_t* stages,
                                                   uint64_t* full_bar, uint64_t* empty_bar, PipeState& ps,
                                                   float* dt, unsigned long long* tk, float* thr_f, int warp,
                                                   int lane) {
    using C = TileCfg<TM>;
    const int tid = warp * 32 + lane;
    const long long lo = p.list_offsets[it.list];
    long long hi = p.list_offsets[it.list + 1];
    if (p.max_rows > 0 && hi - lo > p.max_rows) hi = lo + p.max_rows;
    const int nk = (p.d + KC - 1) / KC;
    const int k = p.k;
    int q0, v0;
    consumer_coords<TM>(warp, lane, q0, v0);

    // reset the per-row state. Safe without a leading barrier: every row-screen of the previous item
    // ended before its final barrier (see end of this function).
    for (int i = tid; i < TM * 32 * S; i += N_CONSUMERS) tk[i] = KEY_INF;
    for (int i = tid; i < TM; i += N_CONSUMERS) thr_f[i] = INFINITY;

    float acc[C::RQ][C::RV];
    for (long long row0 = lo; row0 < hi; row0 += TN) {
        scan_consume_tile<TM, OP>(stages, full_bar, empty_bar, ps, nk, acc, q0, v0, lane);
        named_bar_sync(1, N_CONSUMERS);  // previous screening (and the state reset) is complete
        store_acc_to_dt<TM>(dt, acc, q0, v0, OP == OP_L2 ? 1.0f : -1.0f);
        named_bar_sync(1, N_CONSUMERS);
        const int n_valid = (int)((hi - row0) < TN ? (hi - row0) : TN);
        const uint32_t pos0 = (uint32_t)(row0 - lo);
        for (int r = warp; r < it.q_count; r += N_CONSUMER_WARPS) {
            const float4 v = *reinterpret_cast<const float4*>(dt + r * DT_LD + lane * 4);
            const float th = thr_f[r];
            const int cb = lane * 4;
            const float sc[4] = {v.x, v.y, v.z, v.w};
            uint32_t m[4];
            uint32_t any = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                m[j] = __ballot_sync(0xffffffffu, (cb + j < n_valid) && (sc[j] <= th));
                any |= m[j];
            }
            if (any) {
                unsigned long long key[S];
#pragma unroll
                for (int s = 0; s < S; ++s) key[s] = tk[(r * S + s) * 32 + lane];
                unsigned long long kth = warp_sorted_get<S>(key, k - 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t mm = m[j];
                    while (mm) {
                        const int src = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const float xs = __shfl_sync(0xffffffffu, sc[j], src);
                        const unsigned long long x = make_key(xs, pos0 + src * 4 + j);
                        if (x < kth) {
                            warp_sorted_insert<S>(key, x, lane);
                            kth = warp_sorted_get<S>(key, k - 1);
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < S; ++s) tk[(r * S + s) * 32 + lane] = key[s];
                if (lane == 0) thr_f[r] = (kth == KEY_INF) ? INFINITY : key_score(kth);
            }
        }
    }
    named_bar_sync(1, N_CONSUMERS);  // all screening of this item done, lists final
    // write the item's partial results: slot = q_begin + r, k sorted keys each
    for (int r = warp; r < it.q_count; r += N_CONSUMER_WARPS) {
        unsigned long long* out = p.part_key + (size_t)(it.q_begin + r) * k;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int e = s * 32 + lane;
            if (e < k) {
                unsigned long long x = tk[(r * S + s) * 32 + lane];
                if (x != KEY_INF && !p.store_local) {
                    const int gid = __ldg(p.list_ids + lo + key_pos(x));
                    x = (x & 0xFFFFFFFF00000000ull) | (uint32_t)gid;
                }
                out[e] = x;
            }
        }
    }
    named_bar_sync(1, N_CONSUMERS);  // tk / thr_f may be reset by the next item
}

template <int OP, int S>
__global__ void __launch_bounds__(N_THREADS, 1)
scan_lists_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
    // dynamic shared memory is the only shared allocation of this kernel: the declared alignment holds
    // (the 128-byte TMA swizzle needs 1024-byte aligned stages; checked below). Plain pointer arithmetic on
    // the array keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST).
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stages = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    float* dt = (float*)(stages + (size_t)SCAN_NSTAGE * SCAN_STAGE_BYTES);
    unsigned long long* tk = (unsigned long long*)(dt + SCAN_TM_MAX * DT_LD);
    float* thr_f = (float*)(tk + SCAN_TM_MAX * 32 * S);
    uint64_t* full_bar = (uint64_t*)(thr_f + SCAN_TM_MAX);
    uint64_t* empty_bar = full_bar + SCAN_NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pipe_init(full_bar, empty_bar, SCAN_NSTAGE, threadIdx.x);
    if (threadIdx.x == 32) tma_prefetch_desc(&tmap);
    __syncthreads();

    PipeState ps{0, 0};
    const int n_items = *p.n_items;
    if (warp == N_CONSUMER_WARPS) {
        for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
            const ScanItem it = p.items[i];
            switch (it.tm) {
                case 64: scan_item_producer<64>(p, it, &tmap, stages, full_bar, empty_bar, ps, lane); break;
                case 32: scan_item_producer<32>(p, it, &tmap, stages, full_bar, empty_bar, ps, lane); break;
                case 16: scan_item_producer<16>(p, it, &tmap, stages, full_bar, empty_bar, ps, lane); break;
                default: scan_item_producer<8>(p, it, &tmap, stages, full_bar, empty_bar, ps, lane); break;
            }
        }
    } else {
        for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
            const ScanItem it = p.items[i];
            switch (it.tm) {
                case 64: scan_item_consumer<64, OP, S>(p, it, stages, full_bar, empty_bar, ps, dt, tk, thr_f, warp, lane); break;
                case 32: scan_item_consumer<32, OP, S>(p, it, stages, full_bar, empty_bar, ps, dt, tk, thr_f, warp, lane); break;
                case 16: scan_item_consumer<16, OP, S>(p, it, stages, full_bar, empty_bar, ps, dt, tk, thr_f, warp, lane); break;
                default: scan_item_consumer<8, OP, S>(p, it, stages, full_bar, empty_bar, ps, dt, tk, thr_f, warp, lane); break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: per-query merge of the partial lists of its probed (list) slots, with id de-duplication.
//   dedup = 1: an id present in several probed lists counts once BEFORE selection
//              (north-star semantics == Python recall, LIRA_smallscale.py:210-214)
//   dedup = 0: select k (score, id) pairs first, then collapse equal ids (search.cpp:499-513)
// One warp per query. Also used for the cross-GPU merge (slots = ranks).
// ------------------------------------------------------------------------------------------
struct MergeParams {
    const unsigned long long* part_key;  // [P, k]
    const long long* probe_offsets;      // [Q+1]
    const int* probe_slot;               // [P] slot of the j-th probe of a query
    int k;
    int Q;
    int dedup;
    float* out_dist;                     // [Q, k] metric value (L2sq or IP)
    long long* out_ids;                  // [Q, k]
    int is_ip;
    const int* mask;                     // optional [Q]: only queries with mask[q] != 0 are written
    int R = 0;                           // > 0: cross-GPU merge, part_key is [R, Q, k] rank-major and probe_offsets /
                                         //      probe_slot are unused (slot of rank r = r * Q + q)
};

template <int S>
__global__ void __launch_bounds__(256) merge_topk_kernel(const MergeParams p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    if (p.mask && !p.mask[q]) return;
    const int k = p.k;
    unsigned long long key[S];
#pragma unroll
    for (int s = 0; s < S; ++s) key[s] = KEY_INF;
    unsigned long long kth = KEY_INF;
    const long long j_lo = p.R > 0 ? 0 : p.probe_offsets[q], j_hi = p.R > 0 ? p.R : p.probe_offsets[q + 1];
    for (long long j = j_lo; j < j_hi; ++j) {
        const long long slot = p.R > 0 ? j * p.Q + q : p.probe_slot[j];
        if (slot < 0) continue;  // invalid probe
        const unsigned long long* src = p.part_key + (size_t)slot * k;
        for (int e0 = 0; e0 < k; e0 += 32) {
            const int e = e0 + lane;
            const unsigned long long x = (e < k) ? src[e] : KEY_INF;
            uint32_t mm = __ballot_sync(0xffffffffu, x < kth);
            // source is sorted by score: once every lane's score exceeds the k-th score nothing further can enter
            if (!__ballot_sync(0xffffffffu, (x >> 32) <= (kth >> 32))) break;
            while (mm) {
                const int sl = __ffs(mm) - 1;
                mm &= mm - 1;
                const unsigned long long y = shfl_u64(x, sl);
                if (!(y < kth)) continue;
                bool dup = false;
                if (p.dedup) {
                    bool mine = false;
#pragma unroll
                    for (int s = 0; s < S; ++s) mine |= (key[s] == y);
                    dup = __any_sync(0xffffffffu, mine);
                }
                if (!dup) {
                    warp_sorted_insert<S>(key, y, lane);
                    kth = warp_sorted_get<S>(key, k - 1);
                }
            }
        }
    }
    // emit. dedup == 0: identical keys (same id => same score) are adjacent; keep the first of a run.
    bool valid[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int e = s * 32 + lane;
        unsigned long long prev = shfl_up_u64(key[s], 1);
        if (s > 0) {
            const unsigned long long carry = shfl_u64(key[s - 1], 31);
            if (lane == 0) prev = carry;
        }
        const bool is_first = (e == 0) || (prev != key[s]);
        valid[s] = (e < k) && (key[s] != KEY_INF) && (p.dedup || is_first);
    }
    int base = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t vm = __ballot_sync(0xffffffffu, valid[s]);
        if (valid[s]) {
            const int o = base + __popc(vm & ((1u << lane) - 1u));
            const float sc = key_score(key[s]);
            p.out_dist[(size_t)q * k + o] = p.is_ip ? -sc : sc;
            p.out_ids[(size_t)q * k + o] = (long long)(int)key_pos(key[s]);
        }
        base += __popc(vm);
    }
    for (int o = base + lane; o < k; o += 32) {
        p.out_dist[(size_t)q * k + o] = p.is_ip ? -INFINITY : INFINITY;
        p.out_ids[(size_t)q * k + o] = -1;
    }
}

}  // namespace lira
