// K2-U8: the grouped list scan on the integer tensor cores (tcgen05.mma kind::i8, int32 accumulators in TMEM) for
// byte-valued data -- every stored value and every query value an integer in [0, 255] (SIFT / BigANN-style), d <= 128,
// |x|^2 <= 4 112 895. The list rows stream as ONE byte per component (a quarter of the fp32 bytes the reference scans at
// search.cpp:468-514, half of the fp16 shadow copy of tc_scan_kernels.cuh) plus 64 bytes of norm digits, the tensor pipe
// runs K = 32 per instruction (twice the fp16 rate), and all arithmetic is integer, so distances equal the fp32
// direct-difference distances bit for bit.
//
//   work item  (list b, a segment of at most `seg_rows` of its rows, up to 384 queries of b's group) = up to THREE query
//              tiles of 128 rows (16 KiB each) that share every staged chunk of the list: the L2 -> shared-memory operand
//              stream is paid once per 384 queries instead of once per 128. Long lists are cut into row segments so that no
//              item exceeds a small part of an SM's share.
//   unit       one (query tile, chunk of 256 list rows) into one of the two 256-column accumulators: TWO augmented MMAs that
//              put -floor(|v|^2 / 2) into the accumulator -- A side: the constant signed row (-128 x 63, -1), B side: 64
//              unsigned digits of the row with  sum 128 d_j + l = floor(|v|^2 / 2)  (8-bit digits carry at most 128 x 255 per
//              K step, hence two K = 32 blocks) -- then 4 MMAs of M = 128, N = 256, K = 32 per 128 bytes of d. The
//              accumulator holds  a = q.v - floor(|v|^2 / 2),  so  2 a  is the score-side value u = 2 q.v - |v|^2  up to the
//              parity of |v|^2, and the epilogue needs NO per-column arithmetic before its max tree (a per-column
//              subtraction made the epilogue, not the tensor pipe, the limit of the first version: 17 % tensor pipe active).
//   epilogue   16 warps; a thread owns one query row and 64 of the 256 columns: 3-input-max tree over the raw accumulator,
//              ONE compare + vote per 32 columns (conservative by the parity bit); entries that pass are scored exactly
//              (2 a - (|v|^2 & 1), the parity read from global memory) in the out-of-line append.
//   seed pass  (SEED = true) every thread keeps the 4 largest of its 16-column sub-group maxima over the item's rows
//              (4-column blocks in the first chunk of a short list): 16 distinct entries per (query, list segment) row,
//              whose k-th largest (as 2 a - 1 <= u) bounds the query's final k-th best score (k <= 16); the tightest bound
//              over a query's lists is T[q] (atomicMin). Exhaustive probe sets pool the values of several lists instead.
//   filter     (SEED = false) entries with score <= T[q] are appended to a candidate region (exhaustive probe sets: one
//              private region per (pair, column part); otherwise one per pair with atomic slots); refine_topk_kernel turns
//              the regions into the top k over distinct ids. A region that overflows flags the query for the exact path.
#pragma once
#include "tc_scan_kernels.cuh"

namespace lira {

static constexpr int U8_M = 128;                    // queries per tile (UMMA M)
static constexpr int U8_NS = 256;                   // list rows per chunk (UMMA N, accumulator columns)
static constexpr int U8_KB = 128;                   // bytes (= components) per row of the data operand: one 128-byte swizzle row
static constexpr int U8_KBLK_BYTES = 128 * U8_KB;   // 16 KiB: a box of 128 rows
static constexpr int U8_AUG = 64;                   // norm digits per row (two K = 32 blocks, 64-byte swizzle rows)
static constexpr int U8_AUG_BOX = 128 * U8_AUG;     // 8 KiB: the digits of a box of 128 rows
static constexpr int U8_SLOT_BYTES = 2 * U8_KBLK_BYTES + 2 * U8_AUG_BOX;   // 48 KiB: a chunk (256 rows) and its digits
static constexpr int U8_NT = 3;                     // query tiles per work item at most = buffers of the A ring
static constexpr int U8_ITEM_Q = U8_NT * U8_M;      // queries per work item at most
static constexpr int U8_NSLOT_MAX = 3;              // B ring: 3 slots in the filter pass, 2 in the seed pass (which needs the exchange area)
static constexpr int U8_NQ = 2;                     // work-item queue depth (the scheduler warp runs this far ahead)
static constexpr int U8_PARTS = 4;                  // column parts = epilogue warps per TMEM lane quadrant
static constexpr int U8_EPI_WARPS = 4 * U8_PARTS;
static constexpr int U8_THREADS = (U8_EPI_WARPS + 3) * 32;
static constexpr int U8_MAX_D = U8_KB;              // d <= 128
static constexpr int U8_MAX_NORM = 128 * (63 * 255) * 2 + 255;   // |v|^2 <= 4 112 895: floor(|v|^2 / 2) = 128 H + l, H <= 63 x 255, l <= 127
static constexpr int U8_MASKED = -(3 << 28);         // accumulator value given to columns past the end of a segment: below every real value
                                                    //   (>= -2^21) and every bound, and 2 x it still fits an int32
static constexpr int U8_SEG_ROWS = 4096;            // rows of a list per work item (16 chunks x up to 3 tiles = 48 units at most)
__host__ __device__ constexpr int u8_nslot(bool seed) { return seed ? 2 : 3; }
__host__ __device__ constexpr size_t u8_smem_bytes(bool seed) {
    return (size_t)U8_NT * U8_KBLK_BYTES + (size_t)u8_nslot(seed) * U8_SLOT_BYTES     // operands
           + (size_t)U8_AUG_BOX                                                       // the constant A side of the augmented MMAs
           + 512                                                                      // barriers, item queue, tmem slot
           + (size_t)U8_NQ * U8_ITEM_Q * 12                                           // per queued item and row: query id, |q|^2, bound
           + (seed ? (size_t)U8_PARTS * U8_ITEM_Q * 16                                // seed pass: 4 values per (row, part)
                   : (size_t)U8_PARTS * U8_ITEM_Q * 4);                               // filter pass: private region fill per (row, part)
}

// D = S32, A = B = unsigned 8 bit (the augmented MMAs: A signed), both K-major, M = 128, N = 256 / 128
static constexpr uint32_t U8_IDESC_N256 = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(U8_M >> 4) << 24);
static constexpr uint32_t U8_IDESC_N128 = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(U8_M >> 4) << 24);
static constexpr uint32_t U8_IDESC_AUG = 1u << 7;   // a_format = signed 8 bit

__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 64-byte swizzle: rows of 64 B (two K = 32 steps), 8-row groups 512 B apart (SBO)
__device__ __forceinline__ uint64_t tc_smem_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

struct U8Params {
    const int* group_queries;        // [P] query id per slot
    const long long* list_offsets;   // [B+1]
    const ScanItem* items;           // items of up to 384 queries; ScanItem::tm = row segment of the list
    const int* n_items;
    int* work_counter;               // zeroed before launch: dynamic item scheduler
    int d8;                          // padded row length in bytes (multiple of 16, <= 128)
    int seg_rows;                    // rows per segment (multiple of 256)
    int q_mod;                       // > 0: exhaustive probe sets -- every list's group is the whole batch in query order, so the A tiles
                                     //   are read from ONE [Q, d8] copy of the queries (row = slot % Q) instead of one copy per list
    const int* nv;                   // [E + 256] -|v|^2 per list entry (L2) or 0 (IP): the append reads its parity
    const float* qnorm;              // [Q] |q|^2 (exact integers)
    uint32_t* thr;                   // [Q] bound T[q] on the k-th best score as f32_to_ordered(T): written by the seed pass, read by the filter
    unsigned long long* cand_key;    // [P, cap] (score, list entry) keys: one region per (query, list) pair
    int* cand_count;                 // [P] zeroed before the filter pass (private_regions: [U8_PARTS P], written by the owners)
    int private_regions;             // exhaustive probe sets (one work item per pair): one region per (pair, column part) filled by its
                                     //   owner thread, no atomics (their round trip is what a pass with hundreds of survivors per query waits for)
    int* seed_out;                   // seed pass, exhaustive probe sets: [P, U8_PARTS, 4] the 16 best scores of every pair go here instead
                                     //   of being turned into a bound in the kernel (u8_seed_select_kernel pools them); null otherwise
    int cap;
    int k;
    int is_ip;
    long long* trace;                // LIRA_U8_TRACE: [6][512] SM clock stamps of CTA 0's units (MMA: accumulator free, operands ready, issued;
                                     //   epilogue warp 0: accumulator ready, loaded, done), or null
    int* dbg;                        // LIRA_TC_EXP bit 6: {groups with a survivor, groups, items, tiles, chunks, units} counters
    int exp;                         // experiments (LIRA_TC_EXP, wrong results, timing only): bit 0 = skip the survivor path,
                                     //   bit 1 = skip the whole epilogue arithmetic, bit 2 = skip the MMAs
};

// inserts x into the descending list a[0] >= a[1] >= a[2] >= a[3] (the 4 largest values seen)
__device__ __forceinline__ void u8_top4_insert(int (&a)[4], int x) {
    const int n0 = max(a[0], x), x1 = min(a[0], x);
    const int n1 = max(a[1], x1), x2 = min(a[1], x1);
    const int n2 = max(a[2], x2), x3 = min(a[2], x2);
    a[0] = n0; a[1] = n1; a[2] = n2; a[3] = max(a[3], x3);
}

// filter pass, one block of 4 columns whose accumulator values a0..a3 may pass: u = 2 a - parity(|v|^2) exactly (nvp points at
// the block's -|v|^2, null for the inner product where u = a), and the entries with u >= lim are appended as
// (score = |q|^2 - u, list entry) keys. Out of line on purpose: see the note on code size in the epilogue.
// Shared region of the pair, slot taken with an atomic: returns true when the region overflowed (the query is redone exactly).
__device__ __noinline__ bool u8_append4(int a0, int a1, int a2, int a3, const int* nvp, int lim, int qn, uint32_t e0, unsigned long long* cand,
                                       int* cnt_ptr, int cap) {
    const int u0 = nvp ? 2 * a0 - (nvp[0] & 1) : a0, u1 = nvp ? 2 * a1 - (nvp[1] & 1) : a1;
    const int u2 = nvp ? 2 * a2 - (nvp[2] & 1) : a2, u3 = nvp ? 2 * a3 - (nvp[3] & 1) : a3;
    const int n = (u0 >= lim) + (u1 >= lim) + (u2 >= lim) + (u3 >= lim);
    if (n == 0) return false;
    int pos = atomicAdd(cnt_ptr, n);
    if (u0 >= lim) { if (pos < cap) cand[pos] = make_key((float)(qn - u0), e0); ++pos; }
    if (u1 >= lim) { if (pos < cap) cand[pos] = make_key((float)(qn - u1), e0 + 1); ++pos; }
    if (u2 >= lim) { if (pos < cap) cand[pos] = make_key((float)(qn - u2), e0 + 2); ++pos; }
    if (u3 >= lim) { if (pos < cap) cand[pos] = make_key((float)(qn - u3), e0 + 3); ++pos; }
    return pos > cap;
}
// the same for a region owned by the calling thread: returns the new fill (it may run past cap)
__device__ __noinline__ int u8_append4_private(int a0, int a1, int a2, int a3, const int* nvp, int lim, int qn, uint32_t e0,
                                              unsigned long long* cand, int cnt, int cap) {
    const int u0 = nvp ? 2 * a0 - (nvp[0] & 1) : a0, u1 = nvp ? 2 * a1 - (nvp[1] & 1) : a1;
    const int u2 = nvp ? 2 * a2 - (nvp[2] & 1) : a2, u3 = nvp ? 2 * a3 - (nvp[3] & 1) : a3;
    if (u0 >= lim) { if (cnt < cap) cand[cnt] = make_key((float)(qn - u0), e0); ++cnt; }
    if (u1 >= lim) { if (cnt < cap) cand[cnt] = make_key((float)(qn - u1), e0 + 1); ++cnt; }
    if (u2 >= lim) { if (cnt < cap) cand[cnt] = make_key((float)(qn - u2), e0 + 2); ++cnt; }
    if (u3 >= lim) { if (cnt < cap) cand[cnt] = make_key((float)(qn - u3), e0 + 3); ++cnt; }
    return cnt;
}

template <bool SEED, bool IP>
__global__ void __launch_bounds__(U8_THREADS, 1)   // (96 registers per thread: 19 warps x 104 no longer fit the register file)
u8_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_v,
               const __grid_constant__ CUtensorMap tmap_vaug, const __grid_constant__ CUtensorMap tmap_aaug, const U8Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    constexpr int NSLOT = u8_nslot(SEED);
    uint8_t* sA = smem_raw;                                             // [A ring: U8_NT tiles of 128 x 128 B]
    uint8_t* sB = sA + (size_t)U8_NT * U8_KBLK_BYTES;                   // [NSLOT]{[256 x 128 B] rows, [256 x 64 B] norm digits}
    uint8_t* sGA = sB + (size_t)NSLOT * U8_SLOT_BYTES;                  // [128 x 64 B] constant A side of the augmented MMAs
    uint64_t* bars = (uint64_t*)(sGA + U8_AUG_BOX);
    uint64_t* a_full = bars;                        // [U8_NT] per tile buffer
    uint64_t* a_empty = a_full + U8_NT;             // [U8_NT]
    uint64_t* b_full = a_empty + U8_NT;             // [U8_NSLOT_MAX]
    uint64_t* b_empty = b_full + U8_NSLOT_MAX;      // [U8_NSLOT_MAX]
    uint64_t* t_full = b_empty + U8_NSLOT_MAX;      // [2]
    uint64_t* t_empty = t_full + 2;                 // [2]
    uint64_t* i_full = t_empty + 2;                 // [U8_NQ]
    uint64_t* i_empty = i_full + U8_NQ;             // [U8_NQ]
    uint64_t* ga_full = i_empty + U8_NQ;            // [1]
    TcQItem* iq = (TcQItem*)(ga_full + 1);          // [U8_NQ]   (21 barriers = 168 B, 2 items = 64 B)
    uint32_t* tmem_slot = (uint32_t*)(iq + U8_NQ);
    int* s_q = (int*)((uint8_t*)bars + 512);                    // [U8_NQ][384] query id (-1: padding row)
    int* s_qn = s_q + U8_NQ * U8_ITEM_Q;                        // [U8_NQ][384] |q|^2 (0 for IP)
    int* s_lim = s_qn + U8_NQ * U8_ITEM_Q;                      // [U8_NQ][384] an entry survives iff u >= lim
    int4* s_x = reinterpret_cast<int4*>(s_lim + U8_NQ * U8_ITEM_Q);   // [U8_PARTS][384]: seed pass, the 4 largest values per (row, part)
    int* s_cnt = s_lim + U8_NQ * U8_ITEM_Q;                           // [U8_PARTS][384]: filter pass, fill of the private regions (same area)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_PROD = U8_EPI_WARPS, W_MMA = W_PROD + 1, W_ALLOC = W_PROD + 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < U8_NT; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < U8_NSLOT_MAX; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], U8_EPI_WARPS); }
        for (int i = 0; i < U8_NQ; ++i) { mbar_init(&i_full[i], 1); mbar_init(&i_empty[i], 2 + U8_EPI_WARPS); }   // producer + MMA + epilogue warps
        mbar_init(ga_full, 1);
        mbar_fence_init();
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == W_PROD * 32) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_vaug); tma_prefetch_desc(&tmap_aaug); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // (the prologue above touches nothing another kernel produces)
    pdl_launch();
    const int n_items = *p.n_items;

    if (warp == W_ALLOC) {
        // ===== scheduler (the TMEM allocator warp has nothing else to do): claims work items and stages what the other warps
        // need to start one -- row range, and per row the query id, |q|^2 and the bound -- in the shared-memory queue, U8_NQ
        // items ahead. Its chain of dependent global round trips (ticket -> descriptor -> list range + row ids -> per-row
        // gathers, ~2-3 us) therefore never sits between two items of the pipeline, however short they are.
        // The chain is software-pipelined over FOUR items: every iteration issues the ticket of item n+3, the descriptor load of
        // item n+2, the row-range / row-id loads of item n+1 and the per-row gathers of item n -- four independent round trips in
        // flight together -- and then stages item n. So the scheduler delivers one item per memory latency (< 1 us).
        auto ticket = [&]() { return lane == 0 ? atomicAdd(p.work_counter, 1) : 0; };
        auto descriptor = [&](int tk) {
            const int idx = __shfl_sync(0xffffffffu, tk, 0);
            ScanItem d;
            if (idx < n_items) d = p.items[idx];
            else { d.list = -1; d.q_begin = 0; d.q_count = 0; d.tm = 0; }
            return d;
        };
        constexpr int RPL = U8_ITEM_Q / 32;   // rows per lane (row = lane + 32 j)
        int tk = ticket();
        ScanItem d2 = descriptor(tk);          // item 0
        tk = ticket();
        ScanItem d2n = descriptor(tk);         // item 1
        tk = ticket();                         // item 2's ticket
        TcQItem o1;
        int rq1[RPL];
        auto rows_of = [&](const ScanItem& d, TcQItem& o, int (&rq)[RPL]) {
            o.list = d.list; o.q_begin = d.q_begin; o.q_count = d.q_count; o.pad = 0;
            o.lo = 0; o.hi = 0;
            if (d.list >= 0) {
                const long long l0 = p.list_offsets[d.list], l1 = p.list_offsets[d.list + 1];
                o.lo = min(l1, l0 + (long long)d.tm * p.seg_rows);
                o.hi = min(l1, o.lo + p.seg_rows);
            }
#pragma unroll
            for (int j = 0; j < RPL; ++j) rq[j] = (d.list >= 0 && lane + 32 * j < d.q_count) ? __ldg(p.group_queries + d.q_begin + lane + 32 * j) : -1;
        };
        rows_of(d2, o1, rq1);                  // rows of item 0
        d2 = d2n;                              // d2 = descriptor of item 1
        for (int n = 0;; ++n) {
            // ---- issue: gathers of item n, rows of item n+1, descriptor of item n+2, ticket of item n+3 ----
            const TcQItem o0 = o1;
            int rq0[RPL], rqn[RPL], rlim[RPL];
#pragma unroll
            for (int j = 0; j < RPL; ++j) rq0[j] = rq1[j];
            float qn_f[RPL];
            uint32_t thr_u[RPL];
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                qn_f[j] = (!IP && rq0[j] >= 0) ? __ldg(p.qnorm + rq0[j]) : 0.f;
                thr_u[j] = (!SEED && rq0[j] >= 0) ? __ldg(p.thr + rq0[j]) : 0xFF800000u;
            }
            if (o0.list >= 0) {
                rows_of(d2, o1, rq1);
                d2 = descriptor(tk);
                tk = ticket();
            }
            // ---- consume: stage item n ----
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const int qn = (int)qn_f[j];
                rqn[j] = qn;
                rlim[j] = 0x7FFFFFFF;   // padding rows: nothing ever passes
                if (!SEED && rq0[j] >= 0) {
                    // score = qn - u <= T  <=>  u >= qn - T; T = +inf (no bound): everything real passes (u > -2^26)
                    const float T = ordered_to_f32(thr_u[j]);
                    const int Ti = T >= 1073741824.f ? 1073741824 : (T <= -1073741824.f ? -1073741824 : (int)floorf(T));
                    rlim[j] = qn - Ti;
                }
            }
            const int qs = n % U8_NQ;
            mbar_wait(&i_empty[qs], ((n / U8_NQ) & 1) ^ 1u);
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                s_q[qs * U8_ITEM_Q + lane + 32 * j] = rq0[j];
                s_qn[qs * U8_ITEM_Q + lane + 32 * j] = rqn[j];
                s_lim[qs * U8_ITEM_Q + lane + 32 * j] = rlim[j];
            }
            __syncwarp();   // every lane's rows are written before lane 0 publishes the item (arrive = release)
            if (lane == 0) { iq[qs] = o0; mbar_arrive(&i_full[qs]); }
            __syncwarp();
            if (o0.list < 0) break;
        }
    } else if (warp == W_PROD) {
        // ===== TMA producer (whole warp in the loop, one elected lane issues) =====
        if (!IP) {   // the constant A side of the augmented MMAs: every row is (-128 x 63, -1)
            if (elect_one()) {
                mbar_arrive_expect_tx(ga_full, U8_AUG_BOX);
                tma_load_2d(sGA, &tmap_aaug, 0, 0, ga_full);
            }
            __syncwarp();
        }
        PipeState bs{0, 0};
        uint32_t ga = 0;   // running tile counter -> A ring buffer and phase
        for (int n = 0;; ++n) {
            const int qs = n % U8_NQ;
            mbar_wait(&i_full[qs], (n / U8_NQ) & 1);
            const TcQItem it = iq[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&i_empty[qs]);
            if (it.list < 0) break;
            const int ntile = (it.q_count + U8_M - 1) / U8_M;
            for (int t = 0; t < ntile; ++t, ++ga) {   // the item's query tiles into the next buffers of the A ring
                const uint32_t x = ga % (uint32_t)U8_NT;
                mbar_wait(&a_empty[x], ((ga / (uint32_t)U8_NT) & 1) ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[x], U8_KBLK_BYTES);
                    tma_load_2d(sA + (size_t)x * U8_KBLK_BYTES, &tmap_q, 0, (p.q_mod > 0 ? it.q_begin % p.q_mod : it.q_begin) + t * U8_M, &a_full[x]);
                }
                __syncwarp();
            }
            const long long lo = it.lo, hi = it.hi;
            for (long long row0 = lo; row0 < hi; row0 += U8_NS) {
                const int nh = hi - row0 > 128 ? 2 : 1;   // boxes of 128 rows (the tail of a list may need one only)
                mbar_wait(&b_empty[bs.stage], bs.phase ^ 1u);
                if (elect_one()) {
                    uint8_t* slot = sB + (size_t)bs.stage * U8_SLOT_BYTES;
                    mbar_arrive_expect_tx(&b_full[bs.stage], (uint32_t)nh * (U8_KBLK_BYTES + (IP ? 0 : U8_AUG_BOX)));
                    for (int x = 0; x < nh; ++x) {
                        tma_load_2d(slot + (size_t)x * U8_KBLK_BYTES, &tmap_v, 0, (int)row0 + x * 128, &b_full[bs.stage]);
                        if (!IP) tma_load_2d(slot + 2 * U8_KBLK_BYTES + (size_t)x * U8_AUG_BOX, &tmap_vaug, 0, (int)row0 + x * 128, &b_full[bs.stage]);
                    }
                }
                __syncwarp();
                bs.advance(NSLOT);
            }
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer =====
        PipeState bs{0, 0};
        uint32_t m = 0;   // running unit counter -> accumulator and phase
        uint32_t gm = 0;  // running tile counter -> A ring buffer and phase
        if (!IP) { mbar_wait(ga_full, 0); tc_fence_after(); }
        const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB), sGA_u32 = smem_u32(sGA);
        const int ksteps = min(4, (p.d8 + 31) / 32);
        for (int n = 0;; ++n) {
            const int qs = n % U8_NQ;
            mbar_wait(&i_full[qs], (n / U8_NQ) & 1);
            const TcQItem it = iq[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&i_empty[qs]);
            if (it.list < 0) break;
            const int ntile = (it.q_count + U8_M - 1) / U8_M;
            if ((p.exp & 64) && lane == 0) {   // {.., .., items, tiles, chunks, units}
                const int nch = (int)((it.hi - it.lo + U8_NS - 1) / U8_NS);
                atomicAdd(p.dbg + 2, 1); atomicAdd(p.dbg + 3, ntile); atomicAdd(p.dbg + 4, nch); atomicAdd(p.dbg + 5, nch * ntile);
            }
            for (int t = 0; t < ntile; ++t) {
                const uint32_t g = gm + t;
                mbar_wait(&a_full[g % (uint32_t)U8_NT], (g / (uint32_t)U8_NT) & 1);
            }
            tc_fence_after();
            const long long lo = it.lo, hi = it.hi;
            for (long long row0 = lo; row0 < hi; row0 += U8_NS) {
                const uint32_t nsel = hi - row0 > 128 ? (uint32_t)(256 >> 3) << 17 : (uint32_t)(128 >> 3) << 17;   // N = 256 (tail of a list: 128)
                const uint32_t idesc = (2u << 4) | nsel | ((uint32_t)(U8_M >> 4) << 24);
                const uint32_t b_addr = sB_u32 + (uint32_t)bs.stage * U8_SLOT_BYTES;
                for (int t = 0; t < ntile; ++t, ++m) {
                    const uint32_t acc = m & 1u;
                    mbar_wait(&t_empty[acc], ((m >> 1) & 1) ^ 1u);
                    tc_fence_after();
                    if (p.trace && blockIdx.x == 0 && m < 512 && lane == 0) p.trace[0 * 512 + m] = clock64();
                    if (t == 0) { mbar_wait(&b_full[bs.stage], bs.phase); tc_fence_after(); }
                    if (p.trace && blockIdx.x == 0 && m < 512 && lane == 0) p.trace[1 * 512 + m] = clock64();
                    const uint32_t d_tmem = tmem_base + acc * U8_NS;
                    const uint32_t a_addr = sA_u32 + (uint32_t)((gm + t) % (uint32_t)U8_NT) * U8_KBLK_BYTES;
                    if (elect_one() && !(p.exp & 4)) {
                        if (!IP) {   // a = -floor(|v|^2 / 2) ...
                            tc_mma_i8(d_tmem, tc_smem_desc_sw64(sGA_u32), tc_smem_desc_sw64(b_addr + 2 * U8_KBLK_BYTES), idesc | U8_IDESC_AUG, 0u);
                            tc_mma_i8(d_tmem, tc_smem_desc_sw64(sGA_u32 + 32), tc_smem_desc_sw64(b_addr + 2 * U8_KBLK_BYTES + 32), idesc | U8_IDESC_AUG, 1u);
                        }
                        for (int j = 0; j < ksteps; ++j)   // ... + q.v, K = 32 bytes inside the 128-byte swizzle row
                            tc_mma_i8(d_tmem, tc_smem_desc(a_addr + j * 32), tc_smem_desc(b_addr + j * 32), idesc, (!IP || j) ? 1u : 0u);
                    }
                    __syncwarp();
                    if (elect_one()) tc_commit(&t_full[acc]);
                    __syncwarp();
                    if (p.trace && blockIdx.x == 0 && m < 512 && lane == 0) p.trace[2 * 512 + m] = clock64();
                }
                if (elect_one()) tc_commit(&b_empty[bs.stage]);   // the chunk's slot is free once every MMA above has read it
                __syncwarp();
                bs.advance(NSLOT);
            }
            for (int t = 0; t < ntile; ++t, ++gm) {   // ... and so are the item's query tiles
                if (elect_one()) tc_commit(&a_empty[gm % (uint32_t)U8_NT]);
                __syncwarp();
            }
        }
    } else if (warp < U8_EPI_WARPS) {
        // ===== epilogue =====
        // (code size matters here: with the tile / half loops unrolled and the append code inlined per column the kernel no
        //  longer fits the instruction cache and every survivor costs microseconds -- measured; so the loops below are real
        //  loops, per-tile state lives in shared memory, and the append is one out-of-line function)
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int part = warp >> 2;                // which 32 columns of each 128-column half
        const int cap = p.cap;
        const uint32_t t_full_u32 = smem_u32(t_full), t_empty_u32 = smem_u32(t_empty);
        uint32_t m = 0;
        for (int n = 0;; ++n) {
            const int qs = n % U8_NQ;
            mbar_wait(&i_full[qs], (n / U8_NQ) & 1);
            const TcQItem it = iq[qs];
            if (it.list < 0) break;
            const int* iq_q = s_q + qs * U8_ITEM_Q;
            const int* iq_qn = s_qn + qs * U8_ITEM_Q;
            int* iq_lim = s_lim + qs * U8_ITEM_Q;
            const int ntile = (it.q_count + U8_M - 1) / U8_M;
            const int n_rows = (int)(it.hi - it.lo);
            const int trow = quad * 32 + lane;     // this thread's row inside a tile
            if (SEED) {
                if (n > 0) named_bar_sync(2, U8_EPI_WARPS * 32);   // the previous item's values have been read
                for (int t = 0; t < ntile; ++t) s_x[part * U8_ITEM_Q + t * U8_M + trow] = make_int4(INT_MIN, INT_MIN, INT_MIN, INT_MIN);
            } else if (p.private_regions) {
                for (int t = 0; t < ntile; ++t) s_cnt[part * U8_ITEM_Q + t * U8_M + trow] = 0;   // (only this thread touches its entries)
            }
            uint32_t ebase = (uint32_t)it.lo + part * 32;
            int left = n_rows - part * 32;
            for (int rows_left = n_rows; rows_left > 0; rows_left -= U8_NS, ebase += U8_NS, left -= U8_NS) {
                const int nh = rows_left > 128 ? 2 : 1;
                const bool fine = SEED && rows_left == n_rows && n_rows <= 1024;   // first chunk of a short list: 4-column blocks
#pragma unroll 1
                for (int t = 0; t < ntile; ++t, ++m) {
                    const uint32_t acc = m & 1u;
                    const uint32_t par = (m >> 1) & 1u;
                    mbar_wait_addr(t_full_u32 + acc * 8, par);
                    if (p.trace && blockIdx.x == 0 && m < 512 && threadIdx.x == 0) p.trace[3 * 512 + m] = clock64();
                    // a warp whose 32 rows of this tile are all padding only keeps the accumulator handshake going
                    if (t * U8_M + quad * 32 >= it.q_count) {
                        if (lane == 0) mbar_arrive_addr(t_empty_u32 + acc * 8);
                        __syncwarp();
                        continue;
                    }
                    tc_fence_after();
                    const int row = t * U8_M + trow;
                    // the accumulator value a relates to u by  u = 2 a - parity  (L2) or u = a (IP): a passes (conservatively) iff
                    // 2 a >= lim  <=>  a >= ceil(lim / 2)
                    const int lim = SEED ? 0 : iq_lim[row];
                    const int lima = IP ? lim : (lim >> 1) + (lim & 1);   // ceil(lim / 2) for either sign (arithmetic shift = floor)
                    int a[4];
                    if (SEED) { const int4 v = s_x[part * U8_ITEM_Q + row]; a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * U8_NS + part * 32;
#pragma unroll 1
                    for (int h = 0; h < nh; ++h) {   // the two 128-column halves of the accumulator, one after the other
                        uint32_t ru[32];
                        tc_ld32_async(taddr + h * 128, ru);
                        tc_ld_wait(ru);
                        if (h == nh - 1) {   // the thread's last columns of this unit are in registers: hand the accumulator back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_addr(t_empty_u32 + acc * 8);
                            if (p.trace && blockIdx.x == 0 && m < 512 && threadIdx.x == 0) p.trace[4 * 512 + m] = clock64();
                        }
                        if (p.exp & 2) continue;
                        int r[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = (int)ru[i];
                        // columns past the end of the segment hold other rows
                        const int n_valid = left - h * 128;
                        if (n_valid < 32) {
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                if (c >= n_valid) r[c] = U8_MASKED;
                        }
                        int m4[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) m4[i] = max(max(r[4 * i], r[4 * i + 1]), max(r[4 * i + 2], r[4 * i + 3]));
                        const int m16a = max(max(m4[0], m4[1]), max(m4[2], m4[3]));
                        const int m16b = max(max(m4[4], m4[5]), max(m4[6], m4[7]));
                        if (SEED) {
                            if (fine) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) u8_top4_insert(a, m4[i]);
                            } else { u8_top4_insert(a, m16a); u8_top4_insert(a, m16b); }
                        } else {
                            const int mx = max(m16a, m16b);
                            if ((p.exp & 64) && lane == 0) atomicAdd(p.dbg + 1, 1);
                            if (__any_sync(0xffffffffu, mx >= lima) && !(p.exp & 1)) {
                                if ((p.exp & 64) && lane == 0) atomicAdd(p.dbg, 1);
                                const size_t slot = (size_t)(it.q_begin + row);
                                const uint32_t eb = ebase + h * 128;
                                const int* nvb = IP ? nullptr : p.nv + eb;
                                const int qn = iq_qn[row];
                                bool over = false;
                                // which blocks of 4 columns hold a passing entry in SOME lane: one warp reduction, then warp-uniform
                                // branches (a vote per block would put eight dependent round trips on the warp's critical path)
                                uint32_t qm = 0;
#pragma unroll
                                for (int j = 0; j < 8; ++j) qm |= (m4[j] >= lima) ? (1u << j) : 0u;
                                const uint32_t um = __reduce_or_sync(0xffffffffu, qm);
                                if (p.private_regions) {
                                    unsigned long long* cand = p.cand_key + (slot * U8_PARTS + part) * cap;
                                    int cnt = s_cnt[part * U8_ITEM_Q + row];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        if (um & (1u << j))
                                            cnt = u8_append4_private(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], IP ? nullptr : nvb + 4 * j, lim, qn,
                                                                     eb + 4 * j, cand, cnt, cap);
                                    }
                                    s_cnt[part * U8_ITEM_Q + row] = cnt;
                                    over = cnt > cap;
                                } else {
                                    unsigned long long* cand = p.cand_key + slot * cap;
                                    int* cnt_ptr = p.cand_count + slot;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        if (um & (1u << j))
                                            over |= u8_append4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], IP ? nullptr : nvb + 4 * j, lim, qn,
                                                               eb + 4 * j, cand, cnt_ptr, cap);
                                    }
                                }
                                if (over) iq_lim[row] = 0x7FFFFFFF;   // the region overflowed (the query is redone exactly): stop collecting
                            }
                        }
                    }
                    if (SEED) s_x[part * U8_ITEM_Q + row] = make_int4(a[0], a[1], a[2], a[3]);
                    if (p.trace && blockIdx.x == 0 && m < 512 && threadIdx.x == 0) p.trace[5 * 512 + m] = clock64();
                }
            }
            if (SEED) {
                // a row's four column parts kept 4 accumulator values each: 16 distinct entries of this list segment, each with
                // u >= 2 a - 1. The k-th largest of these lower bounds has k entries at or above it  =>  T = |q|^2 - that value
                // bounds the query's final k-th best score (k <= 16).
                // Exhaustive probe sets (p.seed_out): the 16 score bounds go to global memory and are pooled over the query's lists.
                named_bar_sync(1, U8_EPI_WARPS * 32);
                const int row = warp * 32 + lane;   // 512 epilogue threads, one row each (384 rows at most)
                if (row < it.q_count) {
                    const int qn = iq_qn[row];
                    auto lower_u = [&](int av) { return av > U8_MASKED ? (IP ? av : 2 * av - 1) : INT_MIN; };
                    if (p.seed_out) {
#pragma unroll
                        for (int pp = 0; pp < U8_PARTS; ++pp) {
                            const int4 x = s_x[pp * U8_ITEM_Q + row];
                            int4 o;
                            o.x = x.x > U8_MASKED ? qn - lower_u(x.x) : INT_MAX; o.y = x.y > U8_MASKED ? qn - lower_u(x.y) : INT_MAX;
                            o.z = x.z > U8_MASKED ? qn - lower_u(x.z) : INT_MAX; o.w = x.w > U8_MASKED ? qn - lower_u(x.w) : INT_MAX;
                            *reinterpret_cast<int4*>(p.seed_out + ((size_t)(it.q_begin + row) * U8_PARTS + pp) * 4) = o;
                        }
                    } else {
                        float v[16];   // (descending order wanted: sort the negated values ascending)
#pragma unroll
                        for (int pp = 0; pp < U8_PARTS; ++pp) {
                            const int4 x = s_x[pp * U8_ITEM_Q + row];
                            v[4 * pp + 0] = -(float)lower_u(x.x); v[4 * pp + 1] = -(float)lower_u(x.y);
                            v[4 * pp + 2] = -(float)lower_u(x.z); v[4 * pp + 3] = -(float)lower_u(x.w);
                        }
                        tc_sort16(v);
                        const float tk = tc_pick16(v, p.k - 1);   // = -(k-th largest bound); 2^31 when fewer than k entries were seen
                        if (tk < 1073741824.f) atomicMin(p.thr + iq_q[row], f32_to_ordered((float)qn + tk));
                    }
                }
            }
            if (!SEED && p.private_regions) {
                for (int t = 0; t < ntile; ++t) {
                    const int row = t * U8_M + trow;
                    if (row < it.q_count) p.cand_count[(size_t)(it.q_begin + row) * U8_PARTS + part] = s_cnt[part * U8_ITEM_Q + row];
                }
            }
            // the item's per-row data has been read: its queue slot may be reused
            __syncwarp();
            if (lane == 0) mbar_arrive(&i_empty[qs]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// exhaustive probe sets (exact kNN over disjoint base segments; group slot of (query q, list b) = b Q + q): T[q] = k-th
// smallest of the 16 S scores the seed pass kept for the query's first S lists (S <= 32) -- all distinct entries. One warp per
// query: the values sit in registers (16 per lane) and the k-th smallest is found bit by bit (radix select on the
// order-preserving unsigned image of the int32 scores: 32 rounds of 16 compares + one warp sum).
__global__ void __launch_bounds__(256) u8_seed_select_kernel(const int* __restrict__ seed_out, int Q, int S, int k, uint32_t* __restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    uint32_t v[16];   // lane l holds the 16 values of list l (missing: 0xFFFFFFFF)
    if (lane < S) {
        const int4* src = reinterpret_cast<const int4*>(seed_out + ((size_t)lane * Q + q) * (U8_PARTS * 4));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int4 x = src[j];
            v[4 * j + 0] = (uint32_t)x.x ^ 0x80000000u; v[4 * j + 1] = (uint32_t)x.y ^ 0x80000000u;
            v[4 * j + 2] = (uint32_t)x.z ^ 0x80000000u; v[4 * j + 3] = (uint32_t)x.w ^ 0x80000000u;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0xFFFFFFFFu;
    }
    // (INT_MAX marks a missing value: its image 0xFFFFFFFF is the largest key)
    int valid = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) valid += v[j] != 0xFFFFFFFFu;
    valid = __reduce_add_sync(0xffffffffu, valid);
    if (valid < k) {
        if (lane == 0) thr[q] = 0xFF800000u;   // f32_to_ordered(+inf): fewer than k candidates seen
        return;
    }
    uint32_t res = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t t = res | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) c += v[j] < t;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c < k) res = t;   // fewer than k keys below t: the k-th smallest is at least t
    }
    if (lane == 0) thr[q] = f32_to_ordered((float)(int)(res ^ 0x80000000u));
}

// items of lists 0 .. S-1 (the seed pass of exhaustive probe sets runs on these only)
__global__ void u8_seed_items_kernel(const ScanItem* __restrict__ items, const int* __restrict__ n_items, int S,
                                     ScanItem* __restrict__ out, int* __restrict__ n_out) {
    const int n = *n_items;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const ScanItem it = items[i];
        if (it.list < S) out[atomicAdd(n_out, 1)] = it;
    }
}

// the norm digits of the list rows (L2): aug[n, 64] unsigned bytes with  128 (d_0 + .. + d_62) + d_63 = floor(|x|^2 / 2)
// (greedy: leading digits 255), the B side of the two augmented MMAs whose A side is the constant row (-128 x 63, -1)
__global__ void shadow_digits_u8_kernel(const float* __restrict__ norm, long long n, uint8_t* __restrict__ aug) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * 16; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> 4;
        const int g = (int)(i & 15);          // which 4 digits of the row
        const int nv2 = (int)norm[row] >> 1;  // floor(|x|^2 / 2)
        const int H = nv2 >> 7, L = nv2 & 127;
        uint32_t pk = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = 4 * g + c;
            int dj;
            if (j == 63) dj = L;
            else { const int rem = H - 255 * j; dj = rem <= 0 ? 0 : (rem >= 255 ? 255 : rem); }
            pk |= (uint32_t)dj << (8 * c);
        }
        *reinterpret_cast<uint32_t*>(aug + row * 64 + 4 * g) = pk;
    }
}

// the byte shadow copy of the list rows: x8[n, d8] = uint8(x) (zero padded) and nv[n] = -|x|^2 (L2) or 0 (IP) as int32
__global__ void shadow_rows_u8_kernel(const float* __restrict__ x, long ld, int d, long long n, const float* __restrict__ norm,
                                      int is_ip, uint8_t* __restrict__ x8, int d8, int* __restrict__ nv) {
    const int per_row = d8 / 4;
    const long long total = n * per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / per_row;
        const int c = (int)(i % per_row) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) v = *reinterpret_cast<const float4*>(x + row * ld + c);   // (ld and the padding up to round_up(d, 4) are zero filled)
        const uint32_t pk = (uint32_t)(int)v.x | ((uint32_t)(int)v.y << 8) | ((uint32_t)(int)v.z << 16) | ((uint32_t)(int)v.w << 24);
        *reinterpret_cast<uint32_t*>(x8 + row * d8 + c) = pk;
        if (c == 0) nv[row] = is_ip ? 0 : -(int)norm[row];
    }
}

// *flag is cleared unless every value of x[n, d] is an integer in [0, 255]
__global__ void check_u8_kernel(const float* __restrict__ x, long ld, int d, long long n, int* __restrict__ flag) {
    const int per_row = (d + 3) / 4;
    const long long total = n * per_row;
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / per_row;
        const int c = (int)(i % per_row) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + row * ld + c);
        bad |= !(v.x >= 0.f && v.x <= 255.f && v.y >= 0.f && v.y <= 255.f && v.z >= 0.f && v.z <= 255.f && v.w >= 0.f && v.w <= 255.f);
        bad |= (v.x != rintf(v.x)) | (v.y != rintf(v.y)) | (v.z != rintf(v.z)) | (v.w != rintf(v.w));
    }
    if (bad) *flag = 0;
}

// gq8[slot, :] = uint8(q[group_queries[slot], :]) (queries in group order, zero padded to d8 bytes); *bad_flag is set
// when a value is not an integer in [0, 255] (the batch then takes another path)
__global__ void gather_group_queries_u8_kernel(const float* __restrict__ q, long ldq, int ds, const int* __restrict__ group_queries,
                                               long long P, const long long* __restrict__ n_slots, uint8_t* __restrict__ gq8, int d8,
                                               int* __restrict__ ok_flag) {
    const int per_row = d8 / 4;
    const long long total = (*n_slots < P ? *n_slots : P) * per_row;
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / per_row;
        const int c = (int)(i % per_row) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < ds) v = *reinterpret_cast<const float4*>(q + (long long)group_queries[s] * ldq + c);
        bad |= !(v.x >= 0.f && v.x <= 255.f && v.y >= 0.f && v.y <= 255.f && v.z >= 0.f && v.z <= 255.f && v.w >= 0.f && v.w <= 255.f);
        const uint32_t pk = (uint32_t)(int)v.x | ((uint32_t)(int)v.y << 8) | ((uint32_t)(int)v.z << 16) | ((uint32_t)(int)v.w << 24);
        *reinterpret_cast<uint32_t*>(gq8 + s * d8 + c) = pk;
    }
    if (bad && ok_flag) *ok_flag = 0;
}

}  // namespace lira
