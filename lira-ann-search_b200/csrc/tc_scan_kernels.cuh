// K2-TC: the grouped list scan on the 5th-generation tensor cores (tcgen05 + TMEM), used for query batches
// of 256 and more, where the fp32 CUDA-core scan is compute-bound by an order of magnitude.
//
//   seed     tc_scan_kernel<true>: the same pipeline as the filter over the first rows of every probed list, on the
//            filter's own work items; per (query, list) row it keeps 64 group minima of the score, whose k-th smallest
//            bounds the query's final k-th best score (k real, distinct candidates of one list); the smallest bound
//            over a query's lists is T[q] (atomicMin). Exact kNN (exhaustive probe sets) seeds from two base segments.
//   filter   tc_scan_kernel<false>: for a tile of 128 queries x one inverted list, D = Q_tile . V^T is computed by
//            tcgen05.mma kind::f16 (M = 128, N = 256, K = 16) from the fp16 shadow copy of the list rows that TMA
//            staged in shared memory (128-byte swizzle, K-major); an augmented K block adds -|v|^2, so the
//            accumulator (2 x 256 TMEM columns: the epilogue of one chunk overlaps the MMAs of the next) holds
//            2 q.v - |v|^2. 16 epilogue warps read their query rows with tcgen05.ld, test 32 columns at a time
//            against T[q] - |q|^2 with a 3-input-max tree, append the survivors to private candidate regions and
//            tighten T[q] whenever a region fills up and is compacted to its k best.
//   refine   refine_topk_kernel: per query, candidate regions -> top-k over distinct ids (mode 2: after scoring every
//            candidate that can still make it again exactly in fp32 from the original rows).
//
// Mode 1 (exact): every stored value and every query value of the batch is an integer of at most 11 bits (exactly
// representable in fp16, e.g. SIFT / BigANN-style data) and |x|^2 < 2^22. Then every product and partial sum is an
// integer below 2^24, the tensor-core result (fp16 operands, fp32 accumulation) equals the fp32 direct-difference
// result bit for bit, and ids and distances are identical to the CUDA-core path.
// Mode 2 (approximate filter + exact re-rank, k <= 16): any other finite data; the shadow copy is fp16(sigma v), every
// bound carries a rigorous margin M(q) for the rounding of the operands, and the refine pass re-scores the survivors
// exactly, so the results agree with the CUDA-core path up to fp32 summation order.
// d <= 256: the query tile is resident in shared memory; 256 < d <= 1024: both operands stream (TC_STREAM_*).
#pragma once
#include <cuda_fp16.h>
#include "scan_kernels.cuh"

namespace lira {

static constexpr int TC_M = 128;         // queries per tile (UMMA M, TMEM lanes)
static constexpr int TC_N = 128;         // vectors per TMA box / per half of a chunk
static constexpr int TC_NS = 256;        // vectors per chunk (UMMA N, TMEM columns per accumulator): with N = 256 an MMA reads 12 KiB of
                                         //   shared memory per 128 tensor-pipe cycles instead of 8 KiB per 64 (N = 128 is bound by it)
static constexpr int TC_NH = TC_NS / TC_N;   // halves (TMA boxes of 128 rows) per chunk
static constexpr int TC_NACC = 2;        // TMEM accumulators in flight (2 x 256 = 512 columns)
static constexpr int TC_NSLOT = 2;       // B ring slots; a slot = up to TC_SLOT_KB K blocks of one chunk (2 x 32 KiB) + its augmented-K boxes
static constexpr int TC_SLOT_KB = 2;     //   (one barrier round trip per 8-9 MMAs instead of per 4)
static constexpr int TC_A_KB = 4;        // K blocks of A storage: one tile of d <= 256, or two tiles (double buffered) of d <= 128
static constexpr int TC_KH = 64;         // fp16 values per K block (one 128-byte swizzle row)
static constexpr int TC_MAX_KB = 4;      // K blocks of 64 halves resident per A tile: d <= 256
static constexpr int TC_KBLK_BYTES = TC_M * ROW_BYTES;  // 16 KiB: 128 rows x 128 B
static constexpr int TC_AUG_BYTES = TC_N * 32;          // 4 KiB: 128 rows x 16 halves
static constexpr int TC_SKB_BYTES = TC_NH * B_STAGE_BYTES;                        // 32 KiB: one K block of a chunk (256 rows x 128 B)
static constexpr int TC_SLOT_BYTES = TC_SLOT_KB * TC_SKB_BYTES + TC_NH * TC_AUG_BYTES;   // 72 KiB
static constexpr int TC_MAX_KB_STREAM = 16;   // d <= 1024 in the streaming mode below
// Streaming mode (d > 256, e.g. GIST's 960): the query tile no longer fits next to the B ring, so BOTH operands stream, one K
// block per ring slot: [A: 128 queries x 64 halves = 16 KiB][B: 256 vectors x 64 halves = 32 KiB][augmented-K boxes 8 KiB];
// the A tile is re-read (from L2) for every chunk. The slots overlay the A area + the B ring of the resident mode.
static constexpr int TC_STREAM_NSLOT = 3;
static constexpr int TC_STREAM_SLOT_BYTES = TC_KBLK_BYTES + TC_SKB_BYTES + TC_NH * TC_AUG_BYTES;   // 56 KiB
static_assert(TC_STREAM_NSLOT * TC_STREAM_SLOT_BYTES <= TC_A_KB * TC_KBLK_BYTES + TC_NSLOT * TC_SLOT_BYTES, "streaming ring must fit the operand area");
static constexpr int TC_PARTS = 4;       // filter pass: column parts per accumulator = epilogue warps per TMEM lane quadrant
static constexpr int TC_EPI_WARPS = 4 * TC_PARTS;   // filter pass: each epilogue warp takes 128 / TC_PARTS accumulator columns
static constexpr int TC_SEED_EPI_WARPS = 4;         // seed pass: one warp per quadrant takes all 128 columns
// warp roles: epilogue warps first (a warp may only touch the TMEM lanes 32 (w % 4) .. +31), then the TMA producer,
// the MMA issuer and the TMEM allocator
__host__ __device__ constexpr int tc_w_prod(bool seed) { return seed ? 8 : TC_EPI_WARPS; }
__host__ __device__ constexpr int tc_threads(bool seed) { return (tc_w_prod(seed) + 3) * 32; }

static constexpr int TC_CAPK = 32;       // k <= 16: candidate slots per (query, list, column part) region; a full region is
                                         //          compacted in place to its k best by the whole warp (one key per lane)
static constexpr int TC_CAPP = 64;       // k > 16: slots per region, no compaction (overflow -> the query is redone exactly)

static constexpr int TC_NQ_ = 4;         // work-item queue depth (scheduler -> MMA / epilogue)
static constexpr size_t TC_SMEM_BYTES = (size_t)TC_A_KB * TC_KBLK_BYTES                  // A tile(s)
                                        + (size_t)TC_NSLOT * TC_SLOT_BYTES              // B ring
                                        + (size_t)TC_AUG_BYTES                          // the constant augmented-K block of A
                                        + 512                                           // barriers, item queue, tmem slot
                                        + (size_t)TC_NQ_ * TC_M * 12;                   // per queued item and row: query id, |q|^2, bound

static constexpr int TC_TRACE_ROLES = 9, TC_TRACE_CHUNKS = 512, TC_TRACE_CTAS = 256, TC_TRACE_ITEMS = 24;   // + per CTA: {start ns, end ns, chunks, items}; + per CTA and item: {start ns, list, rows, queries}

// work-item queue entry (scheduler -> MMA / epilogue warps): the tile and its row range of the list
struct TcQItem {
    int list, q_begin, q_count, pad;
    long long lo, hi;
};

struct TcParams {
    const int* group_queries;        // [P] query id per slot
    const long long* list_offsets;   // [B+1]
    const ScanItem* items;           // tiles of up to 128 queries, most expensive first
    const int* n_items;
    int* work_counter;               // zeroed before launch: dynamic item scheduler
    int nk;                          // K blocks (ceil(d / 64)), <= TC_MAX_KB
    int max_rows;                    // > 0: only the first max_rows entries of each list (seed pass)
    const float* qnorm;              // [Q] |q|^2
    uint32_t* thr;                   // [Q] bound T[q] on the k-th best score, as f32_to_ordered(T): written by the
                                     //     seed pass, read AND tightened (atomicMin) by the filter pass
    unsigned long long* cand_key;    // [TC_PARTS P, cap] (score, list entry) keys: one private region per (pair, column part)
    int* cand_count;                 // [TC_PARTS P] survivors of the region's owner (may exceed cap: the query is then redone)
    int cap;
    int k;
    int is_ip;
    // approximate mode (real-valued data, fp16-rounded operands): the accumulator value differs from the exact one by at
    // most M(q) = margin_c * sqrt(qn_scale |q|^2) + margin_abs (all in the units of the scaled shadow copy); 0 = exact mode
    float margin_c, margin_abs;
    float qn_scale;                  // |q|^2 is staged as qn_scale * qnorm[q] (sigma^2 of the shadow copy; 1 in exact mode)
    int exp;                         // experiments (LIRA_TC_EXP): bit 0 = skip the survivor path (wrong results, timing only)
    long long* trace;                // debug (LIRA_TC_TRACE): [TC_TRACE_ROLES][TC_TRACE_CHUNKS] SM clock stamps of CTA 0, or null
};

// ---- tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, 32-byte swizzle: rows of 32 B (one K = 16 fp16 step), 8-row groups 256 B apart (SBO)
__device__ __forceinline__ uint64_t tc_smem_desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// asynchronous TMEM -> register load of 32 consecutive columns of this thread's lane; pair with tc_ld_wait()
__device__ __forceinline__ void tc_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// wait for the outstanding TMEM loads; r[] is threaded through the asm so that no use of the loaded
// registers can be scheduled above the wait
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// instruction descriptors: D = F32, A = B = F16 (format 0) resp. TF32 (format 2), both K-major, M = 128, N = 256 / 128
static constexpr uint32_t TC_IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_NS >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
static constexpr uint32_t TC_IDESC_F16_N128 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
static constexpr uint32_t TC_IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

static constexpr int TC_NQ = TC_NQ_;
static constexpr int TC_G = 64;   // seed pass: group minima per row
static constexpr int TC_KMAX_TIGHTEN = 16;  // in-kernel bound tightening (and the tensor-core seed) need k <= 16

// ---- small static sorting networks (registers only; every index is a compile-time constant) --------
__device__ __forceinline__ void tc_sort16(float (&v)[16]) {  // bitonic, ascending
#pragma unroll
    for (int k = 2; k <= 16; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const float a = v[i], b = v[l];
                    v[i] = up ? fminf(a, b) : fmaxf(a, b);
                    v[l] = up ? fmaxf(a, b) : fminf(a, b);
                }
            }
}
// a, b ascending -> a = the 16 smallest of the union, ascending
__device__ __forceinline__ void tc_merge_low16(float (&a)[16], const float (&b)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fminf(a[i], b[15 - i]);  // bitonic sequence of the 16 smallest
#pragma unroll
    for (int j = 8; j > 0; j >>= 1)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int l = i ^ j;
            if (l > i) {
                const float x = a[i], y = a[l];
                a[i] = fminf(x, y);
                a[l] = fmaxf(x, y);
            }
        }
}
__device__ __forceinline__ float tc_pick16(const float (&v)[16], int idx) {
    float r = v[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) r = (i == idx) ? v[i] : r;
    return r;
}

// The accumulator holds  s = 2 q.v - |v|^2  (L2; the -|v|^2 term comes from the augmented K block, the factor 2
// from the gathered query rows) or s = q.v (IP): LARGER is better, and the score is |q|^2 - s resp. -s.
// m4[i] = max of columns 4i..4i+3; returns the maximum of all 32.
__device__ __forceinline__ float tc_max32(const uint32_t (&r)[32], float (&m4)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        m4[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                      fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
    return fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
}

template <bool SEED, bool TRACE>
__global__ void __launch_bounds__(tc_threads(SEED), 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_v,
               const __grid_constant__ CUtensorMap tmap_vaug, const __grid_constant__ CUtensorMap tmap_aaug, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* sA = smem_raw;                                              // [nabuf][TC_A_KB / nabuf][128 x 128 B]
    uint8_t* sB = sA + (size_t)TC_A_KB * TC_KBLK_BYTES;     // [TC_NSLOT]{[TC_SLOT_KB][128 x 128 B], [128 x 32 B] aug}
    uint8_t* sGA = sB + (size_t)TC_NSLOT * TC_SLOT_BYTES;                // [128 x 32 B] constant augmented-K block of A
    uint64_t* bars = (uint64_t*)(sGA + TC_AUG_BYTES);
    uint64_t* a_full = bars;                        // [2]
    uint64_t* a_empty = a_full + 2;                 // [2]
    constexpr int NSLOT_MAX = TC_NSLOT > TC_STREAM_NSLOT ? TC_NSLOT : TC_STREAM_NSLOT;
    uint64_t* b_full = a_empty + 2;                 // [NSLOT_MAX]
    uint64_t* b_empty = b_full + NSLOT_MAX;         // [NSLOT_MAX]
    uint64_t* t_full = b_empty + NSLOT_MAX;         // [TC_NACC]
    uint64_t* t_empty = t_full + TC_NACC;           // [TC_NACC]
    uint64_t* i_full = t_empty + TC_NACC;           // [TC_NQ]
    uint64_t* i_empty = i_full + TC_NQ;             // [TC_NQ]
    uint64_t* ga_full = i_empty + TC_NQ;            // [1]
    TcQItem* iq = (TcQItem*)(ga_full + 1);          // [TC_NQ]
    uint32_t* tmem_slot = (uint32_t*)(iq + TC_NQ);
    // per queued item and tile row, staged by the scheduler warp so that an epilogue warp starts an item without any
    // global round trip: query id, |q|^2 and the query's bound (a snapshot: bounds only ever decrease, so a stale one is
    // merely looser)
    int* s_q = (int*)((uint8_t*)bars + 512);               // [TC_NQ][TC_M]
    float* s_qn = (float*)(s_q + TC_NQ * TC_M);            // [TC_NQ][TC_M]
    uint32_t* s_thr = (uint32_t*)(s_qn + TC_NQ * TC_M);    // [TC_NQ][TC_M]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TC_W_PROD = tc_w_prod(SEED), TC_W_MMA = TC_W_PROD + 1, TC_W_ALLOC = TC_W_PROD + 2;
    constexpr int N_EPI = SEED ? TC_SEED_EPI_WARPS : TC_EPI_WARPS;
    const bool aug = !p.is_ip;   // inner product: the accumulator is q.v itself, no augmented block
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NSLOT_MAX; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < TC_NACC; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], N_EPI); }
        for (int i = 0; i < TC_NQ; ++i) { mbar_init(&i_full[i], 1); mbar_init(&i_empty[i], 1 + N_EPI); }  // MMA + epilogue warps
        mbar_init(ga_full, 1);
        mbar_fence_init();
    }
    if (warp == TC_W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == TC_W_PROD * 32) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_vaug); tma_prefetch_desc(&tmap_aaug); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // (the prologue above touches nothing another kernel produces)
    pdl_launch();
    const int n_items = *p.n_items;
    const int nk = p.nk;
    // A tiles: with d <= 128 (nk <= 2) two query tiles fit, so the next work item's queries load while this item's MMAs run
    const int nabuf = nk <= TC_A_KB / 2 ? 2 : 1;
    const int a_kb = TC_A_KB / nabuf;
    // d > 256: streaming mode (see TC_STREAM_*): A rides in the ring slots, one K block per slot
    const bool stream_a = nk > TC_A_KB;
    const int slot_kb = stream_a ? 1 : TC_SLOT_KB;                          // K blocks per ring slot
    const int nslot = stream_a ? TC_STREAM_NSLOT : TC_NSLOT;
    const uint32_t slot_bytes = stream_a ? TC_STREAM_SLOT_BYTES : TC_SLOT_BYTES;
    uint8_t* const ring = stream_a ? sA : sB;
    const uint32_t slot_b_off = stream_a ? TC_KBLK_BYTES : 0;               // B operand inside a slot
    const uint32_t slot_aug_off = slot_b_off + slot_kb * TC_SKB_BYTES;      // augmented-K boxes inside a slot

    // debug timeline of CTA 0: one clock stamp per (role, chunk)
    auto stamp = [&](int role, uint32_t chunk) {
        if constexpr (TRACE) {
            if (p.trace && blockIdx.x == 0 && chunk < TC_TRACE_CHUNKS && lane == 0) p.trace[role * TC_TRACE_CHUNKS + chunk] = clock64();
        }
    };

    if (warp == TC_W_PROD) {
        // ===== scheduler + TMA producer: the whole warp runs the loop, one elected lane issues (see elect_one) =====
        if (aug) {   // the constant augmented-K block of A: every row is (-2048, -1, 0, ..., 0)
            if (elect_one()) {
                mbar_arrive_expect_tx(ga_full, TC_AUG_BYTES);
                tma_load_2d(sGA, &tmap_aaug, 0, 0, ga_full);
            }
            __syncwarp();
        }
        PipeState bs{0, 0};
        uint32_t mp = 0;
        // Dynamic scheduler, ONE work item ahead (more would hurt the balance at the tail: an item is ~10 % of a
        // CTA's share). The dependent global round trips that describe the next item -- ticket (atomicAdd),
        // descriptor, list range + row ids, per-row |q|^2 and bound -- are issued one per chunk of the current item,
        // each consuming the result of the previous one, so none of them is waited for between two items.
        auto ticket = [&]() { return lane == 0 ? atomicAdd(p.work_counter, 1) : 0; };
        auto descriptor = [&](int t) {
            const int idx = __shfl_sync(0xffffffffu, t, 0);
            ScanItem d;
            if (idx < n_items) d = p.items[idx];
            else { d.list = -1; d.q_begin = 0; d.q_count = 0; d.tm = 0; }
            return d;
        };
        auto with_rows = [&](const ScanItem& d) {
            TcQItem o;
            o.list = d.list; o.q_begin = d.q_begin; o.q_count = d.q_count; o.pad = 0;
            o.lo = 0; o.hi = 0;
            if (d.list >= 0) {
                o.lo = p.list_offsets[d.list];
                o.hi = p.list_offsets[d.list + 1];
                if (p.max_rows > 0 && o.hi - o.lo > p.max_rows) o.hi = o.lo + p.max_rows;   // seed pass
            }
            return o;
        };
        // per-row data of an item, 4 rows per lane (row = lane + 32 j)
        int rq[4];
        float rqn[4];
        uint32_t rthr[4];
        auto row_ids = [&](const ScanItem& d) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rq[j] = (d.list >= 0 && lane + 32 * j < d.q_count) ? __ldg(p.group_queries + d.q_begin + lane + 32 * j) : -1;
        };
        auto row_data = [&]() {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rqn[j] = rq[j] >= 0 ? __ldg(p.qnorm + rq[j]) * p.qn_scale : 0.f;
                rthr[j] = (rq[j] >= 0 && !SEED) ? *reinterpret_cast<const volatile uint32_t*>(p.thr + rq[j]) : 0xFF800000u;
            }
        };
        TcQItem cur;
        {
            const ScanItem d0 = descriptor(ticket());
            cur = with_rows(d0);
            row_ids(d0);
            row_data();
        }
        constexpr int N_STAGES = 5;
        for (int n = 0;; ++n) {
            const TcQItem it = cur;
            const int qs = n % TC_NQ;
            mbar_wait(&i_empty[qs], ((n / TC_NQ) & 1) ^ 1u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_q[qs * TC_M + lane + 32 * j] = rq[j];
                s_qn[qs * TC_M + lane + 32 * j] = rqn[j];
                s_thr[qs * TC_M + lane + 32 * j] = rthr[j];
            }
            __syncwarp();   // every lane's rows are written before lane 0 publishes the item (arrive = release)
            if (lane == 0) { iq[qs] = it; mbar_arrive(&i_full[qs]); }
            __syncwarp();
            if (it.list < 0) break;
            // next item: ticket -> descriptor -> list range and row ids -> per-row data, one stage per chunk
            int t_next = 0, stage = 0;
            ScanItem d_next;
            d_next.list = -1; d_next.q_begin = 0; d_next.q_count = 0; d_next.tm = 0;
            auto advance_prefetch = [&]() {
                if (stage == 0) t_next = ticket();
                else if (stage == 1) d_next = descriptor(t_next);
                else if (stage == 2) { cur = with_rows(d_next); row_ids(d_next); }
                else if (stage == 3) row_data();
                ++stage;
            };
            const int ab = n % nabuf;
            if (!stream_a) {
                mbar_wait(&a_empty[ab], ((n / nabuf) & 1) ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[ab], (uint32_t)nk * TC_KBLK_BYTES);
                    for (int kb = 0; kb < nk; ++kb)
                        tma_load_2d(sA + (size_t)(ab * a_kb + kb) * TC_KBLK_BYTES, &tmap_q, kb * TC_KH, it.q_begin, &a_full[ab]);
                }
                __syncwarp();
            }
            const long long lo = it.lo, hi = it.hi;
            for (long long row0 = lo; row0 < hi; row0 += TC_NS, ++mp) {
                stamp(0, mp);
                if (stage < N_STAGES) advance_prefetch();
                const int nh = hi - row0 > TC_N ? TC_NH : 1;   // boxes of 128 rows in this chunk (the tail of a list may need one only)
                for (int kb0 = 0; kb0 < nk; kb0 += slot_kb) {   // one slot = up to slot_kb K blocks (+ the aug boxes with the first)
                    const int nkb = min(slot_kb, nk - kb0);
                    const bool with_aug = aug && kb0 == 0;
                    mbar_wait(&b_empty[bs.stage], bs.phase ^ 1u);
                    if (elect_one()) {
                        uint8_t* slot = ring + (size_t)bs.stage * slot_bytes;
                        mbar_arrive_expect_tx(&b_full[bs.stage], (uint32_t)(nkb * nh) * B_STAGE_BYTES + (with_aug ? nh * TC_AUG_BYTES : 0) +
                                                                     (stream_a ? (uint32_t)TC_KBLK_BYTES : 0u));
                        if (stream_a) tma_load_2d(slot, &tmap_q, kb0 * TC_KH, it.q_begin, &b_full[bs.stage]);   // this K block of the queries
                        for (int j = 0; j < nkb; ++j)
                            for (int x = 0; x < nh; ++x)
                                tma_load_2d(slot + slot_b_off + (size_t)j * TC_SKB_BYTES + (size_t)x * B_STAGE_BYTES, &tmap_v, (kb0 + j) * TC_KH,
                                            (int)row0 + x * TC_N, &b_full[bs.stage]);
                        if (with_aug)
                            for (int x = 0; x < nh; ++x)
                                tma_load_2d(slot + slot_aug_off + (size_t)x * TC_AUG_BYTES, &tmap_vaug, 0, (int)row0 + x * TC_N,
                                            &b_full[bs.stage]);
                    }
                    __syncwarp();
                    bs.advance(nslot);
                }
            }
            while (stage < N_STAGES) advance_prefetch();   // short list: the rest of the look-ahead (waited for)
        }
    } else if (warp == TC_W_MMA) {
        // ===== MMA issuer: the whole warp runs the loop (uniform operands), one elected lane issues =====
        PipeState bs{0, 0};
        uint32_t m = 0;  // running chunk counter -> accumulator slot and phase
        long long cta_t0 = 0;
        int cta_items = 0;
        if constexpr (TRACE) cta_t0 = (long long)global_timer_ns();
        if (aug) { mbar_wait(ga_full, 0); tc_fence_after(); }
        const uint32_t sA_u32 = smem_u32(sA), ring_u32 = smem_u32(ring);
        const uint64_t ga_desc = tc_smem_desc_sw32(smem_u32(sGA));
        for (int n = 0;; ++n) {
            const int qs = n % TC_NQ;
            mbar_wait(&i_full[qs], (n / TC_NQ) & 1);
            const TcQItem it = iq[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&i_empty[qs]);
            if (it.list < 0) break;
            const int ab = n % nabuf;
            if (!stream_a) {
                mbar_wait(&a_full[ab], (n / nabuf) & 1);
                tc_fence_after();
            }
            const long long lo = it.lo, hi = it.hi;
            if constexpr (TRACE) {
                if (p.trace && lane == 0 && blockIdx.x < TC_TRACE_CTAS && n < TC_TRACE_ITEMS) {
                    long long* c = p.trace + (size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS + TC_TRACE_CTAS * 4 + ((size_t)blockIdx.x * TC_TRACE_ITEMS + n) * 4;
                    c[0] = (long long)global_timer_ns(); c[1] = it.list; c[2] = hi - lo; c[3] = it.q_count;
                }
            }
            for (long long row0 = lo; row0 < hi; row0 += TC_NS, ++m) {
                const uint32_t acc = m & (TC_NACC - 1);
                mbar_wait(&t_empty[acc], ((m / TC_NACC) & 1) ^ 1u);
                tc_fence_after();
                stamp(1, m);
                const uint32_t d_tmem = tmem_base + acc * TC_NS;
                const uint32_t idesc = hi - row0 > TC_N ? TC_IDESC_F16 : TC_IDESC_F16_N128;   // (the tail of a list: N = 128)
                for (int kb0 = 0; kb0 < nk; kb0 += slot_kb) {
                    const int nkb = min(slot_kb, nk - kb0);
                    mbar_wait(&b_full[bs.stage], bs.phase);
                    tc_fence_after();
                    if (kb0 == 0) stamp(2, m);
                    if (kb0 + slot_kb >= nk) stamp(3, m);
                    const uint32_t slot = ring_u32 + (uint32_t)bs.stage * slot_bytes;
                    if (elect_one()) {
                        if (aug && kb0 == 0)   // s = -|v|^2 ...
                            tc_mma_f16(d_tmem, ga_desc, tc_smem_desc_sw32(slot + slot_aug_off), idesc, 0u);
                        for (int jb = 0; jb < nkb; ++jb) {   // ... + (2 q) . v
                            const uint32_t a_addr = stream_a ? slot : sA_u32 + (uint32_t)(ab * a_kb + kb0 + jb) * TC_KBLK_BYTES;
                            const uint32_t b_addr = slot + slot_b_off + (uint32_t)jb * TC_SKB_BYTES;
#pragma unroll
                            for (int j = 0; j < 4; ++j)  // 4 x K = 16 fp16 (32 bytes) inside the 128-byte swizzle row
                                tc_mma_f16(d_tmem, tc_smem_desc(a_addr + j * 32), tc_smem_desc(b_addr + j * 32), idesc,
                                            (aug || kb0 || jb || j) ? 1u : 0u);
                        }
                        tc_commit(&b_empty[bs.stage]);  // frees the slot when these MMAs have read it
                    }
                    __syncwarp();
                    bs.advance(nslot);
                }
                if (elect_one()) tc_commit(&t_full[acc]);            // accumulator complete -> epilogue
                __syncwarp();
                stamp(4, m);
            }
            if (!stream_a) {
                if (elect_one()) tc_commit(&a_empty[ab]);            // all MMAs reading this A tile are done
                __syncwarp();
            }
            ++cta_items;
        }
        if constexpr (TRACE) {
            if (p.trace && lane == 0 && blockIdx.x < TC_TRACE_CTAS) {
                long long* c = p.trace + (size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS + blockIdx.x * 4;
                c[0] = cta_t0; c[1] = (long long)global_timer_ns(); c[2] = m; c[3] = cta_items;
            }
        }
    } else if (warp < N_EPI) {
        // ===== epilogue: TMEM -> registers -> (seed: group minima | filter: survivors) =====
        // TC_PARTS warps per TMEM lane quadrant: a thread owns one query row and one part (128 / TC_PARTS columns) of
        // every accumulator of the work item (many warps with little work each: the survivor path is a chain of
        // dependent selects, votes and stores, and only other warps can hide its latency). With t = -s (t = score - |q|^2 for L2, t = score for IP; smaller is
        // better) the filter keeps t <= tq. The hot loop is a 3-input-max tree over the raw accumulator values and
        // ONE compare per 32 columns (FMNMX3 only: the norms are already inside the accumulator). About one pair in
        // a thousand survives, i.e. a good part of a warp's 32-column groups hold one, but a thread sees only one
        // or two survivors per list. So the survivor path is kept short and warp-uniform: a lane picks its next
        // passing 4-column block with a select tree (no dynamic register indexing) and APPENDS the passing entries
        // to its PRIVATE candidate region in global memory (plain stores, no atomics, no shared staging, nothing
        // kept in registers). Only a region that fills up (k <= 16: 32 slots) is compacted, by the whole warp, to
        // its k best entries in sorted order; its k-th score is then an exact bound for the row and is published
        // (atomicMin) for the query's rows in other lists and CTAs, which re-read it every chunk.
        // (the seed pass has no survivor path and one bound per ROW to produce: it runs on 4 warps that take all
        //  128 columns, so the 64 group minima of a row cover all its 1024 seed entries)
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int part = SEED ? 0 : warp >> 2;     // which part of the 128 accumulator columns
        const int row = quad * 32 + lane;          // query row of this thread inside the tile
        constexpr int NCOL = SEED ? TC_N : TC_N / TC_PARTS;   // columns per thread and chunk
        constexpr int NG = NCOL / 32;                         // 32-column groups per thread and chunk
        const bool keep_mode = p.k <= TC_KMAX_TIGHTEN;
        const int cap = p.cap;
        const uint32_t t_full_u32 = smem_u32(t_full), t_empty_u32 = smem_u32(t_empty);
        uint32_t m = 0;
        for (int n = 0;; ++n) {
            const int qs = n % TC_NQ;
            mbar_wait(&i_full[qs], (n / TC_NQ) & 1);
            const TcQItem it = iq[qs];
            const int q_s = s_q[qs * TC_M + row];
            const float qn_s = s_qn[qs * TC_M + row];
            const uint32_t thr_s = s_thr[qs * TC_M + row];
            __syncwarp();
            if (lane == 0) mbar_arrive(&i_empty[qs]);
            if (it.list < 0) break;
            const bool row_ok = row < it.q_count;
            int q = 0, cnt = 0;
            float qn = 0.f, tq = -INFINITY, M = 0.f;
            float gmin[SEED ? TC_G : 1];
#pragma unroll
            for (int g = 0; g < (SEED ? TC_G : 1); ++g) gmin[g] = INFINITY;
            bool lost = false;     // a tie at the bound fell out of a compacted region: redo the query exactly
            unsigned long long* cand = nullptr;
            const volatile uint32_t* thr_q = p.thr;   // (rows past the end of the tile read thr[0]; their tq stays -inf)
            uint32_t thr_pref = 0xFF800000u;
            if (row_ok) {
                q = q_s;
                qn = p.is_ip ? 0.f : qn_s;
                M = p.margin_c > 0.f ? fmaf(p.margin_c, sqrtf(qn_s), p.margin_abs) : 0.f;
                if (!SEED) {
                    thr_q = p.thr + q;
                    thr_pref = thr_s;
                    // exact score <= T  <=  accumulator-side t <= T - qn + M  (qn = 0 for IP, M = 0 in exact mode): tq is
                    // the bound the (possibly rounded) accumulator values are compared with
                    tq = ordered_to_f32(thr_pref) - qn + M;
                    cand = p.cand_key + ((size_t)(it.q_begin + row) * TC_PARTS + part) * cap;
                }
            }
            // 32-bit chunk bookkeeping (an index holds fewer than 2^31 entries): first entry of this thread's columns in
            // the current chunk, and how many entries of the list are left from there
            uint32_t ebase = (uint32_t)it.lo + part * NCOL;
            int left = (int)(it.hi - it.lo) - part * NCOL;
            // a warp whose 32 tile rows are all padding (tile of fewer queries) only keeps the accumulator handshake going
            const bool warp_idle = !SEED && quad * 32 >= it.q_count;
            for (int rows_left = (int)(it.hi - it.lo); rows_left > 0; rows_left -= TC_NS, ebase += TC_NS, left -= TC_NS, ++m) {
                const uint32_t acc = m & (TC_NACC - 1);
                const int nh = rows_left > TC_N ? TC_NH : 1;   // halves of 128 columns the MMA warp filled (tail of a list: one)
                if (warp_idle) {
                    mbar_wait_addr(t_full_u32 + acc * 8, (m / TC_NACC) & 1);
                    if (lane == 0) mbar_arrive_addr(t_empty_u32 + acc * 8);
                    __syncwarp();
                    continue;
                }
                // the bound the query's rows in other lists / CTAs have published meanwhile: the value loaded one
                // chunk ago is used now and the next one is requested, so the load latency is never waited for
                const uint32_t thr_now = thr_pref;
                if (!SEED) thr_pref = *thr_q;
                mbar_wait_addr(t_full_u32 + acc * 8, (m / TC_NACC) & 1);
                tc_fence_after();
                if (warp == 0) stamp(5, m);
                if (warp == 5) stamp(7, m);
                if (!SEED && row_ok) tq = fminf(tq, ordered_to_f32(thr_now) - qn + M);
                for (int h = 0; h < nh; ++h) {   // the two 128-column halves of the accumulator, one after the other
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TC_NS + h * TC_N + part * NCOL;
                uint32_t ra[32], rb[NG > 1 ? 32 : 1];
                tc_ld32_async(taddr, ra);
                if constexpr (NG > 1) tc_ld32_async(taddr + 32, rb);
                tc_ld_wait(ra);   // (waits for both loads)
                if constexpr (NG > 1) tc_ld_wait(rb);
                if (!SEED && h == nh - 1) {
                    // the thread's last columns are in registers: hand the accumulator back to the MMA warp at once
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_addr(t_empty_u32 + acc * 8);
                }
                // columns past the end of the list hold other lists' vectors: valid columns of this thread's range
                const int n_valid = min(NCOL, left - h * TC_N);
                const uint32_t ebase_h = ebase + h * TC_N;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if constexpr (SEED) {
                        if (g == 2) {   // second half of the columns (no call is made in the seed pass)
                            tc_ld32_async(taddr + 64, ra);
                            tc_ld32_async(taddr + 96, rb);
                            tc_ld_wait(ra);
                            tc_ld_wait(rb);
                        }
                    }
                    uint32_t (&r)[32] = (NG > 1 && (g & 1)) ? reinterpret_cast<uint32_t (&)[32]>(rb) : ra;
                    if (n_valid < (g + 1) * 32) {   // partial group (only the last chunk of a list): mask the tail
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (g * 32 + c >= n_valid) r[c] = 0xFF800000u;   // -inf: never the maximum, never passes
                    }
                    float m4[8];
                    const float mx = tc_max32(r, m4);
                    if (SEED) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            gmin[(g * 32 + c) & (TC_G - 1)] = fminf(gmin[(g * 32 + c) & (TC_G - 1)], -__uint_as_float(r[c]));
                    } else if (__any_sync(0xffffffffu, -mx <= tq) && !(p.exp & 1)) {
                        // (tq = -inf for rows past the end of the tile, so they never pass)
                        uint32_t qm = 0;   // passing 4-column blocks of this lane still to look at
#pragma unroll
                        for (int i = 0; i < 8; ++i) qm |= (-m4[i] <= tq) ? (1u << i) : 0u;
                        uint32_t pm = 0;   // passing entries of the current block still to append
                        uint32_t e0 = 0;
                        float s4[4] = {0.f, 0.f, 0.f, 0.f};
                        do {   // one round = at most one new block and one appended entry per lane; the warp stays converged
                            const int j = (pm == 0) ? __ffs(qm) - 1 : -1;   // -1: no new block for this lane in this round
                            if (pm == 0) qm &= qm - 1;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {   // select tree: no dynamic register indexing
                                const uint32_t x0 = (j & 1) ? r[4 + u] : r[u], x1 = (j & 1) ? r[12 + u] : r[8 + u];
                                const uint32_t x2 = (j & 1) ? r[20 + u] : r[16 + u], x3 = (j & 1) ? r[28 + u] : r[24 + u];
                                const uint32_t y0 = (j & 2) ? x1 : x0, y1 = (j & 2) ? x3 : x2;
                                const float v = -__uint_as_float((j & 4) ? y1 : y0);
                                s4[u] = (j >= 0) ? v : s4[u];
                                pm |= (j >= 0 && v <= tq) ? (1u << u) : 0u;
                            }
                            e0 = (j >= 0) ? ebase_h + (uint32_t)(g * 32 + j * 4) : e0;
                            if (pm) {
                                const int u = __ffs(pm) - 1;
                                pm &= pm - 1;
                                const float x = (u & 2) ? ((u & 1) ? s4[3] : s4[2]) : ((u & 1) ? s4[1] : s4[0]);
                                // (the bound may have tightened since the block was looked at; masked columns are +inf)
                                if (x <= tq && x < INFINITY) {
                                    if (cnt < cap) cand[cnt] = make_key(x + qn, e0 + u);
                                    ++cnt;   // k > 16: may run past cap -> the query is redone exactly
                                }
                            }
                            if (keep_mode) {
                                // a full region (32 keys, one per lane) -> its k best, in sorted order, by the whole warp
                                uint32_t fm = __ballot_sync(0xffffffffu, cnt >= TC_CAPK);
                                while (fm) {
                                    const int L = __ffs(fm) - 1;
                                    fm &= fm - 1;
                                    __syncwarp();   // lane L's stores are visible to the warp
                                    unsigned long long* reg = (unsigned long long*)shfl_u64((unsigned long long)cand, L);
                                    const unsigned long long key = __ldcg(reg + lane);
                                    int rank = 0;   // keys of one region are distinct (distinct list entries)
#pragma unroll 8
                                    for (int i = 0; i < 32; ++i) rank += (shfl_u64(key, i) < key) ? 1 : 0;
                                    __syncwarp();   // every lane holds its key before the region is rewritten
                                    const uint32_t sc = (uint32_t)(key >> 32);   // ordered score bits
                                    const uint32_t mk = __ballot_sync(0xffffffffu, rank == p.k - 1);
                                    const uint32_t sk = __shfl_sync(0xffffffffu, sc, __ffs(mk) - 1);     // k-th best score
                                    // kept: everything at or below the k-th score (ties stay, their id order is settled by
                                    // the refine pass); ranks are the sorted order, so the kept keys are ranks 0 .. n_keep-1
                                    // (approximate mode: the k kept candidates' exact scores are at most sk + M, and an entry
                                    //  can only be dropped when its exact score is certainly above that: approx > sk + 2 M)
                                    const float sk_f = ordered_to_f32(sk);
                                    const float M_L = __shfl_sync(0xffffffffu, M, L);
                                    const bool keep = sc <= (M_L > 0.f ? f32_to_ordered(sk_f + 2.f * M_L) : sk);
                                    if (keep) reg[rank] = key;
                                    const int n_keep = __popc(__ballot_sync(0xffffffffu, keep));
                                    if (lane == L) {
                                        cnt = n_keep;
                                        if (n_keep >= TC_CAPK - 4) { lost = true; cnt = p.k; }   // (nearly) all ties: give the query up
                                        tq = fminf(tq, sk_f + M - qn + M);
                                        atomicMin(p.thr + q, M > 0.f ? f32_to_ordered(sk_f + M) : sk);
                                    }
                                    __syncwarp();
                                }
                            }
                        } while (__any_sync(0xffffffffu, (qm | pm) != 0));
                    }
                }
                }   // h
                if (SEED) {   // this warp is done with the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[acc]);
                }
                if (warp == 0) stamp(6, m);
                if (warp == 5) stamp(8, m);
            }
            if (SEED) {
                // T = k-th smallest of the row's 64 group minima of t: distinct entries of ONE list, so k real
                // candidates are at or below it (score = t + |q|^2 for L2, t for IP). Needs k <= 16.
                if (row_ok) {
                    float a[16], b[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { a[j] = gmin[j]; b[j] = gmin[16 + j]; }
                    tc_sort16(a); tc_sort16(b); tc_merge_low16(a, b);
                    float c[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { b[j] = gmin[32 + j]; c[j] = gmin[48 + j]; }
                    tc_sort16(b); tc_sort16(c); tc_merge_low16(b, c);
                    tc_merge_low16(a, b);
                    const float tk = tc_pick16(a, p.k - 1);
                    atomicMin(p.thr + q, f32_to_ordered(tk + qn + M));
                }
            } else if (row_ok) {
                p.cand_count[(size_t)(it.q_begin + row) * TC_PARTS + part] = lost ? cap + 1 : cnt;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- helpers around the filter --------------------------------------------------------------------
// |x|^2 per row in fp32 (sequential order); *exact_flag is cleared unless every value is an integer of at
// most 11 bits (exact in fp16) and every |x|^2 < 2^22: then all products / partial sums of the tensor-core
// path are integers below 2^24 and its result equals the fp32 direct-difference result bit for bit.
// max_norm_bits (may be null): atomicMax of the float bits of |x|^2 (a NaN / inf row makes it NaN / inf).
__global__ void row_norms_kernel(const float* __restrict__ x, long ld, int d, long long n, float* __restrict__ out,
                                 int* __restrict__ exact_flag, uint32_t* __restrict__ max_norm_bits) {
    bool bad = false;
    float mx = 0.f;
    bool nan = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* r = x + i * ld;
        float s = 0.f;
        for (int j = 0; j < d; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(r + j);
            s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
            bad |= (v.x != rintf(v.x)) | (v.y != rintf(v.y)) | (v.z != rintf(v.z)) | (v.w != rintf(v.w));
            bad |= !(fabsf(v.x) <= 2047.f && fabsf(v.y) <= 2047.f && fabsf(v.z) <= 2047.f && fabsf(v.w) <= 2047.f);
        }
        out[i] = s;
        bad |= !(s < 4194304.0f);  // |x|^2 < 2^22  =>  |q|^2 + |v|^2 + 2|q.v| < 2^24
        nan |= !(s == s);
        mx = fmaxf(mx, s);
    }
    if (bad && exact_flag) *exact_flag = 0;
    if (max_norm_bits) atomicMax(max_norm_bits, nan ? 0x7FC00000u : __float_as_uint(mx));
}

// The fp16 shadow copy of the rows for the tensor-core scan: x16[n, d16] = fp16(sigma x) (zero padded, d16 % 8 == 0) and
// the row's augmented-K block vaug[n, 16] that puts -sigma^2 |x|^2 into the accumulator:
//   exact mode (integers of <= 11 bits, sigma = 1):  (hi, lo, 0 ..) with |x|^2 = 2048 hi + lo, A side (-2048, -1, 0 ..)
//   approximate mode (real-valued data):             (hi, lo, 0 ..) with sigma^2 |x|^2 = hi + lo (+- 2^-21 relative), A side (-1, -1, 0 ..)
__global__ void shadow_rows_kernel(const float* __restrict__ x, long ld, int d, long long n, const float* __restrict__ norm,
                                   float sigma, int exact_mode, __half* __restrict__ x16, int d16, __half* __restrict__ vaug) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* r = x + i * ld;
        for (int j = 0; j < d; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(r + j);
            __half2 h0 = __floats2half2_rn(v.x * sigma, v.y * sigma), h1 = __floats2half2_rn(v.z * sigma, v.w * sigma);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(x16 + i * d16 + j) = pk;
        }
        for (int j = d; j < d16; j += 4) *reinterpret_cast<uint2*>(x16 + i * d16 + j) = make_uint2(0u, 0u);
        const float s = norm[i] * sigma * sigma;
        float hi, lo;
        if (exact_mode) {
            hi = floorf(s * (1.0f / 2048.0f));
            lo = s - hi * 2048.0f;
        } else {
            hi = __half2float(__float2half_rn(s));
            lo = s - hi;
        }
        __half2 h = __floats2half2_rn(hi, lo);
        *reinterpret_cast<uint4*>(vaug + i * 16) = make_uint4(*reinterpret_cast<uint32_t*>(&h), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(vaug + i * 16 + 8) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// gq[slot, :] = fp16(scale * q[group_queries[slot], :])  (queries in group order, so a tile of a group is one TMA box;
// zero padded to d16 columns). n_slots (device scalar = group_offsets[B]) is the number of valid slots: explicit probe
// sets may hold invalid (-1) entries, so it can be smaller than the host-side bound P the buffers were sized with.
// ok_flag (may be null) is cleared when a scaled value does not fit fp16 (approximate mode: the batch then takes the CUDA cores).
__global__ void gather_group_queries_kernel(const float* __restrict__ q, long ldq, int ds, const int* __restrict__ group_queries,
                                            long long P, const long long* __restrict__ n_slots, __half* __restrict__ gq, int d16,
                                            float scale, int* __restrict__ ok_flag) {
    const int per_row = d16 / 4;
    const long long total = (*n_slots < P ? *n_slots : P) * per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / per_row;
        const int c = (int)(i % per_row) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < ds) v = *reinterpret_cast<const float4*>(q + (long long)group_queries[s] * ldq + c);
        if (ok_flag && !(fabsf(v.x * scale) <= 60000.f && fabsf(v.y * scale) <= 60000.f && fabsf(v.z * scale) <= 60000.f &&
                         fabsf(v.w * scale) <= 60000.f))
            *ok_flag = 0;
        __half2 h0 = __floats2half2_rn(v.x * scale, v.y * scale), h1 = __floats2half2_rn(v.z * scale, v.w * scale);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(gq + s * d16 + c) = pk;
    }
}

// T[q] = k-th best exact score over DISTINCT ids of the seed scan's (up to) two partial lists of the query,
// +inf when fewer than k distinct candidates were seen. Any k real candidates bound the final k-th score.
__global__ void seed_threshold_kernel(const unsigned long long* part_key, const int* probe_slot, const int* seed_ids, int Q,
                                      int k, uint32_t* thr) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += gridDim.x * blockDim.x) {
        const unsigned long long* a = seed_ids[2 * q] >= 0 ? part_key + (size_t)probe_slot[2 * q] * k : nullptr;
        const unsigned long long* b = seed_ids[2 * q + 1] >= 0 ? part_key + (size_t)probe_slot[2 * q + 1] * k : nullptr;
        int ia = 0, ib = 0, n = 0;
        unsigned long long last = KEY_INF, kth = KEY_INF;
        while (n < k) {
            const unsigned long long xa = (a && ia < k) ? a[ia] : KEY_INF;
            const unsigned long long xb = (b && ib < k) ? b[ib] : KEY_INF;
            const unsigned long long x = xa < xb ? xa : xb;
            if (x == KEY_INF) break;
            if (xa < xb) ++ia; else ++ib;
            if (x == last) continue;  // the same id in both lists (learned redundancy): identical key
            last = x;
            kth = x;
            ++n;
        }
        thr[q] = f32_to_ordered((n == k) ? key_score(kth) : INFINITY);
    }
}

__global__ void fill_u32_kernel(uint32_t* x, long long n, uint32_t v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
// seed1[q] = seed_ids[2 q]: the single best list per query for the tensor-core seed pass
__global__ void first_of_pairs_kernel(const int* seed_ids, int Q, int* seed1) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += gridDim.x * blockDim.x) seed1[q] = seed_ids[2 * q];
}

// explicit probe sets: seed with the first two probed lists of each query
__global__ void first_probes_kernel(const long long* probe_offsets, const int* probe_ids, int Q, int* seed_ids) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += gridDim.x * blockDim.x) {
        const long long lo = probe_offsets[q], hi = probe_offsets[q + 1];
        seed_ids[2 * q + 0] = hi > lo ? probe_ids[lo] : -1;
        seed_ids[2 * q + 1] = hi > lo + 1 ? probe_ids[lo + 1] : -1;
    }
}

// exhaustive probe sets: seed every query with lists 0 and 1
__global__ void seed_first_lists_kernel(int Q, int B, int* seed_ids) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += gridDim.x * blockDim.x) {
        seed_ids[2 * q + 0] = 0;
        seed_ids[2 * q + 1] = B > 1 ? 1 : -1;
    }
}

// refine: the candidate regions (score, list entry) of a query's probed (list, column part) pairs -> exact top-k over
// distinct ids. One warp per query.
struct RefineParams {
    const unsigned long long* cand_key;   // [PARTS P, cap]  (PARTS = TC_PARTS regions per pair, or 1 for the byte scan)
    const int* cand_count;                // [PARTS P]
    int cap;
    const long long* probe_offsets;       // [Q+1]
    const int* probe_slot;                // [P] slot of the j-th probe of a query (-1: invalid probe)
    const int* list_ids;
    int k, Q, dedup, is_ip;
    float* out_dist;
    long long* out_ids;
    int* redo;      // [Q] 1 when a region of the query overflowed `cap` (its row is left for the exact path)
    int* n_redo;    // number of such queries
    // approximate mode (EXACT = true): the regions hold approximate scores; every candidate that can still make the top k
    // is scored again exactly from the fp32 rows (direct difference / dot product in fp32, as the CUDA-core scan does)
    const float* vecs;      // [E, ldv] fp32 list rows
    long long ldv;
    const float* q;         // [Q, ldq] queries
    long long ldq;
    int d;
    const float* qnorm;     // [Q] |q|^2
    float qn_scale, margin_c, margin_abs, inv_scale;   // scaled units of the shadow copy; inv_scale = 1 / sigma^2
};

template <int S, bool EXACT, int PARTS = TC_PARTS>
__global__ void __launch_bounds__(256) refine_topk_kernel(const RefineParams p) {
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int k = p.k;
    // EXACT: this lane's dimensions (lane, lane + 32, ...) of the query, d <= 256; the error bound M of the approximate scores
    float qreg[EXACT ? 8 : 1];
    float M = 0.f;
    if (EXACT) {
#pragma unroll
        for (int j = 0; j < 8; ++j) qreg[j] = (lane + 32 * j < p.d) ? p.q[(size_t)q * p.ldq + lane + 32 * j] : 0.f;
        M = fmaf(p.margin_c, sqrtf(p.qnorm[q] * p.qn_scale), p.margin_abs);
    }
    unsigned long long key[S];
#pragma unroll
    for (int s = 0; s < S; ++s) key[s] = KEY_INF;
    unsigned long long kth = KEY_INF;
    bool overflow = false;
    const long long lo = p.probe_offsets[q], hi = p.probe_offsets[q + 1];
    constexpr int PPP = 32 / PARTS;   // probes per 32 regions
    for (long long j0 = lo; j0 < hi; j0 += 2 * PPP) {
        // 2 PPP probes = 64 regions per pass, TWO per lane: lane l looks at regions (probe j0 + l / TC_PARTS, part l % TC_PARTS) and
        // (probe j0 + PPP + l / TC_PARTS, same part); the loads of both are in flight together, which halves the chain of dependent
        // global round trips (slot -> count -> keys -> ids) a query with more than PPP probes has to wait for
        int cnt2[2];
        const unsigned long long* src2[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const long long j = j0 + w * PPP + (lane / PARTS);
            int region = -1, cnt = 0;
            if (j < hi) {
                const int slot = p.probe_slot[j];
                if (slot >= 0) {
                    region = slot * PARTS + (lane % PARTS);
                    cnt = p.cand_count[region];
                }
            }
            overflow |= cnt > p.cap;
            cnt2[w] = cnt < p.cap ? cnt : p.cap;
            src2[w] = p.cand_key + (size_t)(region < 0 ? 0 : region) * p.cap;
        }
        const int max_cnt = __reduce_max_sync(0xffffffffu, max(cnt2[0], cnt2[1]));
        for (int r0 = 0; r0 < max_cnt; r0 += 8) {
            // every lane pulls up to 8 entries of EACH of its two regions: all key loads, then all id loads, are in flight together
            unsigned long long x[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = (r0 + (r & 7) < cnt2[r >> 3]) ? src2[r >> 3][r0 + (r & 7)] : KEY_INF;
            if (!EXACT) {
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (x[r] != KEY_INF) x[r] = (x[r] & 0xFFFFFFFF00000000ull) | (uint32_t)__ldg(p.list_ids + key_pos(x[r]));  // entry -> global id
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (r0 + (r & 7) >= max_cnt) continue;
                uint32_t mm = EXACT ? __ballot_sync(0xffffffffu, x[r] != KEY_INF) : __ballot_sync(0xffffffffu, x[r] < kth);
                while (mm) {
                    const int sl = __ffs(mm) - 1;
                    mm &= mm - 1;
                    unsigned long long y = shfl_u64(x[r], sl);
                    if (EXACT) {
                        // the approximate score minus its error bound is a lower bound of the exact score
                        const float lower = (key_score(y) - M) * p.inv_scale;
                        if (kth != KEY_INF && lower > key_score(kth)) continue;
                        const uint32_t entry = key_pos(y);
                        const float* v = p.vecs + (size_t)entry * p.ldv;
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (lane + 32 * j < p.d) {
                                const float vv = __ldg(v + lane + 32 * j);
                                if (p.is_ip) acc = fmaf(qreg[j], vv, acc);
                                else { const float df = qreg[j] - vv; acc = fmaf(df, df, acc); }
                            }
                        }
                        for (int jj = lane + 256; jj < p.d; jj += 32) {   // d > 256: the rest of the query row from L1 / L2
                            const float vv = __ldg(v + jj), qq = __ldg(p.q + (size_t)q * p.ldq + jj);
                            if (p.is_ip) acc = fmaf(qq, vv, acc);
                            else { const float df = qq - vv; acc = fmaf(df, df, acc); }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        y = make_key(p.is_ip ? -acc : acc, (uint32_t)__ldg(p.list_ids + entry));
                    }
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
    }
    overflow = __any_sync(0xffffffffu, overflow);
    if (overflow) {
        if (lane == 0) { p.redo[q] = 1; atomicAdd(p.n_redo, 1); }
        return;
    }
    if (lane == 0) p.redo[q] = 0;
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
