// K0 / MLP on the 5th-generation tensor cores: out[m, n] = epi( sum_k A[m, k] * B[n, k] ) with fp32 accuracy.
//
// tcgen05.mma kind::tf32 keeps 10 mantissa bits of each operand, which is not enough for the probing model
// (its scores are compared with thresholds; the reference evaluates it in fp32, model_probing.py:33-39).
// Every operand is therefore carried as an error-free pair  x = hi + lo,  hi = x with the low 13 mantissa
// bits cleared (exact in TF32), lo = x - hi (exact in fp32, |lo| < 2^-10 |x|), and a product is formed as
//     a * b  ~=  a_hi * b_hi + a_hi * b_lo + a_lo * b_hi          (three MMAs into one fp32 TMEM accumulator)
// the dropped lo * lo term and the rounding of the lo parts are O(2^-21) relative per product -- the same
// order as fp32 accumulation itself. Weights and centroids are split once at model creation; every layer's
// epilogue writes its activations already split, so the next layer's TMA loads feed the MMAs directly.
//
// Centroid distances use |q-c|^2 = |q'|^2 + |c'|^2 - 2 q'.c' on data centred by the centroid mean
// (q' = q - mu, c' = c - mu): centring removes the common offset that would otherwise make the expansion
// cancel catastrophically for SIFT-like data (norms 1e5-1e6, squared distances 1e4).
//
// One persistent CTA per SM; warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue (one TMEM
// lane quadrant and half of the columns each, eight warps). 128 x 128 output tiles, K blocks of 32 floats, 3-stage ring of {A_hi, A_lo, B_hi, B_lo}
// boxes (64 KiB per stage), two TMEM accumulators so the epilogue of a tile overlaps the MMAs of the next.
#pragma once
#include "tc_scan_kernels.cuh"

namespace lira {

enum { TD_EPI_FEATURE = 0, TD_EPI_BIAS_RELU = 1, TD_EPI_BIAS_SIGMOID = 2, TD_EPI_SELECT = 3 };

static constexpr int TD_THREADS = 320;   // warp 0 TMA, warp 1 TMEM + MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
static constexpr int TD_EPI_WARPS = 8;
static constexpr int TD_NSTAGE = 3;
static constexpr int TD_STAGE_BYTES = 4 * B_STAGE_BYTES;   // A_hi, A_lo, B_hi, B_lo: 128 rows x 128 B each
static constexpr size_t TD_SMEM_BYTES = (size_t)TD_NSTAGE * TD_STAGE_BYTES + 3 * TC_N * 4 + TD_EPI_WARPS * 32 * 33 * 4 + 256;

struct TdParams {
    int M, N, K;             // out is M x N, reduction over K (zero-filled past the end by TMA)
    float* out_hi;           // [M, ldo] split activations (may be null)
    float* out_lo;
    float* out;              // [M, ldo_f] plain fp32 result (may be null)
    long ldo, ldo_f;
    int col_off;             // column offset into out_hi / out_lo (concatenation of two branches)
    const float* v0;         // FEATURE: |c'|^2 [N]          BIAS_*: bias [N]
    const float* v1;         // FEATURE: scaler mean [N] (may be null)
    const float* v2;         // FEATURE: scaler scale [N]
    const float* rown;       // FEATURE: |q'|^2 [M]
    // SELECT (the last layer fused with the partition selection: the scores never leave the SM): a score that passes the
    // threshold appends its partition to the query's probe list (slot from atomicAdd on nsel[m]) and counts in the
    // per-partition histogram; entries past sel_cap are dropped (nsel keeps counting: the caller sees the truncation)
    int* sel;                // [M, sel_cap] selected partitions, in arrival order
    int* nsel;               // [M] zeroed before the launch
    int* list_count;         // [N] zeroed before the launch
    unsigned long long* rowbest;   // [M] zeroed: (score bits << 32 | ~partition) maximum = first argmax (sel_mode 1 only)
    int sel_cap;
    int sel_mode;            // 0: score > thr (LIRA_smallscale.py:206)   1: score >= thr, argmax recorded (search.cpp:448-466)
    float sel_thr;
};

// MUFU.SQRT: max relative error 2^-22 (the reference's own two feature paths differ by more: fp64 cdist vs fp32 loop)
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int EPI>
__global__ void __launch_bounds__(TD_THREADS, 1)
tc_dense_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
                const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl, const TdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* stages = smem_raw;
    float* vec_s = (float*)(stages + (size_t)TD_NSTAGE * TD_STAGE_BYTES);   // [3][128] per-column epilogue vectors
    float* tr_s = vec_s + 3 * TC_N;                   // [8 warps][32][33] transpose tiles of the epilogue
    uint64_t* full = (uint64_t*)(tr_s + TD_EPI_WARPS * 32 * 33);   // [TD_NSTAGE]
    uint64_t* empty = full + TD_NSTAGE;               // [TD_NSTAGE]
    uint64_t* t_full = empty + TD_NSTAGE;             // [2]
    uint64_t* t_empty = t_full + 2;                   // [2]
    uint32_t* tmem_slot = (uint32_t*)(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TD_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], TD_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { tma_prefetch_desc(&tm_ah); tma_prefetch_desc(&tm_al); tma_prefetch_desc(&tm_bh); tma_prefetch_desc(&tm_bl); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // (the prologue above touches nothing another kernel produces)
    pdl_launch();

    const int tiles_n = (p.N + TC_N - 1) / TC_N;
    const int tiles_m = (p.M + TC_M - 1) / TC_M;
    const int n_tiles = tiles_m * tiles_n;
    const int nk = (p.K + KC - 1) / KC;

    if (warp == 0) {
        // TMA producer: the whole warp runs the loop, one elected lane issues (see elect_one)
        PipeState ps{0, 0};
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int m0 = (t / tiles_n) * TC_M, n0 = (t % tiles_n) * TC_N;   // consecutive CTAs share the A rows
            for (int kb = 0; kb < nk; ++kb) {
                mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                uint8_t* s = stages + (size_t)ps.stage * TD_STAGE_BYTES;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[ps.stage], TD_STAGE_BYTES);
                    tma_load_2d(s, &tm_ah, kb * KC, m0, &full[ps.stage]);
                    tma_load_2d(s + B_STAGE_BYTES, &tm_al, kb * KC, m0, &full[ps.stage]);
                    tma_load_2d(s + 2 * B_STAGE_BYTES, &tm_bh, kb * KC, n0, &full[ps.stage]);
                    tma_load_2d(s + 3 * B_STAGE_BYTES, &tm_bl, kb * KC, n0, &full[ps.stage]);
                }
                __syncwarp();
                ps.advance(TD_NSTAGE);
            }
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform loop, one elected lane issues
        PipeState ps{0, 0};
        uint32_t it = 0;
        const uint32_t stages_u32 = smem_u32(stages);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1u;
            mbar_wait(&t_empty[acc], ((it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * TC_N;
            for (int kb = 0; kb < nk; ++kb) {
                mbar_wait(&full[ps.stage], ps.phase);
                tc_fence_after();
                const uint32_t s = stages_u32 + (uint32_t)ps.stage * TD_STAGE_BYTES;
                const uint32_t ah = s, al = s + B_STAGE_BYTES, bh = s + 2 * B_STAGE_BYTES, bl = s + 3 * B_STAGE_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {   // small terms first, then the leading one
                        tc_mma_tf32(d_tmem, tc_smem_desc(al + j * 32), tc_smem_desc(bh + j * 32), TC_IDESC_TF32, (kb | j) ? 1u : 0u);
                        tc_mma_tf32(d_tmem, tc_smem_desc(ah + j * 32), tc_smem_desc(bl + j * 32), TC_IDESC_TF32, 1u);
                        tc_mma_tf32(d_tmem, tc_smem_desc(ah + j * 32), tc_smem_desc(bh + j * 32), TC_IDESC_TF32, 1u);
                    }
                    tc_commit(&empty[ps.stage]);
                }
                __syncwarp();
                ps.advance(TD_NSTAGE);
            }
            if (elect_one()) tc_commit(&t_full[acc]);
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..9 -> TMEM lane quadrants 2, 3, 0, 1, 2, 3, 0, 1; warps 2-5 take columns 0-63, 6-9 columns 64-127 =====
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;   // 0..255
        const int half = (warp - 2) >> 2;
        float* tr = tr_s + (warp - 2) * 32 * 33;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1u;
            const int m0 = (t / tiles_n) * TC_M, n0 = (t % tiles_n) * TC_N;
            // per-column vectors of this tile -> shared (previous tile's readers are past their last read: bar below)
            named_bar_sync(2, TD_EPI_WARPS * 32);
            if (et < TC_N) {
                const int n = n0 + et;
                const bool ok = n < p.N;
                vec_s[et] = (ok && p.v0) ? __ldg(p.v0 + n) : 0.0f;
                if (EPI == TD_EPI_FEATURE) {
                    vec_s[TC_N + et] = (ok && p.v1) ? __ldg(p.v1 + n) : 0.0f;
                    float sc = (ok && p.v2) ? __ldg(p.v2 + n) : 1.0f;
                    if (sc == 0.0f) sc = 1.0f;   // search.cpp:246
                    vec_s[2 * TC_N + et] = 1.0f / sc;   // the epilogue multiplies (1 ulp from the reference's division)
                }
            }
            named_bar_sync(2, TD_EPI_WARPS * 32);
            const int m = m0 + row;
            float rn = 0.0f;
            if (EPI == TD_EPI_FEATURE && m < p.M) rn = __ldg(p.rown + m);
            mbar_wait(&t_full[acc], (it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TC_N;
            if constexpr (EPI == TD_EPI_SELECT) {
                const bool row_ok = m < p.M;
                float best = -1.0f;
                int best_n = 0;
#pragma unroll 1
                for (int g = half * 2; g < half * 2 + 2; ++g) {
                    uint32_t r[32];
                    tc_ld32_async(taddr + g * 32, r);
                    tc_ld_wait(r);
                    // pass 1 (ALU only): which of the 32 columns pass; pass 2: ONE returning atomic reserves the thread's slots in the
                    // query's probe list (a returning atomic per hit would put a global round trip on every hit of the warp)
                    uint32_t mask = 0;
#pragma unroll
                    for (int u = 0; u < 32; ++u) {
                        const int c = g * 32 + u;
                        const int n = n0 + c;
                        const float v = __fdividef(1.0f, 1.0f + __expf(-(__uint_as_float(r[u]) + vec_s[c])));   // same sigmoid as TD_EPI_BIAS_SIGMOID
                        const bool ok = row_ok && n < p.N;
                        if (p.sel_mode == 1 && ok && v > best) { best = v; best_n = n; }
                        const bool hit = ok && (p.sel_mode == 0 ? v > p.sel_thr : v >= p.sel_thr);
                        mask |= hit ? (1u << u) : 0u;
                    }
                    if (mask) {
                        int pos = atomicAdd(p.nsel + m, __popc(mask));
                        while (mask) {
                            const int n = n0 + g * 32 + __ffs(mask) - 1;
                            mask &= mask - 1;
                            if (pos < p.sel_cap) {
                                p.sel[(size_t)m * p.sel_cap + pos] = n;
                                atomicAdd(p.list_count + n, 1);
                            }
                            ++pos;
                        }
                    }
                }
                if (p.sel_mode == 1 && row_ok && best >= 0.0f)
                    atomicMax(p.rowbest + m, ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_n));
            } else {
#pragma unroll 1
            for (int g = half * 2; g < half * 2 + 2; ++g) {
                uint32_t r[32];
                tc_ld32_async(taddr + g * 32, r);
                tc_ld_wait(r);
                float y[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const int c = g * 32 + u;
                    const float a = __uint_as_float(r[u]);
                    if (EPI == TD_EPI_FEATURE) {
                        // utils.py:98-118 (Euclidean distance, with sqrt) + StandardScaler.transform (:142-167)
                        const float d2 = fmaf(-2.0f, a, rn + vec_s[c]);
                        float v = fast_sqrt(fmaxf(d2, 0.0f));
                        if (p.v1) v = (v - vec_s[TC_N + c]) * vec_s[2 * TC_N + c];
                        y[u] = v;
                    } else if (EPI == TD_EPI_BIAS_RELU) {
                        y[u] = fmaxf(a + vec_s[c], 0.0f);
                    } else {
                        y[u] = __fdividef(1.0f, 1.0f + __expf(-(a + vec_s[c])));   // sigmoid, ~2 ulp
                    }
                }
                // transpose the warp's 32 x 32 block through shared memory so that every store instruction writes
                // 128 contiguous bytes of one output row (a thread owns a ROW in TMEM: direct stores would touch
                // 32 rows per instruction, half a sector each)
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 32; ++u) tr[lane * 33 + u] = y[u];
                __syncwarp();
                const int n = n0 + g * 32 + lane;
                if (n < p.N) {
                    const int rows = min(32, p.M - (m0 + quad * 32));
                    for (int rr = 0; rr < rows; ++rr) {
                        const float v = tr[rr * 33 + lane];
                        const size_t mm = (size_t)(m0 + quad * 32 + rr);
                        if (p.out) p.out[mm * p.ldo_f + n] = v;
                        if (p.out_hi) {
                            const float h = tf32_hi(v);
                            p.out_hi[mm * p.ldo + p.col_off + n] = h;
                            p.out_lo[mm * p.ldo + p.col_off + n] = v - h;
                        }
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// x' = x - mu (mu may be null); hi / lo split of x' and |x'|^2 per row (fp32, sequential order). One warp per row.
__global__ void split_rows_kernel(const float* __restrict__ x, long ld, int K, long long M, const float* __restrict__ mu,
                                  float* __restrict__ hi, float* __restrict__ lo, long ldo, float* __restrict__ norm) {
    const int lane = threadIdx.x & 31;
    const long long m = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    float s = 0.0f;
    for (int k = lane; k < ldo; k += 32) {
        float v = 0.0f;
        if (k < K) v = x[m * ld + k] - (mu ? mu[k] : 0.0f);
        const float h = tf32_hi(v);
        hi[m * ldo + k] = h;
        lo[m * ldo + k] = v - h;
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && norm) norm[m] = s;
}

}  // namespace lira
