// Tile engine shared by the list scan (K2), the centroid-feature / probing-MLP GEMMs (K0/K1)
// and the SIMT exact-kNN path: a warp-specialised, mbarrier-pipelined fp32 "pairwise" tile
//
//      out[q, v] = sum_k op(A[q, k], B[v, k])        op = (a-b)^2  (L2)   or   a*b  (IP / linear)
//
// B rows (inverted-list vectors, centroids, weight rows) are contiguous row-major [rows, ld]
// in HBM and are brought in by TMA as 128-row x 32-float boxes with the hardware 128-byte
// swizzle; A rows (queries of one probe group, gathered by index; activations) are brought in
// by the producer warp with 16-byte cp.async into the same swizzled layout. One producer warp
// feeds NSTAGE stages; 8 consumer warps (256 threads) hold a TM x 128 accumulator tile in
// registers (RQ x RV per thread) and read operands with conflict-free 128-bit LDS.
//
// Thread layout inside a warp is fixed at 4 (query direction) x 8 (vector direction) lanes:
// one LDS.128 touches 4 (A) or 8 (B) distinct rows whose swizzled 16-byte chunks fall in
// distinct bank groups, the rest is broadcast -> 1 shared-memory wavefront per load.
#pragma once
#include "common.cuh"

namespace lira {

static constexpr int TN = 128;            // B rows (vectors) per tile
static constexpr int KC = 32;             // floats per K step = one 128-byte swizzle row
static constexpr int ROW_BYTES = KC * 4;  // 128
static constexpr int B_STAGE_BYTES = TN * ROW_BYTES;  // 16 KiB
static constexpr int N_CONSUMER_WARPS = 8;
static constexpr int N_CONSUMERS = N_CONSUMER_WARPS * 32;
static constexpr int N_THREADS = N_CONSUMERS + 32;  // + producer warp
static constexpr int LQ = 4, LV = 8;

enum { OP_L2 = 0, OP_DOT = 1 };

template <int TM_>
struct TileCfg;
template <> struct TileCfg<128> { static constexpr int TM = 128, RQ = 8, RV = 8, WQ = 4, WV = 2; };
template <> struct TileCfg<64>  { static constexpr int TM = 64,  RQ = 8, RV = 4, WQ = 2, WV = 4; };
template <> struct TileCfg<32>  { static constexpr int TM = 32,  RQ = 8, RV = 2, WQ = 1, WV = 8; };
template <> struct TileCfg<16>  { static constexpr int TM = 16,  RQ = 4, RV = 2, WQ = 1, WV = 8; };
template <> struct TileCfg<8>   { static constexpr int TM = 8,   RQ = 2, RV = 2, WQ = 1, WV = 8; };

template <int TM>
__host__ __device__ constexpr int stage_bytes() { return TM * ROW_BYTES + B_STAGE_BYTES; }

struct PipeState {
    int stage;
    uint32_t phase;
    __device__ __forceinline__ void advance(int nstage) {
        if (++stage == nstage) { stage = 0; phase ^= 1u; }
    }
};

// ------------------------------------------------------------------------------------------
// producer side: one K step of one tile (called by all 32 lanes of the producer warp)
//   arow[t]  : global row index of A row (lane/8 + 4 t), or -1 (zero-fill)
// ------------------------------------------------------------------------------------------
template <int TM>
__device__ __forceinline__ void produce_kstep(uint8_t* stage_base, uint64_t* full_bar, uint64_t* empty_bar,
                                              PipeState& ps, int nstage, const CUtensorMap* tmap_b, int b_row0,
                                              const float* __restrict__ a_base, long lda, int kdim,
                                              const int (&arow)[TM / 4 > 0 ? TM / 4 : 1], int kc, int lane) {
    mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1u);
    uint8_t* sA = stage_base + (size_t)ps.stage * stage_bytes<TM>();
    uint8_t* sB = sA + TM * ROW_BYTES;
    if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[ps.stage], B_STAGE_BYTES);
        tma_load_2d(sB, tmap_b, kc * KC, b_row0, &full_bar[ps.stage]);
    }
    const int c = lane & 7;
    const int col = kc * KC + c * 4;
    const bool col_ok = col < kdim;  // kdim % 4 == 0, so a 16-byte chunk is all-in or all-out
    const uint32_t sA_u32 = smem_u32(sA);
#pragma unroll
    for (int t = 0; t < TM / 4; ++t) {
        if (t * 4 + (lane >> 3) < TM) {
            const int r = (lane >> 3) + 4 * t;
            const int g = arow[t];
            const bool ok = col_ok && g >= 0;
            const float* src = ok ? (a_base + (long)g * lda + col) : a_base;
            cp_async_16(sA_u32 + r * ROW_BYTES + ((c ^ (r & 7)) << 4), src, ok ? 16u : 0u);
        }
    }
    cp_async_mbar_arrive_noinc(&full_bar[ps.stage]);
    ps.advance(nstage);
}

// ------------------------------------------------------------------------------------------
// consumer side: accumulate one K step from a landed stage
// ------------------------------------------------------------------------------------------
template <int TM, int OP>
__device__ __forceinline__ void consume_kstep(const uint8_t* sA, const uint8_t* sB,
                                              float (&acc)[TileCfg<TM>::RQ][TileCfg<TM>::RV], int q0, int v0) {
    using C = TileCfg<TM>;
    // q rows: q0 + 4 i (q0 = warp base + lq), v rows: v0 + 8 j (v0 = warp base + lv)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 a[C::RQ], b[C::RV];
#pragma unroll
        for (int i = 0; i < C::RQ; ++i) {
            const int r = q0 + 4 * i;
            a[i] = *reinterpret_cast<const float4*>(sA + r * ROW_BYTES + ((c ^ (r & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < C::RV; ++j) {
            const int r = v0 + 8 * j;
            b[j] = *reinterpret_cast<const float4*>(sB + r * ROW_BYTES + ((c ^ (r & 7)) << 4));
        }
#pragma unroll
        for (int i = 0; i < C::RQ; ++i) {
#pragma unroll
            for (int j = 0; j < C::RV; ++j) {
                if (OP == OP_L2) {
                    // direct difference, the arithmetic of the reference's scan
                    // (search.cpp:253-260 l2_sq; Faiss fvec_L2sqr for nx == 1)
                    float d0 = a[i].x - b[j].x, d1 = a[i].y - b[j].y, d2 = a[i].z - b[j].z, d3 = a[i].w - b[j].w;
                    acc[i][j] = fmaf(d0, d0, acc[i][j]);
                    acc[i][j] = fmaf(d1, d1, acc[i][j]);
                    acc[i][j] = fmaf(d2, d2, acc[i][j]);
                    acc[i][j] = fmaf(d3, d3, acc[i][j]);
                } else {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
            }
        }
    }
}

// One TM x 128 tile over the whole K extent (nk K steps); consumer warps only.
template <int TM, int OP>
__device__ __forceinline__ void consume_tile(uint8_t* stage_base, uint64_t* full_bar, uint64_t* empty_bar,
                                             PipeState& ps, int nstage, int nk,
                                             float (&acc)[TileCfg<TM>::RQ][TileCfg<TM>::RV], int q0, int v0, int lane) {
    using C = TileCfg<TM>;
#pragma unroll
    for (int i = 0; i < C::RQ; ++i)
#pragma unroll
        for (int j = 0; j < C::RV; ++j) acc[i][j] = 0.0f;
    for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full_bar[ps.stage], ps.phase);
        const uint8_t* sA = stage_base + (size_t)ps.stage * stage_bytes<TM>();
        consume_kstep<TM, OP>(sA, sA + TM * ROW_BYTES, acc, q0, v0);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[ps.stage]);
        ps.advance(nstage);
    }
}

// thread coordinates of a consumer thread inside the TM x 128 tile
template <int TM>
__device__ __forceinline__ void consumer_coords(int warp, int lane, int& q0, int& v0) {
    using C = TileCfg<TM>;
    const int wq = warp / C::WV, wv = warp % C::WV;
    q0 = wq * (LQ * C::RQ) + (lane >> 3);
    v0 = wv * (LV * C::RV) + (lane & 7);
}

// Shared-memory staging tile for epilogues: [TM][DT_LD] fp32. DT_LD % 32 == 8 makes the
// 4-row x 8-column footprint of one warp store conflict-free.
static constexpr int DT_LD = TN + 8;

template <int TM>
__device__ __forceinline__ void store_acc_to_dt(float* dt, const float (&acc)[TileCfg<TM>::RQ][TileCfg<TM>::RV],
                                                int q0, int v0, float sign) {
    using C = TileCfg<TM>;
#pragma unroll
    for (int i = 0; i < C::RQ; ++i)
#pragma unroll
        for (int j = 0; j < C::RV; ++j) dt[(q0 + 4 * i) * DT_LD + v0 + 8 * j] = sign * acc[i][j];
}

__device__ __forceinline__ void pipe_init(uint64_t* full_bar, uint64_t* empty_bar, int nstage, int tid) {
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full_bar[s], 1 + 32);              // expect_tx arrive + 32 cp.async noinc arrives
            mbar_init(&empty_bar[s], N_CONSUMER_WARPS);   // one elected lane per consumer warp
        }
        mbar_fence_init();
    }
}

}  // namespace lira
