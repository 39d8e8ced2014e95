// The query-phase front end around the probing model, in three launches instead of twelve (search.cpp:421-466 per batch):
//
//   prep_queries_kernel   one pass over the query rows: centred hi/lo split + |q'|^2 (centroid-distance GEMM), raw hi/lo
//                         split (vector_net), |q|^2 + exactness check (tensor-core scan), and the per-batch resets
//                         (probe counters, histogram, bounds) that used to be memsets and fill kernels;
//   [tc_dense_kernel<TD_EPI_SELECT>: the last Linear layer + sigmoid + threshold selection, tc_dense_kernels.cuh]
//   finish_select_kernel  one CTA: truncation / argmax fix-up of the per-query counts, both prefix sums (probe offsets
//                         over the queries, group offsets over the partitions) and the work items of the scan;
//   scatter_queries_kernel  one warp per query: inverts its probe list into the per-partition query groups AND writes
//                         the query's fp16 row into every group slot (the A operand of the tensor-core scan), so the
//                         separate gather pass and its indirection are gone; also nprobe[q] and cmp[q].
#pragma once
#include "probe_kernels.cuh"
#include "tc_dense_kernels.cuh"

namespace lira {

struct PrepParams {
    const float* q;          // [Q, ldq]
    long ldq;
    int d, ds;               // valid columns, row stride of the split outputs (zero padded)
    long long Q;
    const float* mu;         // [ds] centroid mean
    float *qch, *qcl, *qn;   // centred split [Q, ds] x 2, |q'|^2 [Q]
    float *qrh, *qrl;        // raw split [Q, ds] x 2
    float* qnorm;            // [Q] |q|^2
    int* inexact_flag;       // set to 1 when a value is not an integer of <= 11 bits or |q|^2 >= 2^22 (null: not checked)
    uint32_t* thr;           // [Q] <- f32_to_ordered(+inf)
    int* nsel;               // [Q] <- 0
    unsigned long long* rowbest;   // [Q] <- 0
    int* list_count;         // [B + 1] <- 0
    int* cursor;             // [B + 1] <- 0
    int B;
};

__global__ void __launch_bounds__(256) prep_queries_kernel(const PrepParams p) {
    const int lane = threadIdx.x & 31;
    const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long i = gt; i <= p.B; i += (long long)gridDim.x * blockDim.x) {
        p.list_count[i] = 0;
        p.cursor[i] = 0;
    }
    const long long m = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= p.Q) return;
    float s = 0.0f, sr = 0.0f;
    bool bad = false;
    for (int k = lane; k < p.ds; k += 32) {
        float x = 0.0f;
        if (k < p.d) x = p.q[m * p.ldq + k];
        const float v = k < p.d ? x - p.mu[k] : 0.0f;
        const float h = tf32_hi(v);
        p.qch[m * p.ds + k] = h;
        p.qcl[m * p.ds + k] = v - h;
        const float hr = tf32_hi(x);
        p.qrh[m * p.ds + k] = hr;
        p.qrl[m * p.ds + k] = x - hr;
        s = fmaf(v, v, s);
        sr = fmaf(x, x, sr);
        bad |= (x != rintf(x)) | !(fabsf(x) <= 2047.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
    }
    bad |= !(sr < 4194304.0f);
    if (p.inexact_flag && bad) *p.inexact_flag = 1;
    if (lane == 0) {
        p.qn[m] = s;
        p.qnorm[m] = sr;
        p.thr[m] = 0xFF800000u;
        p.nsel[m] = 0;
        p.rowbest[m] = 0ull;
    }
}

// block-wide exclusive scan of int counts into long long offsets (n + 1 entries); every thread of the 1024-thread CTA calls it
// `raw` is a plain load (all of a thread's loads are issued together), `in` turns the loaded value into the count (it may have
// side effects: stores and atomics, which would otherwise serialise the loads into one global round trip per element).
// Element layout: a warp owns 256 consecutive elements of a pass and lane l takes elements l, l + 32, ..: every load and store
// instruction of a warp touches ONE contiguous run (8 consecutive elements per THREAD made each instruction touch 32 sectors
// in 8-16 cache lines, and the single CTA spent its time queueing in the load/store unit). Counts per warp pass stay below 2^31.
template <class Raw, class Load>
__device__ __forceinline__ void block_exclusive_scan_1024(Raw raw, Load in, long long* out, int n, long long* warp_sum, long long* carry_s) {
    constexpr int PER = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024 * PER) {
        const int i0 = base + warp * (32 * PER) + lane;   // element j of this thread: i0 + 32 j
        int v[PER], inc[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] = (i0 + 32 * j < n) ? raw(i0 + 32 * j) : 0;
        int run = 0;   // elements of this warp before row j
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = (i0 + 32 * j < n) ? in(i0 + 32 * j, v[j]) : 0;
            int x = v[j];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            inc[j] = run + x;   // inclusive prefix inside the warp's 256 elements
            run += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 31) warp_sum[warp] = run;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const long long carry = *carry_s;
        const long long before = carry + (warp ? warp_sum[warp - 1] : 0);
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (i0 + 32 * j < n) out[i0 + 32 * j] = before + (inc[j] - v[j]);
        __syncthreads();
        if (threadIdx.x == 1023) *carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = *carry_s;
    __syncthreads();
}

struct FinishSelectParams {
    int* nsel;                      // [Q] raw counts from the selection epilogue -> clamped counts
    int* sel;                       // [Q, cap]
    int cap;
    int mode;                       // 1: a query with no selection takes its argmax (search.cpp:456-466)
    const unsigned long long* rowbest;
    int* list_count;                // [B]
    int Q, B;
    long long* probe_offsets;       // [Q + 1]
    long long* group_offsets;       // [B + 1]
    int* trunc_flag;                // set when a query selected more than cap partitions; trunc_flag[1] <- the largest such count
    const int* list_order;          // [B] lists by size, descending
    const long long* list_offsets;  // [B + 1]
    int tile;
    int seg_rows;                   // > 0 (byte scan): one item per (query tile group, segment of seg_rows list rows)
    ScanItem* items;
    int* n_items;
    unsigned long long* stats;      // {E_p, pairs}
};

__global__ void __launch_bounds__(1024) finish_select_kernel(const FinishSelectParams p) {
    pdl_wait();
    pdl_launch();
    __shared__ long long warp_sum[32];
    __shared__ long long carry_s;
    __shared__ int iwarp_sum[32];
    __shared__ int icarry_s;
    __shared__ unsigned long long stats_s[2];
    // The prefix over the queries and the partition side (prefix over the partitions + work items) are independent as long as
    // no query needs the argmax fix-up (mode 0): the launch then has TWO CTAs, one per half; mode 1 runs both in one CTA.
    const bool do_queries = gridDim.x == 1 || blockIdx.x == 0, do_lists = gridDim.x == 1 || blockIdx.x == 1;
    // probe offsets over the queries; the loader clamps a query's count to the cap (flagging the truncation) and gives a query
    // without any selection its argmax (mode 1, search.cpp:456-466), which also enters the partition histogram
    if (do_queries)
    block_exclusive_scan_1024([&](int q) { return __ldcg(p.nsel + q); }, [&](int q, int n) {
        if (n > p.cap) { atomicMax(p.trunc_flag + 1, n); n = p.cap; *p.trunc_flag = 1; }
        if (n == 0 && p.mode == 1) {
            const int b = (int)(0xFFFFFFFFu - (uint32_t)(p.rowbest[q] & 0xFFFFFFFFull));
            p.sel[(size_t)q * p.cap] = b;
            atomicAdd(p.list_count + b, 1);
            n = 1;
        }
        p.nsel[q] = n;
        return n;
    }, p.probe_offsets, p.Q, warp_sum, &carry_s);
    if (!do_lists) return;
    block_exclusive_scan_1024([&](int b) { return p.list_count[b]; }, [&](int, int c) { return c; }, p.group_offsets, p.B, warp_sum, &carry_s);
    // work items: lists in size order, each list's group cut into tiles of `tile` queries (build_items_kernel's arithmetic)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { icarry_s = 0; stats_s[0] = 0; stats_s[1] = 0; }
    __syncthreads();
    for (int base = 0; base < p.B; base += 1024) {
        const int i = base + threadIdx.x;
        int b = 0, g = 0, cnt = 0, nseg = 1;
        unsigned long long st_rows = 0, st_pairs = 0;
        if (i < p.B) {
            b = p.list_order[i];
            g = p.list_count[b];   // (= group_offsets[b + 1] - group_offsets[b]: one 4-byte gather instead of two 8-byte ones)
            cnt = (g + p.tile - 1) / p.tile;
            const unsigned long long nb = (unsigned long long)(p.list_offsets[b + 1] - p.list_offsets[b]);
            if (g > 0) { st_rows = nb; st_pairs = nb * (unsigned long long)g; }
            if (p.seg_rows > 0) { nseg = (int)((nb + p.seg_rows - 1) / p.seg_rows); if (nseg < 1) nseg = 1; cnt *= nseg; }
        }
        // {E_p, pairs}: one shared-memory atomic per warp (1024 threads adding to ONE global address are 2048 serialised L2
        // operations: they were most of this kernel's time)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_rows += shfl_u64(st_rows, lane ^ o);
            st_pairs += shfl_u64(st_pairs, lane ^ o);
        }
        if (lane == 0 && (st_rows | st_pairs)) { atomicAdd(&stats_s[0], st_rows); atomicAdd(&stats_s[1], st_pairs); }
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) iwarp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = iwarp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            iwarp_sum[lane] = w;
        }
        __syncthreads();
        const int carry = icarry_s;
        const int at = carry + (warp ? iwarp_sum[warp - 1] : 0) + (x - cnt);
        if (i < p.B) {
            const int gb = (int)p.group_offsets[b];
            for (int t = 0; t < cnt; ++t) {
                const int tq = t / nseg;
                const int left = g - tq * p.tile;
                ScanItem it;
                it.list = b;
                it.q_begin = gb + tq * p.tile;
                it.q_count = left < p.tile ? left : p.tile;
                it.tm = p.seg_rows > 0 ? t % nseg : p.tile;
                p.items[at + t] = it;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) icarry_s = carry + iwarp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *p.n_items = icarry_s;
        if (p.stats) { atomicAdd(p.stats + 0, stats_s[0]); atomicAdd(p.stats + 1, stats_s[1]); }
    }
}

struct ScatterQueriesParams {
    const float* q;                 // [Q, ldq]
    long ldq;
    int ds;                         // valid (padded to 4) columns of q
    const int* sel;                 // [Q, cap]
    const int* nsel;                // [Q]
    int cap;
    const long long* probe_offsets; // [Q + 1]
    const long long* group_offsets; // [B + 1]
    const long long* list_offsets;  // [B + 1]
    int* cursor;                    // [B] zeroed
    int* group_queries;             // [P]
    int* probe_slot;                // [P]
    int* probe_ids;                 // [P] the probe sets as a CSR over probe_offsets (the exact redo of flagged queries reads it)
    __half* gq;                     // [P, d16] fp16(scale q) in group order
    int d16;
    float scale;
    int* cand_count;                // byte scan: [P] <- 0 for every valid pair slot (may be null)
    uint8_t* gq8;                   // byte-valued index: [P, d8] uint8(q) in group order instead of gq (null otherwise)
    int d8;
    int* bad_flag;                  // set when a scaled value does not fit fp16 / a value is not an integer in [0, 255] (may be null)
    int* nprobe;                    // [Q] (may be null)
    long long* cmp;                 // [Q] (may be null)
    int Q;
};

__global__ void __launch_bounds__(256, 8) scatter_queries_kernel(const ScatterQueriesParams p) {
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    // (the first 32 probe slots of the row are requested together with the count: one global round trip less on the
    //  chain count -> list id -> returning atomic -> stores that a warp is waiting on most of the time)
    const int sel_first = lane < p.cap ? p.sel[(size_t)q * p.cap + lane] : 0;
    const int n = p.nsel[q];
    const long long po = p.probe_offsets[q];
    // this lane's part of the fp16 row: 4 values per lane and 128-column block (d <= 128: one block, kept in registers)
    const float* qr = p.q + (size_t)q * p.ldq;
    const int c0 = lane * 4;
    uint2 pk0 = make_uint2(0u, 0u);
    bool bad = false;
    auto pack = [&](int c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < p.ds) v = *reinterpret_cast<const float4*>(qr + c);
        bad |= !(fabsf(v.x * p.scale) <= 60000.f && fabsf(v.y * p.scale) <= 60000.f && fabsf(v.z * p.scale) <= 60000.f &&
                 fabsf(v.w * p.scale) <= 60000.f);
        __half2 h0 = __floats2half2_rn(v.x * p.scale, v.y * p.scale), h1 = __floats2half2_rn(v.z * p.scale, v.w * p.scale);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        return pk;
    };
    auto pack8 = [&](int c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < p.ds) v = *reinterpret_cast<const float4*>(qr + c);
        bad |= !(v.x >= 0.f && v.x <= 255.f && v.y >= 0.f && v.y <= 255.f && v.z >= 0.f && v.z <= 255.f && v.w >= 0.f && v.w <= 255.f);
        return (uint32_t)(int)v.x | ((uint32_t)(int)v.y << 8) | ((uint32_t)(int)v.z << 16) | ((uint32_t)(int)v.w << 24);
    };
    uint32_t b0 = 0, b1 = 0;   // this lane's 4 bytes of the first and second 128-byte block of the row
    if (p.gq8) {
        if (c0 < p.d8) b0 = pack8(c0);
        if (c0 + 128 < p.d8) b1 = pack8(c0 + 128);
    } else if (c0 < p.d16) pk0 = pack(c0);
    long long cmp = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        int pos = 0;
        if (j < n) {
            const int b = j0 == 0 ? sel_first : p.sel[(size_t)q * p.cap + j];
            pos = (int)p.group_offsets[b] + atomicAdd(p.cursor + b, 1);
            p.group_queries[pos] = q;
            if (p.cand_count) p.cand_count[pos] = 0;
            p.probe_slot[po + j] = pos;
            p.probe_ids[po + j] = b;
            cmp += p.list_offsets[b + 1] - p.list_offsets[b];
        }
        const int cnt = min(32, n - j0);
        for (int jj = 0; jj < cnt; ++jj) {
            const int ps = __shfl_sync(0xffffffffu, pos, jj);
            if (p.gq8) {
                uint8_t* dst8 = p.gq8 + (size_t)ps * p.d8;
                if (c0 < p.d8) *reinterpret_cast<uint32_t*>(dst8 + c0) = b0;
                if (c0 + 128 < p.d8) *reinterpret_cast<uint32_t*>(dst8 + c0 + 128) = b1;
                continue;
            }
            __half* dst = p.gq + (size_t)ps * p.d16;
            if (c0 < p.d16) *reinterpret_cast<uint2*>(dst + c0) = pk0;
            for (int c = c0 + 128; c < p.d16; c += 128) *reinterpret_cast<uint2*>(dst + c) = pack(c);
        }
    }
    if (p.bad_flag && bad) *p.bad_flag = 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmp += __shfl_xor_sync(0xffffffffu, cmp, o);
    if (lane == 0) {
        if (p.nprobe) p.nprobe[q] = n;
        if (p.cmp) p.cmp[q] = cmp;
    }
}

}  // namespace lira
