// K0 / K1: probing path.
//   K0  centroid-distance features + standardisation (utils.py:98-180 query half; search.cpp:220-250)
//   MLP the six Linear layers of MLP_2_Input (model_probing.py:12-39) as fp32 tile GEMMs with fused
//       bias+ReLU / bias+sigmoid epilogues (weights are nn.Linear layout [out, in] = K-contiguous rows,
//       exactly the B operand of the tile engine)
//   K1  partition selection (threshold ">" as LIRA_smallscale.py:206, ">=" + argmax fallback as
//       search.cpp:448-466, top-nprobe as utils.py:512), inversion of the probe sets into per-list
//       query groups, and work-item generation for the scan.
#pragma once
#include "scan_kernels.cuh"

namespace lira {

enum { EPI_FEATURE = 0, EPI_BIAS_RELU = 1, EPI_BIAS_SIGMOID = 2, EPI_BIAS = 3, EPI_NONE = 4 };

struct DenseParams {
    const float* a;  // [M, lda]
    long lda;
    int M, K, N;     // out[m, n] = epi(sum_k op(a[m,k], b[n,k]));  K % 4 == 0
    float* out;      // [M, ldo], written at column col_off + n
    long ldo;
    int col_off;
    const float* v0;  // EPI_FEATURE: mean[N] (may be null)   EPI_BIAS_*: bias[N]
    const float* v1;  // EPI_FEATURE: scale[N]
};

static constexpr size_t DENSE_SMEM_BYTES =
    1024 + (size_t)SCAN_NSTAGE * SCAN_STAGE_BYTES + (size_t)SCAN_TM_MAX * DT_LD * 4 + 2 * SCAN_NSTAGE * 8 + 64;

template <int TM, int OP, int EPI>
__global__ void __launch_bounds__(N_THREADS, 1)
dense_tile_kernel(const __grid_constant__ CUtensorMap tmap_b, const DenseParams p) {
    // dynamic shared memory is the only shared allocation of this kernel: the declared alignment holds
    // (the 128-byte TMA swizzle needs 1024-byte aligned stages; checked below). Plain pointer arithmetic on
    // the array keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST).
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* stages = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    float* dt = (float*)(stages + (size_t)SCAN_NSTAGE * SCAN_STAGE_BYTES);
    uint64_t* full_bar = (uint64_t*)(dt + SCAN_TM_MAX * DT_LD);
    uint64_t* empty_bar = full_bar + SCAN_NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pipe_init(full_bar, empty_bar, SCAN_NSTAGE, threadIdx.x);
    if (threadIdx.x == 32) tma_prefetch_desc(&tmap_b);
    __syncthreads();

    const int n0 = blockIdx.x * TN;
    const int m0 = blockIdx.y * TM;
    const int nk = (p.K + KC - 1) / KC;
    PipeState ps{0, 0};
    if (warp == N_CONSUMER_WARPS) {
        int arow[TM / 4];
#pragma unroll
        for (int t = 0; t < TM / 4; ++t) {
            const int r = m0 + (lane >> 3) + 4 * t;
            arow[t] = r < p.M ? r : -1;
        }
        for (int kc = 0; kc < nk; ++kc)
            scan_produce_kstep<TM>(stages, full_bar, empty_bar, ps, &tmap_b, n0, p.a, p.lda, p.K, arow, kc, lane);
    } else {
        using C = TileCfg<TM>;
        int q0, v0;
        consumer_coords<TM>(warp, lane, q0, v0);
        float acc[C::RQ][C::RV];
        scan_consume_tile<TM, OP>(stages, full_bar, empty_bar, ps, nk, acc, q0, v0, lane);
        store_acc_to_dt<TM>(dt, acc, q0, v0, 1.0f);
        named_bar_sync(1, N_CONSUMERS);
        for (int r = warp; r < TM; r += N_CONSUMER_WARPS) {
            const int m = m0 + r;
            if (m >= p.M) break;
            const float4 v = *reinterpret_cast<const float4*>(dt + r * DT_LD + lane * 4);
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + lane * 4 + j;
                if (n < p.N) {
                    float y = x[j];
                    if (EPI == EPI_FEATURE) {
                        y = sqrtf(y);
                        if (p.v0) {
                            float s = p.v1[n];
                            if (s == 0.0f) s = 1.0f;  // search.cpp:246
                            y = (y - p.v0[n]) / s;
                        }
                    } else if (EPI == EPI_BIAS_RELU) {
                        y = fmaxf(y + p.v0[n], 0.0f);
                    } else if (EPI == EPI_BIAS_SIGMOID) {
                        y = 1.0f / (1.0f + expf(-(y + p.v0[n])));
                    } else if (EPI == EPI_BIAS) {
                        y = y + p.v0[n];
                    }
                    p.out[(size_t)m * p.ldo + p.col_off + n] = y;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1: selection. One warp per query over scores[Q, lds] (probabilities in (0,1)).
// ------------------------------------------------------------------------------------------
enum { SEL_GT = 0, SEL_GE_ARGMAX = 1, SEL_TOPN = 2, SEL_ALL = 3 };

struct SelectParams {
    const float* scores;  // [Q, lds]
    long lds;
    int Q, B;
    int mode;
    double value;         // threshold or nprobe. Thresholds are compared in fp32, i.e. score > float(value): search.cpp's
                          // threshold is a float (search.cpp:413), and torch / numpy compare an fp32 array with a Python float in
                          // fp32 as well (LIRA_smallscale.py:206: float32(0.1) > 0.1 is False)
    const long long* list_offsets;  // [B+1] for the Computations column
    int* sel;             // [Q, B] selected partitions of each query, in output order
    int* nsel;            // [Q]
    long long* cmp;       // [Q] sum of probed list sizes (search.cpp:477, LIRA_smallscale.py:208)
    int* list_count;      // [B] histogram (atomic)
    int* seed_ids;        // [Q, 2] the two best-scoring selected partitions (-1 padded); may be null
    const int* mask;      // optional [Q]: queries with mask[q] == 0 select nothing
    int cap;              // > 0 (threshold modes): at most `cap` partitions per query are kept, so that Q * cap bounds the
    int* trunc_flag;      //   number of pairs without a host round trip; *trunc_flag = 1 when a query had more (caller reruns)
};

// the two smallest keys of a warp's per-lane (k1 <= k2) pairs, broadcast to all lanes
__device__ __forceinline__ void warp_two_smallest(unsigned long long k1, unsigned long long k2, unsigned long long& m1,
                                                  unsigned long long& m2) {
    unsigned long long a = k1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = shfl_u64(a, (threadIdx.x & 31) ^ o);
        a = t < a ? t : a;
    }
    m1 = a;
    unsigned long long b = (k1 == m1) ? k2 : k1;  // keys are unique (partition id in the low word)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = shfl_u64(b, (threadIdx.x & 31) ^ o);
        b = t < b ? t : b;
    }
    m2 = b;
}

template <int S>
__global__ void __launch_bounds__(256) select_kernel(const SelectParams p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const float* s = p.scores + (size_t)q * p.lds;
    int* sel = p.sel + (size_t)q * p.B;
    int n = 0;
    long long cmp = 0;
    if (p.mask && !p.mask[q]) {
        if (lane == 0) p.nsel[q] = 0;
        return;
    }
    if (p.mode == SEL_TOPN) {
        int want = (int)p.value;
        if (want > p.B) want = p.B;
        if (want > 32 * S) want = 32 * S;
        if (want < 1) want = 1;
        unsigned long long key[S];
#pragma unroll
        for (int t = 0; t < S; ++t) key[t] = KEY_INF;
        unsigned long long kth = KEY_INF;
        for (int b0 = 0; b0 < p.B; b0 += 32) {
            const int b = b0 + lane;
            const unsigned long long x = (b < p.B && want > 0) ? make_key(-s[b], (uint32_t)b) : KEY_INF;
            uint32_t mm = __ballot_sync(0xffffffffu, x < kth);
            while (mm) {
                const int sl = __ffs(mm) - 1;
                mm &= mm - 1;
                const unsigned long long y = shfl_u64(x, sl);
                if (y < kth) {
                    warp_sorted_insert<S>(key, y, lane);
                    kth = warp_sorted_get<S>(key, want - 1);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < S; ++t) {
            const int e = t * 32 + lane;
            if (e < want && key[t] != KEY_INF) {
                const int b = (int)key_pos(key[t]);
                sel[e] = b;
                atomicAdd(p.list_count + b, 1);
                cmp += p.list_offsets[b + 1] - p.list_offsets[b];
            }
        }
        n = want;
        if (p.seed_ids) {
            const unsigned long long b0 = shfl_u64(key[0], 0), b1 = shfl_u64(key[0], 1);
            if (lane == 0) {
                p.seed_ids[2 * q + 0] = (b0 == KEY_INF || want < 1) ? -1 : (int)key_pos(b0);
                p.seed_ids[2 * q + 1] = (b1 == KEY_INF || want < 2) ? -1 : (int)key_pos(b1);
            }
        }
    } else {
        float best = -INFINITY;
        int best_b = 0;
        const float thr = (float)p.value;
        unsigned long long h1 = KEY_INF, h2 = KEY_INF;  // this lane's two best selected partitions
        for (int base = 0; base < p.B; base += 1024) {
            // 32 independent coalesced loads in flight per lane (the loop below is compute only)
            float vv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int b = base + i * 32 + lane;
                vv[i] = b < p.B ? s[b] : -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int b = base + i * 32 + lane;
                if (base + i * 32 >= p.B) break;
                bool hit = false;
                const float v = vv[i];
                if (b < p.B) {
                    if (p.mode == SEL_GT) hit = v > thr;
                    else if (p.mode == SEL_GE_ARGMAX) hit = v >= thr;
                    else hit = true;
                    if (v > best) { best = v; best_b = b; }  // first maximum per lane (b ascending)
                }
                const uint32_t mm = __ballot_sync(0xffffffffu, hit);
                if (mm == 0) continue;
                if (hit && (p.cap <= 0 || n + __popc(mm & ((1u << lane) - 1u)) < p.cap)) {
                    sel[n + __popc(mm & ((1u << lane) - 1u))] = b;
                    atomicAdd(p.list_count + b, 1);
                    cmp += p.list_offsets[b + 1] - p.list_offsets[b];
                    const unsigned long long hk = make_key(-v, (uint32_t)b);
                    if (hk < h1) { h2 = h1; h1 = hk; } else if (hk < h2) { h2 = hk; }
                }
                n += __popc(mm);
            }
        }
        if (p.cap > 0 && n > p.cap) {
            n = p.cap;
            if (lane == 0 && p.trunc_flag) *p.trunc_flag = 1;
        }
        // global argmax, first maximum wins (search.cpp:456-466); it is always selected when anything is
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int obb = __shfl_xor_sync(0xffffffffu, best_b, o);
            if (ob > best || (ob == best && obb < best_b)) { best = ob; best_b = obb; }
        }
        if (p.seed_ids) {
            unsigned long long m1, m2;
            warp_two_smallest(h1, h2, m1, m2);
            if (lane == 0) {
                int s0 = m1 == KEY_INF ? -1 : (int)key_pos(m1);
                if (n == 0 && p.mode == SEL_GE_ARGMAX) s0 = best_b;
                p.seed_ids[2 * q + 0] = s0;
                p.seed_ids[2 * q + 1] = m2 == KEY_INF ? -1 : (int)key_pos(m2);
            }
        }
        if (p.mode == SEL_GE_ARGMAX && n == 0) {
            if (lane == 0) {
                sel[0] = best_b;
                atomicAdd(p.list_count + best_b, 1);
                cmp += p.list_offsets[best_b + 1] - p.list_offsets[best_b];
            }
            n = 1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmp += __shfl_xor_sync(0xffffffffu, cmp, o);
    if (lane == 0) {
        p.nsel[q] = n;
        if (p.cmp) p.cmp[q] = cmp;
    }
}

// exclusive scan of int counts into long long offsets (n+1 entries), single CTA of 1024 threads, 8 elements per
// thread and round
// (in1 / out1 / n1, optional: a second, independent scan done by block 1 of the same launch)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int* in, long long* out, int n, const int* in1 = nullptr,
                                                              long long* out1 = nullptr, int n1 = 0) {
    if (blockIdx.x == 1) { in = in1; out = out1; n = n1; }
    __shared__ long long warp_sum[32];
    __shared__ long long carry_s;
    constexpr int PER = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024 * PER) {
        const int i0 = base + threadIdx.x * PER;
        int v[PER];
        long long tsum = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = (i0 + j < n) ? in[i0 + j] : 0;
            tsum += v[j];
        }
        long long x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;  // inclusive
        }
        __syncthreads();
        const long long carry = carry_s;
        long long before = carry + (warp ? warp_sum[warp - 1] : 0) + (x - tsum);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (i0 + j < n) out[i0 + j] = before;
            before += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// invert the probe sets: slot of (q, j-th probe) inside its list's group
struct ScatterParams {
    const int* sel;                 // [Q, B]
    const int* nsel;                // [Q]
    const long long* probe_offsets; // [Q+1]
    const long long* group_offsets; // [B+1]
    int* cursor;                    // [B], zeroed
    int* group_queries;             // [P]
    int* probe_slot;                // [P]
    int Q, B;
};

__global__ void __launch_bounds__(256) scatter_groups_kernel(const ScatterParams p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int n = p.nsel[q];
    const long long po = p.probe_offsets[q];
    for (int j = lane; j < n; j += 32) {
        const int b = p.sel[(size_t)q * p.B + j];
        const int pos = (int)p.group_offsets[b] + atomicAdd(p.cursor + b, 1);
        p.group_queries[pos] = q;
        p.probe_slot[po + j] = pos;
    }
}

// work items: lists in `list_order` (size-descending), each list's group cut into tiles of 64 queries
// plus one remainder tile of class 8/16/32/64. Single CTA.
__device__ __forceinline__ int tile_class(int rem) { return rem > 32 ? 64 : rem > 16 ? 32 : rem > 8 ? 16 : 8; }

// stats[0] += entries in the union of probed lists (E_p), stats[1] += (query, vector) pairs -- the
// algorithmic work of SURVEY.md 8(d), reported by lira_index_last_timing.
__global__ void __launch_bounds__(1024) build_items_kernel(const int* list_order, const long long* group_offsets,
                                                           const long long* list_offsets, int B, int tile, ScanItem* items,
                                                           int* n_items_out, unsigned long long* stats, int seg_rows = 0) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int i = base + threadIdx.x;
        int b = 0, g = 0, cnt = 0, nseg = 1;
        if (i < B) {
            b = list_order[i];
            g = (int)(group_offsets[b + 1] - group_offsets[b]);
            // (a list with no vectors but a non-empty group still gets items: its all-INF rows must be written)
            cnt = (g + tile - 1) / tile;
            const unsigned long long nb = (unsigned long long)(list_offsets[b + 1] - list_offsets[b]);
            if (g > 0 && stats) {
                atomicAdd(stats + 0, nb);
                atomicAdd(stats + 1, nb * (unsigned long long)g);
            }
            // seg_rows > 0 (byte scan): one item per (query tile group, segment of seg_rows list rows); ScanItem::tm = segment
            if (seg_rows > 0) { nseg = (int)((nb + seg_rows - 1) / seg_rows); if (nseg < 1) nseg = 1; cnt *= nseg; }
        }
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        int at = carry + (warp ? warp_sum[warp - 1] : 0) + (x - cnt);
        if (i < B) {
            const int gb = (int)group_offsets[b];
            for (int t = 0; t < cnt; ++t) {
                const int tq = t / nseg;
                const int left = g - tq * tile;
                ScanItem it;
                it.list = b;
                it.q_begin = gb + tq * tile;
                it.q_count = left < tile ? left : tile;
                it.tm = seg_rows > 0 ? t % nseg : (tile > SCAN_TM_MAX ? tile : tile_class(it.q_count));
                items[at + t] = it;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_items_out = carry_s;
}

// ------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------
// list_vecs[e, :] = base[ids[e], :]  (utils.py:411-412 x_d[xd_id_bkt]; search.cpp:388-402), zero padded to ds
__global__ void gather_rows_kernel(const float* __restrict__ base, long ldb, int d, const int* __restrict__ ids,
                                   long long E, float* __restrict__ out, int ds) {
    const int per_row = ds / 4;
    const long long total = E * per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / per_row;
        const int c = (int)(i % per_row) * 4;
        const float* src = base + (long long)ids[e] * ldb + c;
        float4 v;
        v.x = c + 0 < d ? src[0] : 0.0f;
        v.y = c + 1 < d ? src[1] : 0.0f;
        v.z = c + 2 < d ? src[2] : 0.0f;
        v.w = c + 3 < d ? src[3] : 0.0f;
        *reinterpret_cast<float4*>(out + e * ds + c) = v;
    }
}

// get_cmp_recall layout (LIRA_smallscale.py:154-171): found[q, b, :k] = global ids best-first; a non-empty
// list shorter than k repeats its LAST id (xd_id_bid[-1] quirk, :169); empty lists stay -1 (:161); cmp = size.
__global__ void found_from_partials_kernel(const unsigned long long* part_key, const int* probe_slot,
                                           const long long* list_offsets, const int* list_ids, int Q, int B, int k,
                                           long long* found, long long* cmp) {
    const long long total = (long long)Q * B * k;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i % k);
        const long long qb = i / k;
        const int b = (int)(qb % B);
        const long long lo = list_offsets[b], hi = list_offsets[b + 1];
        const unsigned long long x = part_key[(size_t)probe_slot[qb] * k + e];
        long long id;
        if (hi == lo) id = -1;
        else if (x == KEY_INF) id = list_ids[hi - 1];
        else id = (long long)(int)key_pos(x);
        found[i] = id;
        if (e == 0 && cmp) cmp[qb] = hi - lo;
    }
}

// exhaustive probe sets (every query probes every list, in list order): the get_cmp_recall shape
__global__ void fill_all_pairs_kernel(int Q, int B, int* group_queries, int* probe_slot) {
    const long long total = (long long)Q * B;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i / B), b = (int)(i % B);
        const long long slot = (long long)b * Q + q;
        group_queries[slot] = q;
        probe_slot[i] = (int)slot;
    }
}

__global__ void iota_offsets_kernel(long long* out, long long n, long long step) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= n; i += (long long)gridDim.x * blockDim.x)
        out[i] = i * step;
}

}  // namespace lira
