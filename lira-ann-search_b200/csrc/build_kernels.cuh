// Build-side kernels (the callers either side of the query path): the K-Means update step, the statistics of
// the centroid-distance features a StandardScaler needs, and the redundancy rule of the probing model.
//   faiss.Kmeans (utils.py:321-330)            -> assignment = exact 1-NN against the centroid table on the kNN path
//                                                 (lira_knn_*), update = kmeans_accumulate_kernel + kmeans_finalize_kernel
//   StandardScaler.fit on get_dist_cid output  -> feature_stats_kernel (fp64 column sums of the [n, B] distance matrix,
//   (utils.py:120-215)                            which never leaves the device)
//   mul_partition_by_model                     -> mul_partition_kernel (top-n_mul scores of a row + the three-branch rule of
//   (LIRA_smallscale.py:77-97,                    LIRA_smallscale.py:79-97, one warp per point)
//    LIRA_largescale.py:51-72)
#pragma once
#include "common.cuh"

namespace lira {

// out[e, :] = base[rows[e], :] (int64 row ids), zero padded to ds columns
__global__ void gather_rows64_kernel(const float* __restrict__ base, long ldb, int d, const long long* __restrict__ rows, long long E,
                                     float* __restrict__ out, int ds) {
    const int per_row = ds / 4;
    const long long total = E * per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / per_row;
        const int c = (int)(i % per_row) * 4;
        const float* src = base + rows[e] * ldb + c;
        float4 v;
        v.x = c + 0 < d ? src[0] : 0.0f;
        v.y = c + 1 < d ? src[1] : 0.0f;
        v.z = c + 2 < d ? src[2] : 0.0f;
        v.w = c + 3 < d ? src[3] : 0.0f;
        *reinterpret_cast<float4*>(out + e * ds + c) = v;
    }
}

// sums[a, :] += x[i, :], counts[a] += 1 for a = assign[i]; one warp per row
__global__ void __launch_bounds__(256) kmeans_accumulate_kernel(const float* __restrict__ x, long ld, int d, long long n,
                                                                const long long* __restrict__ assign, float* __restrict__ sums, int lds,
                                                                int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long i = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const long long a = assign[i];
    if (a < 0) return;
    for (int j = lane; j < d; j += 32) atomicAdd(sums + a * lds + j, x[i * ld + j]);
    if (lane == 0) atomicAdd(counts + a, 1);
}

// centroid = sums / count where count > 0 (an empty cluster keeps its row: the host re-seeds it)
__global__ void kmeans_finalize_kernel(const float* __restrict__ sums, int lds, const int* __restrict__ counts, int B, int d,
                                       float* __restrict__ centroids, long ldc) {
    const long long total = (long long)B * d;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / d), j = (int)(t % d);
        const int c = counts[b];
        if (c > 0) centroids[b * ldc + j] = sums[(long long)b * lds + j] / (float)c;
    }
}

// sum_out[b] += sum_i f[i, b], sq_out[b] += sum_i f[i, b]^2 in fp64 (StandardScaler.partial_fit's accumulators); grid = (column
// blocks of 32, row slabs), block = 32 x 8
__global__ void __launch_bounds__(256) feature_stats_kernel(const float* __restrict__ f, long ldf, long long n, int B,
                                                            double* __restrict__ sum_out, double* __restrict__ sq_out) {
    __shared__ double s1[8][33], s2[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + tx;
    double a1 = 0.0, a2 = 0.0;
    if (b < B)
        for (long long i = blockIdx.y * 8ll + ty; i < n; i += (long long)gridDim.y * 8) {
            const double v = (double)f[i * ldf + b];
            a1 += v;
            a2 += v * v;
        }
    s1[ty][tx] = a1;
    s2[ty][tx] = a2;
    __syncthreads();
    if (ty == 0 && b < B) {
#pragma unroll
        for (int r = 1; r < 8; ++r) { a1 += s1[r][tx]; a2 += s2[r][tx]; }
        atomicAdd(sum_out + b, a1);
        atomicAdd(sq_out + b, a2);
    }
}

// The redundancy rule for one point per warp (LIRA_smallscale.py:79-97 / LIRA_largescale.py:53-72), n_mul <= 8:
//   partitions ranked by score, descending (equal scores: lower partition id first); n_eff = #(score > sigma);
//   n_act = min(n_mul - 1, n_eff); loc = rank of the point's current partition d2b[t, 0]
//     loc >= n_act              -> the n_act best go to columns 1 .. n_act (column 0 keeps the current partition)
//     else, n_eff == n_act      -> the n_act best replace columns 0 .. n_act - 1
//     else                      -> the n_act + 1 best replace columns 0 .. n_act
//   added[t, j] = partition newly holding t in column j (-1: none / the current one): the caller appends t to those lists.
// rows: score[i, :] belongs to point pts[i] (pts == null: point i + first).
__global__ void __launch_bounds__(256) mul_partition_kernel(const float* __restrict__ score, long lds, long long n_rows, int B, float sigma,
                                                            const long long* __restrict__ pts, long long first, int n_mul,
                                                            int* __restrict__ d2b, int* __restrict__ added) {
    const int lane = threadIdx.x & 31;
    const long long i = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_rows) return;
    const long long t = pts ? pts[i] : first + i;
    const float* s = score + i * lds;
    const int cur = d2b[t * n_mul];
    // top n_mul keys (descending score, ascending id) in a warp-distributed sorted list of 32 keys
    unsigned long long key[1] = {KEY_INF};
    unsigned long long kth = KEY_INF;
    int n_eff = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {
        const int b = b0 + lane;
        const float v = b < B ? s[b] : -INFINITY;
        n_eff += (b < B && v > sigma) ? 1 : 0;
        const unsigned long long x = b < B ? make_key(-v, (uint32_t)b) : KEY_INF;
        uint32_t mm = __ballot_sync(0xffffffffu, x < kth);
        while (mm) {
            const int sl = __ffs(mm) - 1;
            mm &= mm - 1;
            const unsigned long long y = shfl_u64(x, sl);
            if (y < kth) {
                warp_sorted_insert<1>(key, y, lane);
                kth = warp_sorted_get<1>(key, n_mul - 1);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_eff += __shfl_xor_sync(0xffffffffu, n_eff, o);
    const int n_act = min(n_mul - 1, n_eff);
    const int mine = (lane < n_mul && key[0] != KEY_INF) ? (int)key_pos(key[0]) : -1;   // lane j holds the j-th best partition
    const uint32_t hit = __ballot_sync(0xffffffffu, lane < n_mul && mine == cur);
    const int loc = hit ? __ffs(hit) - 1 : n_mul;     // rank of the current partition among the n_mul best (n_mul: not among them)
    if (lane < n_mul) {
        int col = -1;   // column this lane's partition goes to
        if (loc >= n_act) { if (lane < n_act) col = lane + 1; }
        else if (n_eff == n_act) { if (lane < n_act) col = lane; }
        else { if (lane <= n_act) col = lane; }
        if (col >= 0 && col < n_mul) {
            d2b[t * n_mul + col] = mine;
            added[t * n_mul + col] = (mine != cur) ? mine : -1;
        }
    }
}

}  // namespace lira
