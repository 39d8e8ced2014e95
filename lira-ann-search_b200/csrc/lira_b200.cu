// liblira_b200: host side of the C ABI declared in include/lira_b200.h.
// Owns device memory behind the handles, builds the TMA tensor maps, sequences the kernels of
// probe_kernels.cuh / scan_kernels.cuh on one stream per handle. No torch types, no CPU fallback.
#include "../../include/lira_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <thread>
#include <mutex>
#include <numeric>
#include <random>
#include <vector>

#include "probe_kernels.cuh"
#include "tc_scan_kernels.cuh"
#include "u8_scan_kernels.cuh"
#include "tc_dense_kernels.cuh"
#include "fused_probe_kernels.cuh"
#include "build_kernels.cuh"

namespace lira {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
static std::atomic<long long> g_launches{0};

#define LIRA_LAUNCH_CHECK()                                                                    \
    do {                                                                                       \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        LIRA_CUDA_OK(cudaGetLastError());                                                      \
    } while (0)

// Kernel launch with programmatic dependent launch (see pdl_wait in common.cuh): the kernel may be scheduled while the
// previous kernel of the stream drains; it calls pdl_wait() before it touches anything that kernel produces. Only used for
// kernels that do so on every thread. LIRA_NO_PDL=1 launches them fully serialised (A/B timing).
static bool pdl_enabled() {
    static const bool on = getenv("LIRA_NO_PDL") == nullptr;
    return on;
}
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int grid_for(long long n, int block, int cap = 148 * 8) {
    long long g = (n + block - 1) / block;
    return (int)std::max<long long>(1, std::min<long long>(g, cap));
}

// ---------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point resolved at run time: the .so has no link dependency on
// libcuda, so it loads -- and its symbols can be checked -- on a machine without a GPU)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(PFN_encodeTiled* out) {
    static PFN_encodeTiled fn = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LIRA_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        LIRA_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        fn = (PFN_encodeTiled)p;
    }
    *out = fn;
    return 0;
}

// rows x cols fp32 matrix with row stride ld (floats); box = 128 rows x 32 floats, 128-byte swizzle.
static int make_tmap(CUtensorMap* m, const float* base, long long rows, int cols, long long ld) {
    PFN_encodeTiled enc;
    if (int rc = get_encode_fn(&enc)) return rc;
    LIRA_REQUIRE(((uintptr_t)base & 15) == 0 && (ld % 4) == 0, "tensor map: base must be 16-byte aligned, ld % 4 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)std::max(cols, 1), (cuuint64_t)std::max<long long>(rows, 1)};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)TN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

// rows x cols fp16 matrix with row stride ld (halves); box = 128 rows x 64 halves, 128-byte swizzle (tensor-core scan).
static int make_tmap_f16(CUtensorMap* m, const void* base, long long rows, int cols, long long ld) {
    PFN_encodeTiled enc;
    if (int rc = get_encode_fn(&enc)) return rc;
    LIRA_REQUIRE(((uintptr_t)base & 15) == 0 && (ld % 8) == 0, "tensor map: base must be 16-byte aligned, ld % 8 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)std::max(cols, 1), (cuuint64_t)std::max<long long>(rows, 1)};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_KH, (cuuint32_t)TN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp16) failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

// rows x 16 fp16 matrix (32-byte rows): box = 128 rows x 16 halves, 32-byte swizzle (the augmented-K blocks).
static int make_tmap_aug(CUtensorMap* m, const void* base, long long rows) {
    PFN_encodeTiled enc;
    if (int rc = get_encode_fn(&enc)) return rc;
    LIRA_REQUIRE(((uintptr_t)base & 31) == 0, "tensor map: augmented block must be 32-byte aligned");
    cuuint64_t dims[2] = {16, (cuuint64_t)std::max<long long>(rows, 1)};
    cuuint64_t strides[1] = {32};
    cuuint32_t box[2] = {16, (cuuint32_t)TN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (augmented block) failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

// rows x cols byte matrix with row stride ld (bytes, % 16 == 0); box = 128 rows x 128 bytes, 128-byte swizzle (u8 scan).
static int make_tmap_u8(CUtensorMap* m, const void* base, long long rows, int cols, long long ld) {
    PFN_encodeTiled enc;
    if (int rc = get_encode_fn(&enc)) return rc;
    LIRA_REQUIRE(((uintptr_t)base & 15) == 0 && (ld % 16) == 0, "tensor map: base must be 16-byte aligned, ld % 16 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)std::max(cols, 1), (cuuint64_t)std::max<long long>(rows, 1)};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {(cuuint32_t)U8_KB, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (u8) failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

// rows x 64 byte matrix (the norm digits of the byte scan); box = 128 rows x 64 bytes, 64-byte swizzle.
static int make_tmap_u8_aug(CUtensorMap* m, const void* base, long long rows, CUtensorMapDataType dt) {
    PFN_encodeTiled enc;
    if (int rc = get_encode_fn(&enc)) return rc;
    LIRA_REQUIRE(((uintptr_t)base & 63) == 0, "tensor map: digit block must be 64-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)U8_AUG, (cuuint64_t)std::max<long long>(rows, 1)};
    cuuint64_t strides[1] = {(cuuint64_t)U8_AUG};
    cuuint32_t box[2] = {(cuuint32_t)U8_AUG, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (norm digits) failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// grow-only device buffers
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) LIRA_CUDA_OK(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        LIRA_CUDA_OK(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return (T*)p; }
};

struct Workspace {
    DevBuf q, sel, nsel, cmp, list_count, cursor, group_offsets, probe_offsets, group_queries, probe_slot, items,
        n_items, part_key, D, I, scores, probe_ids, nprobe, top1, thr, gq, cand_key, cand_count, qnorm, flags, redo, trace,
        seed_items, seed_out;
    void release() {
        for (DevBuf* b : {&q, &sel, &nsel, &cmp, &list_count, &cursor, &group_offsets, &probe_offsets, &group_queries,
                          &probe_slot, &items, &n_items, &part_key, &D, &I, &scores, &probe_ids, &nprobe, &top1, &thr, &gq,
                          &cand_key, &cand_count, &qnorm, &flags, &redo, &trace, &seed_items, &seed_out})
            b->release();
    }
};

}  // namespace lira

using namespace lira;

// a batch that was enqueued without waiting for its status words (lira_probe_search_enqueue_dev / _submit): everything
// lira_index_finish needs to check it and, in the rare cases the optimistic run is void, to answer it again
struct PendingBatch {
    cudaEvent_t ev = nullptr;      // recorded behind the copy of the status words (and, for host batches, of the results)
    int* h_flags = nullptr;        // pinned [4]: not exact in fp16 / queries left for the exact path / probe set truncated / --
    lira_model* m = nullptr;
    const float* d_q = nullptr;
    long long ldq = 0, Q = 0;
    int mode = 0;
    double value = 0;
    int k = 0, dedup = 1;
    float* d_D = nullptr;
    long long* d_I = nullptr;
    int* d_nprobe = nullptr;
    long long* d_cmp = nullptr;
    cudaStream_t st = nullptr;
    int cap = 0;                   // partitions-per-query cap the batch ran with
    bool enqueued = false;         // false: the batch did not qualify for the fused flow and nothing was launched
    int slot = -1;                 // host-buffer batches (lira_probe_search_submit): staging slot
};

struct lira_index {
    int device = 0, B = 0, d = 0, ds = 0, metric = 0;
    long long E = 0;
    float* vecs = nullptr;
    int* ids = nullptr;
    bool owns = true;
    long long* d_offsets = nullptr;
    int* d_list_order = nullptr;
    std::vector<long long> h_offsets;
    CUtensorMap tmap;
    cudaStream_t stream = nullptr;
    Workspace ws, ws_seed;
    int* h_flags = nullptr;      // pinned: the four status words of a tensor-core batch land here (no pageable staging on the way back)
    struct HostSlot {                            // lira_probe_search_submit / _wait: one batch of host queries in flight
        DevBuf q, D, I, nprobe, cmp;
        void* pin_in = nullptr;                  // pinned copy of the queries when the caller's array is pageable
        size_t pin_in_cap = 0;
        void* pin_out = nullptr;                 // pinned landing area of the results
        size_t pin_out_cap = 0;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
        cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};   // timing: upload begin / end, compute begin / end (see stage_pinned)
        bool timed_direct = false;               // this batch's direct upload from the caller's pinned array is being timed
        PendingBatch pb;
        bool busy = false, sync_done = false;
        const float* user_q = nullptr;
        size_t oI = 0, oC = 0, oN = 0;
    } slots[4];
    cudaStream_t st_in = nullptr, st_out = nullptr;   // copy streams of the submit / wait pipeline
    // A caller's PINNED query array: uploaded from directly (-> 0), or staged through this handle's own pinned buffer like a
    // pageable one (-> 1)? Pages pinned by another allocator can upload at half the rate of a cudaHostAlloc buffer on some
    // hosts; then the upload, not the kernels, sets the pipeline's period and the 0.3 ms host memcpy is the cheaper price.
    // Decided from the first timed direct uploads: staged iff the upload takes more than 0.7 of the time of the batch's kernels.
    int stage_pinned = -1;       // -1 undecided
    int direct_samples = 0;
    float direct_h2d_ms = 0.f, direct_compute_ms = 0.f;
    std::deque<PendingBatch> pending;            // enqueued, not yet checked (oldest first)
    std::vector<int*> flag_pool;                 // free pinned status slots
    std::vector<cudaEvent_t> event_pool;         // free events (timing disabled)
    void* h_stage = nullptr;     // pinned host staging for results (D2H into pageable user buffers is staged by the driver otherwise,
    size_t h_stage_cap = 0;      //   synchronously and in small pieces)
    DevBuf stats;                // {E_p, pairs} of the last timed scan (copied out of ws.n_items before it is reused)
    float* vnorm = nullptr;      // |v|^2 per list entry
    __half* vaug = nullptr;      // [E, 16] fp16 augmented-K block of every entry: (hi, lo, 0...) with |v|^2 = 2048 hi + lo (tensor-core path)
    __half* vecs16 = nullptr;    // [E, d16] fp16 shadow copy of vecs (exact when tc_ok): the tensor-core scan streams this
    int d16 = 0;                 // round_up(d, 8)
    CUtensorMap tmap16;
    __half* aaug = nullptr;      // [128, 16] constant fp16 augmented-K block of the query side: (-2048, -1, 0...)
    CUtensorMap tmap_vaug, tmap_aaug;
    int nprobe_cap = 64;         // threshold selection without a host round trip keeps at most this many lists per query (adaptive)
    bool tc_force_sync = false;  // next tensor-core attempt uses the exact pair count (after a truncated / inexact optimistic run)
    bool tc_ok = false;          // the tensor-core scan can serve this index (tc_mode != 0)
    int tc_mode = 0;             // 1: exact (small integers, bit-identical results); 2: approximate filter + exact re-rank (real-valued data)
    // byte-valued data (tc_mode 1 with every value in [0, 255], d <= 256): the integer tensor-core scan (u8_scan_kernels.cuh)
    // streams a one-byte-per-component shadow copy; the fp16 shadow copy is then built only when a batch needs it
    bool u8_ok = false;
    bool has16 = false;          // vecs16 / vaug exist
    bool has8 = false;           // vecs8 / nv_i exist
    bool prefer_u8 = false;      // exact-kNN handles: the byte copy is the one built at create time
    uint8_t* vaug8 = nullptr;    // [E + 256, 64] norm digits of every entry (the B side of the augmented MMAs)
    uint8_t* aaug8 = nullptr;    // [128, 64] their constant A side: (-128 x 63, -1) as signed bytes
    CUtensorMap tmap_vaug8, tmap_aaug8;
    uint8_t* vecs8 = nullptr;    // [E + 256, d8]
    int* nv_i = nullptr;         // [E + 256] -|v|^2 (L2) or 0 (IP)
    int d8 = 0;                  // round_up(d, 16)
    int u8_max_nseg = 1;         // most row segments (of U8_SEG_ROWS rows) any list is cut into
    long long u8_nseg_total = 0; // segments of all lists
    int u8_item_q() const { return U8_ITEM_Q; }   // queries per work item
    CUtensorMap tmap8;
    float tc_sigma = 1.0f;       // power-of-two scale of the fp16 shadow copy
    float tc_vmax = 0.0f;        // largest |v| of the index (mode 2: error margin)
    bool use_tc = true;
    bool last_u8 = false;        // the last tensor-core batch ran the byte scan
    int last_parts = 4;          // candidate regions per pair of the last tensor-core batch
    int last_path = 0;           // 0 = CUDA-core scan, 1 = tensor-core scan
    int last_redo = 0;           // queries of the last tensor-core batch redone on the CUDA cores
    bool timing = false, timing_pending = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [0,1] filter kernel, [2,3] whole search, [4,5] seed + filter + refine
    float last_scan_ms = 0.f, last_total_ms = 0.f, last_scan_total_ms = 0.f;
    bool scan_total_valid = false;
    long long last_scan_bytes = 0, last_scan_pairs = 0, last_Q = 0;
    int last_k = 0;
    int num_sms = 148;
};

struct lira_model {
    int device = 0, B = 0, Bp = 0, d = 0, ds = 0;
    float* centroids = nullptr;  // [B, ds]
    float *mean = nullptr, *scale = nullptr;
    float* W[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* bias[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int out_dim[6], in_dim[6], in_ld[6];
    CUtensorMap tm_cent, tm_w[6];
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;                 // vector_net runs here, concurrently with the centroid features and distance_net
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int num_sms = 148;
    DevBuf feats, h1, cat, h2, h5, scores, q;
    // tensor-core front end (tc_dense_kernels.cuh): error-free hi / lo splits of the static operands ...
    bool use_tc = true;
    float *mu = nullptr, *cent_h = nullptr, *cent_l = nullptr, *cn = nullptr;   // centred centroids, |c'|^2
    float* W_h[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* W_l[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    CUtensorMap tm_ch, tm_cl, tm_wh[6], tm_wl[6];
    // ... and of the activations of the current batch
    DevBuf qch, qcl, qn, qrh, qrl, fh, fl, h1h, h1l, cath, catl, h2h, h2l, h5h, h5l;
};

namespace lira {

// ---------------------------------------------------------------------------------------------
// kernel attribute setup (once per process and device)
// ---------------------------------------------------------------------------------------------
template <class K>
static int set_smem(K kernel, size_t bytes) {
    LIRA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

static int init_kernels(int device) {
    static bool done_dev[64] = {false};
    if (device < 64 && done_dev[device]) return 0;
    int rc = 0;
    rc |= set_smem(scan_lists_kernel<OP_L2, 1>, scan_smem_bytes<1>());
    rc |= set_smem(scan_lists_kernel<OP_DOT, 1>, scan_smem_bytes<1>());
    rc |= set_smem(scan_lists_kernel<OP_L2, 4>, scan_smem_bytes<4>());
    rc |= set_smem(scan_lists_kernel<OP_DOT, 4>, scan_smem_bytes<4>());
    rc |= set_smem(tc_scan_kernel<false, false>, TC_SMEM_BYTES);
    rc |= set_smem(tc_scan_kernel<false, true>, TC_SMEM_BYTES);
    rc |= set_smem(tc_scan_kernel<true, false>, TC_SMEM_BYTES);
    rc |= set_smem(u8_scan_kernel<true, false>, u8_smem_bytes(true));
    rc |= set_smem(u8_scan_kernel<false, false>, u8_smem_bytes(false));
    rc |= set_smem(u8_scan_kernel<true, true>, u8_smem_bytes(true));
    rc |= set_smem(u8_scan_kernel<false, true>, u8_smem_bytes(false));
    rc |= set_smem(dense_tile_kernel<64, OP_L2, EPI_FEATURE>, DENSE_SMEM_BYTES);
    rc |= set_smem(dense_tile_kernel<32, OP_DOT, EPI_BIAS_RELU>, DENSE_SMEM_BYTES);
    rc |= set_smem(dense_tile_kernel<64, OP_DOT, EPI_BIAS_SIGMOID>, DENSE_SMEM_BYTES);
    rc |= set_smem(tc_dense_kernel<TD_EPI_FEATURE>, TD_SMEM_BYTES);
    rc |= set_smem(tc_dense_kernel<TD_EPI_BIAS_RELU>, TD_SMEM_BYTES);
    rc |= set_smem(tc_dense_kernel<TD_EPI_BIAS_SIGMOID>, TD_SMEM_BYTES);
    rc |= set_smem(tc_dense_kernel<TD_EPI_SELECT>, TD_SMEM_BYTES);
    if (device < 64) done_dev[device] = (rc == 0);
    return rc;
}

static int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: liblira_b200 has no CPU fallback");
        return 3;
    }
    LIRA_REQUIRE(device >= 0 && device < n, "device ordinal out of range");
    LIRA_CUDA_OK(cudaSetDevice(device));
    return init_kernels(device);
}

// ---------------------------------------------------------------------------------------------
// dense tile GEMM launcher
// ---------------------------------------------------------------------------------------------
template <int TM, int OP, int EPI>
static int launch_dense(const CUtensorMap& tm, const DenseParams& p, cudaStream_t st) {
    dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM);
    if (p.M <= 0 || p.N <= 0) return 0;
    dense_tile_kernel<TM, OP, EPI><<<grid, N_THREADS, DENSE_SMEM_BYTES, st>>>(tm, p);
    LIRA_LAUNCH_CHECK();
    return 0;
}

static int model_forward_simt(lira_model* m, const float* d_q, long long ldq, long long Q, float* d_scores, long long lds,
                              float* d_feats_out, long long ldf_out, cudaStream_t st) {
    LIRA_REQUIRE((ldq % 4) == 0 && ((uintptr_t)d_q & 15) == 0, "queries must be 16-byte aligned with ld % 4 == 0");
    const size_t Qs = (size_t)Q;
    if (int rc = m->feats.ensure(Qs * m->Bp * 4)) return rc;
    if (int rc = m->h1.ensure(Qs * 128 * 4)) return rc;
    if (int rc = m->cat.ensure(Qs * 128 * 4)) return rc;
    if (int rc = m->h2.ensure(Qs * 128 * 4)) return rc;
    if (int rc = m->h5.ensure(Qs * 128 * 4)) return rc;
    float* feats = d_feats_out ? d_feats_out : m->feats.as<float>();
    const long long ldf = d_feats_out ? ldf_out : m->Bp;
    LIRA_REQUIRE((ldf % 4) == 0, "feats output needs a row stride that is a multiple of 4");
    DenseParams p;
    // K0: ||q - c_b||_2, standardised (utils.py:98-118,142-167; search.cpp:220-250)
    p = DenseParams{d_q, ldq, (int)Q, m->ds, m->B, feats, ldf, 0, m->mean, m->scale};
    if (int rc = launch_dense<64, OP_L2, EPI_FEATURE>(m->tm_cent, p, st)) return rc;
    // distance_net (model_probing.py:12-17)
    p = DenseParams{feats, ldf, (int)Q, m->in_ld[0], 128, m->h1.as<float>(), 128, 0, m->bias[0], nullptr};
    if (int rc = launch_dense<32, OP_DOT, EPI_BIAS_RELU>(m->tm_w[0], p, st)) return rc;
    p = DenseParams{m->h1.as<float>(), 128, (int)Q, 128, 64, m->cat.as<float>(), 128, 0, m->bias[1], nullptr};
    if (int rc = launch_dense<32, OP_DOT, EPI_BIAS_RELU>(m->tm_w[1], p, st)) return rc;
    // vector_net (model_probing.py:19-24)
    p = DenseParams{d_q, ldq, (int)Q, m->ds, 128, m->h2.as<float>(), 128, 0, m->bias[2], nullptr};
    if (int rc = launch_dense<32, OP_DOT, EPI_BIAS_RELU>(m->tm_w[2], p, st)) return rc;
    p = DenseParams{m->h2.as<float>(), 128, (int)Q, 128, 64, m->cat.as<float>(), 128, 64, m->bias[3], nullptr};
    if (int rc = launch_dense<32, OP_DOT, EPI_BIAS_RELU>(m->tm_w[3], p, st)) return rc;
    // fc (model_probing.py:26-31): cat -> 128 -> B, sigmoid
    p = DenseParams{m->cat.as<float>(), 128, (int)Q, 128, 128, m->h5.as<float>(), 128, 0, m->bias[4], nullptr};
    if (int rc = launch_dense<32, OP_DOT, EPI_BIAS_RELU>(m->tm_w[4], p, st)) return rc;
    p = DenseParams{m->h5.as<float>(), 128, (int)Q, 128, m->B, d_scores, lds, 0, m->bias[5], nullptr};
    if (int rc = launch_dense<64, OP_DOT, EPI_BIAS_SIGMOID>(m->tm_w[5], p, st)) return rc;
    return 0;
}

// One layer on the tensor cores: out = epi(A . B^T), A and B given as hi / lo pairs (tc_dense_kernels.cuh).
template <int EPI>
static int launch_tc_dense(const float* a_h, const float* a_l, long long M, int K, long long lda, const CUtensorMap& tm_bh,
                           const CUtensorMap& tm_bl, int N, const TdParams& tp, int num_sms, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    CUtensorMap tm_ah, tm_al;
    if (int rc = make_tmap(&tm_ah, a_h, M, K, lda)) return rc;
    if (int rc = make_tmap(&tm_al, a_l, M, K, lda)) return rc;
    const long long tiles = ((M + TC_M - 1) / TC_M) * ((N + TC_N - 1) / TC_N);
    const int grid = (int)std::min<long long>(tiles, num_sms);
    LIRA_CUDA_OK(launch_pdl(tc_dense_kernel<EPI>, dim3(grid), dim3(TD_THREADS), TD_SMEM_BYTES, st, tm_ah, tm_al, tm_bh, tm_bl, tp));
    LIRA_LAUNCH_CHECK();
    return 0;
}

// selection fused into the last layer (tc_dense_kernel<TD_EPI_SELECT>): where the probe lists go
struct FusedSelect {
    int* sel;
    int* nsel;
    int* list_count;
    unsigned long long* rowbest;
    int cap, mode;
    float thr;
};

static int model_ensure_tc(lira_model* m, long long Q) {
    const size_t Qs = (size_t)std::max<long long>(Q, 1);
    for (DevBuf* b : {&m->qch, &m->qcl, &m->qrh, &m->qrl})
        if (int rc = b->ensure(Qs * m->ds * 4)) return rc;
    if (int rc = m->qn.ensure(Qs * 4)) return rc;
    for (DevBuf* b : {&m->fh, &m->fl})
        if (int rc = b->ensure(Qs * m->Bp * 4)) return rc;
    for (DevBuf* b : {&m->h1h, &m->h1l, &m->cath, &m->catl, &m->h2h, &m->h2l, &m->h5h, &m->h5l})
        if (int rc = b->ensure(Qs * 128 * 4)) return rc;
    return 0;
}

// prepped: the hi / lo splits of the queries and |q'|^2 are already in m->qch .. m->qn (prep_queries_kernel);
// fs != null: the last layer selects instead of writing scores (d_scores unused)
static int model_forward_tc(lira_model* m, const float* d_q, long long ldq, long long Q, float* d_scores, long long lds,
                            float* d_feats_out, long long ldf_out, cudaStream_t st, bool prepped = false, const FusedSelect* fs = nullptr) {
    LIRA_REQUIRE((ldq % 4) == 0 && ((uintptr_t)d_q & 15) == 0 && (lds % 4) == 0, "queries / scores must be 16-byte aligned with ld % 4 == 0");
    LIRA_REQUIRE(!d_feats_out || (ldf_out % 4) == 0, "feats output needs a row stride that is a multiple of 4");
    const int ds = m->ds, B = m->B, Bp = m->Bp;
    if (int rc = model_ensure_tc(m, Q)) return rc;
    const int num_sms = m->num_sms;
    const int warps = 8;
    const int sgrid = (int)((Q + warps - 1) / warps);
    if (!prepped) {
        // queries: centred split (+ |q'|^2) for the distance features, raw split for vector_net
        split_rows_kernel<<<sgrid, warps * 32, 0, st>>>(d_q, ldq, m->d, Q, m->mu, m->qch.as<float>(), m->qcl.as<float>(), ds, m->qn.as<float>());
        LIRA_LAUNCH_CHECK();
        split_rows_kernel<<<sgrid, warps * 32, 0, st>>>(d_q, ldq, m->d, Q, nullptr, m->qrh.as<float>(), m->qrl.as<float>(), ds, nullptr);
        LIRA_LAUNCH_CHECK();
    }
    TdParams tp;
    // vector_net (model_probing.py:19-24) depends on the raw queries only: it runs on a second stream next to the centroid
    // features and distance_net, and joins before fc
    cudaStream_t sv = (m->side && !getenv("LIRA_NO_SIDE_STREAM")) ? m->side : st;
    if (sv != st) {
        LIRA_CUDA_OK(cudaEventRecord(m->ev_fork, st));
        LIRA_CUDA_OK(cudaStreamWaitEvent(sv, m->ev_fork, 0));
    }
    tp = TdParams{(int)Q, 128, m->d, m->h2h.as<float>(), m->h2l.as<float>(), nullptr, 128, 0, 0, m->bias[2], nullptr, nullptr, nullptr};
    if (int rc = launch_tc_dense<TD_EPI_BIAS_RELU>(m->qrh.as<float>(), m->qrl.as<float>(), Q, m->d, ds, m->tm_wh[2], m->tm_wl[2], 128, tp, num_sms, sv)) return rc;
    tp = TdParams{(int)Q, 64, 128, m->cath.as<float>(), m->catl.as<float>(), nullptr, 128, 0, 64, m->bias[3], nullptr, nullptr, nullptr};
    if (int rc = launch_tc_dense<TD_EPI_BIAS_RELU>(m->h2h.as<float>(), m->h2l.as<float>(), Q, 128, 128, m->tm_wh[3], m->tm_wl[3], 64, tp, num_sms, sv)) return rc;
    if (sv != st) LIRA_CUDA_OK(cudaEventRecord(m->ev_join, sv));
    // K0: ||q - c_b||_2, standardised (utils.py:98-118,142-167; search.cpp:220-250)
    tp = TdParams{(int)Q, B, ds, m->fh.as<float>(), m->fl.as<float>(), d_feats_out, Bp, (long)ldf_out, 0,
                  m->cn, m->mean, m->scale, m->qn.as<float>()};
    if (int rc = launch_tc_dense<TD_EPI_FEATURE>(m->qch.as<float>(), m->qcl.as<float>(), Q, ds, ds, m->tm_ch, m->tm_cl, B, tp, num_sms, st)) return rc;
    // distance_net (model_probing.py:12-17)
    tp = TdParams{(int)Q, 128, B, m->h1h.as<float>(), m->h1l.as<float>(), nullptr, 128, 0, 0, m->bias[0], nullptr, nullptr, nullptr};
    if (int rc = launch_tc_dense<TD_EPI_BIAS_RELU>(m->fh.as<float>(), m->fl.as<float>(), Q, B, Bp, m->tm_wh[0], m->tm_wl[0], 128, tp, num_sms, st)) return rc;
    tp = TdParams{(int)Q, 64, 128, m->cath.as<float>(), m->catl.as<float>(), nullptr, 128, 0, 0, m->bias[1], nullptr, nullptr, nullptr};
    if (int rc = launch_tc_dense<TD_EPI_BIAS_RELU>(m->h1h.as<float>(), m->h1l.as<float>(), Q, 128, 128, m->tm_wh[1], m->tm_wl[1], 64, tp, num_sms, st)) return rc;
    if (sv != st) LIRA_CUDA_OK(cudaStreamWaitEvent(st, m->ev_join, 0));
    // fc (model_probing.py:26-31): cat -> 128 -> B, sigmoid
    tp = TdParams{(int)Q, 128, 128, m->h5h.as<float>(), m->h5l.as<float>(), nullptr, 128, 0, 0, m->bias[4], nullptr, nullptr, nullptr};
    if (int rc = launch_tc_dense<TD_EPI_BIAS_RELU>(m->cath.as<float>(), m->catl.as<float>(), Q, 128, 128, m->tm_wh[4], m->tm_wl[4], 128, tp, num_sms, st)) return rc;
    if (fs) {
        tp = TdParams{(int)Q, B, 128, nullptr, nullptr, nullptr, 0, 0, 0, m->bias[5], nullptr, nullptr, nullptr};
        tp.sel = fs->sel; tp.nsel = fs->nsel; tp.list_count = fs->list_count; tp.rowbest = fs->rowbest;
        tp.sel_cap = fs->cap; tp.sel_mode = fs->mode; tp.sel_thr = fs->thr;
        if (int rc = launch_tc_dense<TD_EPI_SELECT>(m->h5h.as<float>(), m->h5l.as<float>(), Q, 128, 128, m->tm_wh[5], m->tm_wl[5], B, tp, num_sms, st)) return rc;
        return 0;
    }
    tp = TdParams{(int)Q, B, 128, nullptr, nullptr, d_scores, 0, (long)lds, 0, m->bias[5], nullptr, nullptr, nullptr};
    if (int rc = launch_tc_dense<TD_EPI_BIAS_SIGMOID>(m->h5h.as<float>(), m->h5l.as<float>(), Q, 128, 128, m->tm_wh[5], m->tm_wl[5], B, tp, num_sms, st)) return rc;
    return 0;
}

static int model_forward(lira_model* m, const float* d_q, long long ldq, long long Q, float* d_scores, long long lds,
                         float* d_feats_out, long long ldf_out, cudaStream_t st) {
    if (Q <= 0) return 0;
    if (m->use_tc) return model_forward_tc(m, d_q, ldq, Q, d_scores, lds, d_feats_out, ldf_out, st);
    return model_forward_simt(m, d_q, ldq, Q, d_scores, lds, d_feats_out, ldf_out, st);
}

// ---------------------------------------------------------------------------------------------
// the search core: probe sets (from scores, explicit CSR, or all pairs) -> groups -> scan -> merge
// ---------------------------------------------------------------------------------------------
struct ProbeSpec {
    int kind;  // 0 = select from scores, 1 = explicit CSR (device), 2 = all pairs
    const float* d_scores = nullptr;
    long long lds = 0;
    int mode = 0;
    double value = 0;
    const long long* d_probe_offsets = nullptr;
    const int* d_probe_ids = nullptr;
    long long P = 0;
    int* d_bad_flag = nullptr;   // kind 1: set to 1 when a probed list id is outside [0, B) (default: a scratch word nobody reads)
};

__global__ void csr_hist_kernel(const long long* probe_offsets, const int* probe_ids, const long long* list_offsets,
                                int Q, int B, int* list_count, long long* cmp, int* nsel, int* bad, const int* mask) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    if (mask && !mask[q]) return;
    long long c = 0;
    int nv = 0;
    const long long lo = probe_offsets[q], hi = probe_offsets[q + 1];
    for (long long j = lo + lane; j < hi; j += 32) {
        const int b = probe_ids[j];
        if (b < 0 || b >= B) { *bad = 1; continue; }   // ignored (and reported where the caller synchronises)
        atomicAdd(list_count + b, 1);
        c += list_offsets[b + 1] - list_offsets[b];
        ++nv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
    }
    if (lane == 0) {
        if (cmp) cmp[q] = c;
        nsel[q] = nv;
    }
}

__global__ void scatter_csr_kernel(const long long* probe_offsets, const int* probe_ids, const long long* group_offsets,
                                   int* cursor, int* group_queries, int* probe_slot, int Q, int B, const int* mask) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    if (mask && !mask[q]) return;
    for (long long j = probe_offsets[q] + lane; j < probe_offsets[q + 1]; j += 32) {
        const int b = probe_ids[j];
        if (b < 0 || b >= B) { probe_slot[j] = -1; continue; }
        const int pos = (int)group_offsets[b] + atomicAdd(cursor + b, 1);
        group_queries[pos] = q;
        probe_slot[j] = pos;
    }
}

__global__ void uniform_groups_kernel(long long* group_offsets, int B, long long Q) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= B; i += gridDim.x * blockDim.x) group_offsets[i] = i * Q;
}

__global__ void iota_i32_kernel(int* out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (int)i;
}

__global__ void copy_nprobe_kernel(const int* nsel, int* out, int Q) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += gridDim.x * blockDim.x) out[i] = nsel[i];
}

static int launch_scan(lira_index* h, const ScanParams& sp, int k, cudaStream_t st) {
    const int grid = h->num_sms;
    const int S = k <= 32 ? 1 : 4;
    if (h->metric == LIRA_METRIC_L2) {
        if (S == 1) scan_lists_kernel<OP_L2, 1><<<grid, N_THREADS, scan_smem_bytes<1>(), st>>>(h->tmap, sp);
        else scan_lists_kernel<OP_L2, 4><<<grid, N_THREADS, scan_smem_bytes<4>(), st>>>(h->tmap, sp);
    } else {
        if (S == 1) scan_lists_kernel<OP_DOT, 1><<<grid, N_THREADS, scan_smem_bytes<1>(), st>>>(h->tmap, sp);
        else scan_lists_kernel<OP_DOT, 4><<<grid, N_THREADS, scan_smem_bytes<4>(), st>>>(h->tmap, sp);
    }
    LIRA_LAUNCH_CHECK();
    return 0;
}

static int launch_merge(const MergeParams& mp, cudaStream_t st) {
    const int warps = 8;
    const int grid = (mp.Q + warps - 1) / warps;
    if (mp.Q <= 0) return 0;
    if (mp.k <= 32) merge_topk_kernel<1><<<grid, warps * 32, 0, st>>>(mp);
    else merge_topk_kernel<4><<<grid, warps * 32, 0, st>>>(mp);
    LIRA_LAUNCH_CHECK();
    return 0;
}

// Selection (or explicit / exhaustive probe sets) -> per-list query groups -> work items of `tile` query
// rows. On return ws.group_queries / ws.probe_slot / ws.items / ws.n_items are set and *P_out is the number
// of (query, list) pairs. `probe_offsets_out` is the CSR the merge must use.
static int prepare_groups(lira_index* h, Workspace& ws, long long Q, const ProbeSpec& ps, int tile, long long* d_cmp,
                          long long* P_out, const long long** probe_offsets_out, int* extra_flag_host,
                          const int* d_extra_flag, const int* d_mask, cudaStream_t st, int* d_trunc_flag = nullptr, int seg_rows = 0) {
    LIRA_REQUIRE(Q >= 0 && Q < (1ll << 31), "Q out of range");
    const int B = h->B;
    const int warps = 8;
    const int qgrid = (int)((Q + warps - 1) / warps);
    if (int rc = ws.nsel.ensure((size_t)(Q + 1) * 4)) return rc;
    if (int rc = ws.top1.ensure((size_t)(Q + 1) * 8)) return rc;  // seed ids: two per query
    if (int rc = ws.probe_offsets.ensure((size_t)(Q + 1) * 8)) return rc;
    if (int rc = ws.group_offsets.ensure((size_t)(B + 1) * 8)) return rc;
    if (int rc = ws.list_count.ensure((size_t)(B + 1) * 4)) return rc;
    if (int rc = ws.cursor.ensure((size_t)(B + 1) * 4)) return rc;
    if (int rc = ws.n_items.ensure(128)) return rc;
    LIRA_CUDA_OK(cudaMemsetAsync(ws.list_count.p, 0, (size_t)(B + 1) * 4, st));
    LIRA_CUDA_OK(cudaMemsetAsync(ws.cursor.p, 0, (size_t)(B + 1) * 4, st));
    LIRA_CUDA_OK(cudaMemsetAsync(ws.n_items.p, 0, 128, st));
    long long P = 0;
    *P_out = 0;
    if (Q == 0) return 0;

    if (ps.kind == 0) {
        LIRA_REQUIRE(ps.mode >= 0 && ps.mode <= 2, "unknown selection mode");
        if (ps.mode == LIRA_SELECT_TOPN) LIRA_REQUIRE(ps.value >= 1 && ps.value <= 128, "top-nprobe must be in [1, 128]");
        if (int rc = ws.sel.ensure((size_t)Q * B * 4)) return rc;
        // no host round trip when the caller asked for it (d_trunc_flag) and the bound Q * cap is affordable
        int per_q_cap = 0;
        if (d_trunc_flag) {
            per_q_cap = ps.mode == LIRA_SELECT_TOPN ? std::max(1, std::min({(int)ps.value, B, 128})) : std::min(B, h->nprobe_cap);
            if (Q * (long long)per_q_cap > (4ll << 20)) per_q_cap = 0;
        }
        const bool nosync = per_q_cap > 0;
        SelectParams sp{ps.d_scores, ps.lds, (int)Q, B, ps.mode, ps.value, h->d_offsets, ws.sel.as<int>(),
                        ws.nsel.as<int>(), d_cmp, ws.list_count.as<int>(), ws.top1.as<int>(), d_mask,
                        (nosync && ps.mode != LIRA_SELECT_TOPN) ? per_q_cap : 0, d_trunc_flag};
        select_kernel<4><<<qgrid, warps * 32, 0, st>>>(sp);
        LIRA_LAUNCH_CHECK();
        // both prefix sums in one launch: block 0 the probe offsets over the queries, block 1 the group offsets over the lists
        exclusive_scan_kernel<<<2, 1024, 0, st>>>(ws.nsel.as<int>(), ws.probe_offsets.as<long long>(), (int)Q,
                                                  ws.list_count.as<int>(), ws.group_offsets.as<long long>(), B);
        LIRA_LAUNCH_CHECK();
        // the one host round trip of the query phase: P sizes the partial-result buffers
        if (nosync) {
            // upper bound instead of the round trip: top-nprobe selects exactly `want` per query, the threshold modes keep at
            // most h->nprobe_cap per query (select_kernel raises flags[2] when a query had more; the caller checks it --
            // and flags[0], the exactness of the query batch -- at its own final synchronisation and reruns if needed)
            P = Q * (long long)per_q_cap;
            if (extra_flag_host) *extra_flag_host = 1;
        } else {
            LIRA_CUDA_OK(cudaMemcpyAsync(&P, ws.probe_offsets.as<long long>() + Q, 8, cudaMemcpyDeviceToHost, st));
            if (extra_flag_host && d_extra_flag)
                LIRA_CUDA_OK(cudaMemcpyAsync(extra_flag_host, d_extra_flag, 4, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
        }
    } else if (ps.kind == 1) {
        P = ps.P;
        int* bad = ps.d_bad_flag ? ps.d_bad_flag : ws.n_items.as<int>() + 8;
        csr_hist_kernel<<<qgrid, warps * 32, 0, st>>>(ps.d_probe_offsets, ps.d_probe_ids, h->d_offsets, (int)Q, B,
                                                      ws.list_count.as<int>(), d_cmp, ws.nsel.as<int>(), bad, d_mask);
        LIRA_LAUNCH_CHECK();
        exclusive_scan_kernel<<<1, 1024, 0, st>>>(ws.list_count.as<int>(), ws.group_offsets.as<long long>(), B);
        LIRA_LAUNCH_CHECK();
        if (extra_flag_host && d_extra_flag) {
            LIRA_CUDA_OK(cudaMemcpyAsync(extra_flag_host, d_extra_flag, 4, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
        }
    } else {
        P = Q * (long long)B;
    }
    LIRA_REQUIRE(P < (1ll << 31), "too many (query, list) pairs in one call; split the query batch");
    *P_out = P;
    const size_t Ps = (size_t)std::max<long long>(P, 1);
    if (int rc = ws.group_queries.ensure(Ps * 4)) return rc;
    if (int rc = ws.probe_slot.ensure(Ps * 4)) return rc;
    size_t max_items = Ps / 8 + B + 1;  // generous: also covers a later re-tiling with a smaller tile
    if (seg_rows > 0) max_items = std::max(max_items, (Ps / tile + 1) * (size_t)h->u8_max_nseg + (size_t)h->u8_nseg_total + B + 1);
    if (int rc = ws.items.ensure(max_items * sizeof(ScanItem))) return rc;

    if (ps.kind == 0) {
        ScatterParams sc{ws.sel.as<int>(), ws.nsel.as<int>(), ws.probe_offsets.as<long long>(),
                         ws.group_offsets.as<long long>(), ws.cursor.as<int>(), ws.group_queries.as<int>(),
                         ws.probe_slot.as<int>(), (int)Q, B};
        scatter_groups_kernel<<<qgrid, warps * 32, 0, st>>>(sc);
        LIRA_LAUNCH_CHECK();
        *probe_offsets_out = ws.probe_offsets.as<long long>();
    } else if (ps.kind == 1) {
        scatter_csr_kernel<<<qgrid, warps * 32, 0, st>>>(ps.d_probe_offsets, ps.d_probe_ids,
                                                         ws.group_offsets.as<long long>(), ws.cursor.as<int>(),
                                                         ws.group_queries.as<int>(), ws.probe_slot.as<int>(), (int)Q, B, d_mask);
        LIRA_LAUNCH_CHECK();
        *probe_offsets_out = ps.d_probe_offsets;
    } else {
        uniform_groups_kernel<<<grid_for(B + 1, 256), 256, 0, st>>>(ws.group_offsets.as<long long>(), B, Q);
        LIRA_LAUNCH_CHECK();
        fill_all_pairs_kernel<<<grid_for(P, 256), 256, 0, st>>>((int)Q, B, ws.group_queries.as<int>(),
                                                                ws.probe_slot.as<int>());
        LIRA_LAUNCH_CHECK();
        iota_offsets_kernel<<<grid_for(Q + 1, 256), 256, 0, st>>>(ws.probe_offsets.as<long long>(), Q, B);
        LIRA_LAUNCH_CHECK();
        *probe_offsets_out = ws.probe_offsets.as<long long>();
    }
    build_items_kernel<<<1, 1024, 0, st>>>(h->d_list_order, ws.group_offsets.as<long long>(), h->d_offsets, B, tile,
                                           ws.items.as<ScanItem>(), ws.n_items.as<int>(),
                                           (unsigned long long*)((char*)ws.n_items.p + 64), seg_rows);
    LIRA_LAUNCH_CHECK();
    return 0;
}

// exact CUDA-core scan of the prepared work items -> ws.part_key[P, k]
static int simt_scan(lira_index* h, Workspace& ws, const float* d_q, long long ldq, long long P, int k, int store_local,
                     int max_rows, bool timed, cudaStream_t st) {
    LIRA_REQUIRE((ldq % 4) == 0 && ((uintptr_t)d_q & 15) == 0, "queries must be 16-byte aligned with ld % 4 == 0");
    if (int rc = ws.part_key.ensure((size_t)std::max<long long>(P, 1) * k * 8)) return rc;
    ScanParams sp;
    sp.q = d_q;
    sp.ldq = ldq;
    sp.d = h->ds;
    sp.group_queries = ws.group_queries.as<int>();
    sp.list_offsets = h->d_offsets;
    sp.list_ids = h->ids;
    sp.items = ws.items.as<ScanItem>();
    sp.n_items = ws.n_items.as<int>();
    sp.part_key = ws.part_key.as<unsigned long long>();
    sp.k = k;
    sp.store_local = store_local;
    sp.max_rows = max_rows;
    if (timed && h->timing) { LIRA_CUDA_OK(cudaEventRecord(h->ev[0], st)); h->scan_total_valid = false; }
    if (int rc = launch_scan(h, sp, k, st)) return rc;
    if (timed && h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[1], st));
    return 0;
}

static int save_stats(lira_index* h, Workspace& ws, cudaStream_t st) {
    if (!h->timing) return 0;
    if (int rc = h->stats.ensure(16)) return rc;
    LIRA_CUDA_OK(cudaMemcpyAsync(h->stats.p, (char*)ws.n_items.p + 64, 16, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// grouping + exact scan (get_cmp_recall, list_search, kNN, and the online path on the CUDA cores)
static int run_grouped_scan(lira_index* h, const float* d_q, long long ldq, long long Q, const ProbeSpec& ps, int k,
                            int store_local, long long* d_cmp, long long* P_out, const long long** po_out,
                            cudaStream_t st, const int* d_mask = nullptr, bool timed = true) {
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    const long long* po = nullptr;
    if (int rc = prepare_groups(h, h->ws, Q, ps, SCAN_TM_MAX, d_cmp, P_out, &po, nullptr, nullptr, d_mask, st)) return rc;
    if (po_out) *po_out = po;
    if (Q == 0) return 0;
    if (timed) if (int rc = save_stats(h, h->ws, st)) return rc;
    return simt_scan(h, h->ws, d_q, ldq, *P_out, k, store_local, 0, timed, st);
}

static int ensure_fp16_shadow(lira_index* h);
static int ensure_u8_shadow(lira_index* h);
static constexpr int U8_CAPK = 64;            // byte scan: candidate slots per (query, list) pair, k <= 16
static constexpr int U8_CAPP = 128;           //            ... k > 16 (exhaustive probe sets over base segments)
static constexpr int U8_SEED_LISTS = 16;      // byte copy, exhaustive probe sets: lists (base segments) the seed pass scans
static constexpr int TC_SEED_ROWS = 384;      // CUDA-core seed (k > 16): rows of each of the two best probed lists
static constexpr int TC_SEED_ROWS_MAIN = 2048;   // seed pass on the filter's own work items: first rows of every probed list (0 = all; measured best)
static constexpr int TC_SEED_ROWS_TC = 0;     // tensor-core seed (k <= 16): rows of each of the two best probed lists (0 = all: the pass streams
                                              // (nearly) every list once anyway, and whole lists halve the survivors of the filter pass)

struct TcStage {   // one tensor-core batch after the grouping: what the seed / filter / refine launches need
    long long Q = 0, P = 0;          // P: (upper bound of the) number of (query, list) pairs -- sizes the candidate regions
    const long long* po = nullptr;   // probe offsets of the queries (refine)
    int k = 0, dedup = 1;
    const float* d_q = nullptr;
    long long ldq = 0;
    float* d_D = nullptr;
    long long* d_I = nullptr;
    int* redo_count = nullptr;       // device counter of queries left for the exact path
    const CUtensorMap* tmap_q = nullptr;
    float margin_c = 0.f, margin_abs = 0.f;
    int* seed_counter = nullptr;     // zeroed ticket counters of the two passes
    int* filter_counter = nullptr;
    bool u8 = false;                 // byte-valued index and batch: the integer tensor-core scan (u8_scan_kernels.cuh)
    bool counts_zeroed = false;      // byte scan: the pair counters were zeroed by the front end (fused flow)
    int q_mod = 0;                   // byte scan, exhaustive probe sets: the batch size (one copy of the query rows serves every list)
    bool private_regions = false;    // byte scan, exhaustive probe sets: one candidate region per (pair, column part), no atomics
    int seg_rows = U8_SEG_ROWS;      // byte scan: rows of a list per work item (exhaustive probe sets: whole lists)
};

static U8Params u8_params(const lira_index* h, Workspace& ws, const TcStage& sg) {
    U8Params up;
    up.group_queries = ws.group_queries.as<int>();
    up.list_offsets = h->d_offsets;
    up.items = ws.items.as<ScanItem>();
    up.n_items = ws.n_items.as<int>();
    up.work_counter = nullptr;
    up.d8 = h->d8;
    up.seg_rows = sg.seg_rows;
    up.q_mod = sg.q_mod;
    up.nv = h->nv_i;
    up.dbg = ws.n_items.as<int>() + 10;
    up.trace = nullptr;
    up.qnorm = ws.qnorm.as<float>();
    up.thr = ws.thr.as<uint32_t>();
    up.cand_key = nullptr;
    up.cand_count = nullptr;
    up.seed_out = nullptr;
    up.private_regions = sg.private_regions ? 1 : 0;
    up.cap = 0;
    up.k = sg.k;
    up.is_ip = h->metric == LIRA_METRIC_IP;
    up.exp = 0;
    return up;
}

// approximate mode: |accumulator - exact| <= M(q) = margin_c sqrt(sigma^2 |q|^2) + margin_abs (units of the scaled copy):
// operand rounding (2^-11 relative each, 5 % slack), fp32 accumulation inside the tensor core (2^-21 per term, generous),
// fp16 subnormal flushing of tiny components, and the (hi, lo) representation of sigma^2 |v|^2
static void tc_margins(const lira_index* h, float* margin_c, float* margin_abs) {
    *margin_c = 0.f;
    *margin_abs = 0.f;
    if (h->tc_mode != 2) return;
    const bool is_ip = h->metric == LIRA_METRIC_IP;
    const float W = h->tc_sigma * h->tc_vmax, sd = std::sqrt((float)h->d);
    const float per = 1.05f * 0.0009765625f + (float)h->d16 * 4.76837158e-7f;
    *margin_c = W * ((is_ip ? 1.0f : 2.0f) * per + sd * 1.1920929e-7f);
    *margin_abs = W * sd * 5.9604645e-8f + (is_ip ? 0.0f : W * W * 9.5367432e-7f);
}

// seed pass on the filter's own work items (ws.thr must hold +inf): see tc_search
static int tc_seed_main(lira_index* h, Workspace& ws, const TcStage& sg, cudaStream_t st) {
    TcParams sp;
    sp.group_queries = ws.group_queries.as<int>();
    sp.list_offsets = h->d_offsets;
    sp.items = ws.items.as<ScanItem>();
    sp.n_items = ws.n_items.as<int>();
    sp.work_counter = sg.seed_counter;
    sp.nk = (h->d16 + TC_KH - 1) / TC_KH;
    sp.max_rows = getenv("LIRA_TC_SEED_ROWS") ? atoi(getenv("LIRA_TC_SEED_ROWS")) : TC_SEED_ROWS_MAIN;
    sp.exp = 0;
    sp.margin_c = sg.margin_c;
    sp.margin_abs = sg.margin_abs;
    sp.qn_scale = h->tc_sigma * h->tc_sigma;
    sp.qnorm = ws.qnorm.as<float>();
    sp.thr = ws.thr.as<uint32_t>();
    sp.cand_key = nullptr;
    sp.cand_count = nullptr;
    sp.cap = 0;
    sp.trace = nullptr;
    sp.k = sg.k;
    sp.is_ip = h->metric == LIRA_METRIC_IP;
    if (getenv("LIRA_TC_NO_SEED")) return 0;
    if (sg.u8) {
        U8Params up = u8_params(h, ws, sg);
        up.exp = getenv("LIRA_TC_EXP") ? atoi(getenv("LIRA_TC_EXP")) & ~1 : 0;
        up.work_counter = sg.seed_counter;
        if (up.is_ip) u8_scan_kernel<true, true><<<h->num_sms, U8_THREADS, u8_smem_bytes(true), st>>>(*sg.tmap_q, h->tmap8, h->tmap_vaug8, h->tmap_aaug8, up);
        else u8_scan_kernel<true, false><<<h->num_sms, U8_THREADS, u8_smem_bytes(true), st>>>(*sg.tmap_q, h->tmap8, h->tmap_vaug8, h->tmap_aaug8, up);
        LIRA_LAUNCH_CHECK();
        return 0;
    }
    LIRA_CUDA_OK(launch_pdl(tc_scan_kernel<true, false>, dim3(h->num_sms), dim3(tc_threads(true)), TC_SMEM_BYTES, st, *sg.tmap_q, h->tmap16,
                            h->tmap_vaug, h->tmap_aaug, sp));
    LIRA_LAUNCH_CHECK();
    return 0;
}

// filter on the tensor cores + refine: ws.thr holds the bounds, the queries sit in ws.gq in group order
static int tc_filter_refine(lira_index* h, Workspace& ws, const TcStage& sg, cudaStream_t st) {
    const long long Q = sg.Q, P = sg.P;
    const int k = sg.k;
    const bool approx = h->tc_mode == 2;
    // one private candidate region per (pair, column part); every valid pair's owner writes its count
    // (byte scan: ONE region per pair, slots taken with atomics, counters zeroed beforehand)
    const int cap = sg.u8 ? (sg.private_regions ? (k <= TC_KMAX_TIGHTEN ? TC_CAPK : TC_CAPP) : (k <= TC_KMAX_TIGHTEN ? U8_CAPK : U8_CAPP))
                          : (k <= TC_KMAX_TIGHTEN ? TC_CAPK : TC_CAPP);   // fp16, k <= 16: full regions are compacted in the kernel
    const int parts = (sg.u8 && !sg.private_regions) ? 1 : TC_PARTS;
    h->last_parts = parts;
    h->last_u8 = sg.u8;
    if (int rc = ws.cand_key.ensure((size_t)P * parts * cap * 8)) return rc;
    if (int rc = ws.cand_count.ensure((size_t)P * parts * 4)) return rc;
    if (sg.u8 && !sg.private_regions && !sg.counts_zeroed) LIRA_CUDA_OK(cudaMemsetAsync(ws.cand_count.p, 0, (size_t)P * 4, st));
    TcParams tp;
    tp.group_queries = ws.group_queries.as<int>();
    tp.list_offsets = h->d_offsets;
    tp.items = ws.items.as<ScanItem>();
    tp.n_items = ws.n_items.as<int>();
    tp.work_counter = sg.filter_counter;
    tp.nk = (h->d16 + TC_KH - 1) / TC_KH;
    tp.max_rows = 0;
    tp.qnorm = ws.qnorm.as<float>();
    tp.thr = ws.thr.as<uint32_t>();
    tp.cand_key = ws.cand_key.as<unsigned long long>();
    tp.cand_count = ws.cand_count.as<int>();
    tp.cap = cap;
    tp.k = k;
    tp.is_ip = h->metric == LIRA_METRIC_IP;
    tp.margin_c = sg.margin_c;
    tp.margin_abs = sg.margin_abs;
    tp.qn_scale = h->tc_sigma * h->tc_sigma;
    tp.trace = nullptr;
    tp.exp = getenv("LIRA_TC_EXP") ? atoi(getenv("LIRA_TC_EXP")) : 0;
    if (getenv("LIRA_TC_TRACE")) {   // debug: per-chunk clock stamps of CTA 0 -> CSV
        if (int rc = ws.trace.ensure(((size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS + TC_TRACE_CTAS * 4 + (size_t)TC_TRACE_CTAS * TC_TRACE_ITEMS * 4) * 8)) return rc;
        LIRA_CUDA_OK(cudaMemsetAsync(ws.trace.p, 0, ((size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS + TC_TRACE_CTAS * 4 + (size_t)TC_TRACE_CTAS * TC_TRACE_ITEMS * 4) * 8, st));
        tp.trace = ws.trace.as<long long>();
    }
    if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[0], st));
    if (sg.u8) {
        U8Params up = u8_params(h, ws, sg);
        up.work_counter = sg.filter_counter;
        up.cand_key = tp.cand_key;
        up.cand_count = tp.cand_count;
        up.cap = cap;
        up.exp = tp.exp;
        if (getenv("LIRA_U8_TRACE")) {   // debug: clock stamps of CTA 0's first 512 units -> CSV (tools/u8_trace_summary.py)
            if (int rc = ws.trace.ensure(6 * 512 * 8)) return rc;
            LIRA_CUDA_OK(cudaMemsetAsync(ws.trace.p, 0, 6 * 512 * 8, st));
            up.trace = ws.trace.as<long long>();
        }
        if (up.is_ip) u8_scan_kernel<false, true><<<h->num_sms, U8_THREADS, u8_smem_bytes(false), st>>>(*sg.tmap_q, h->tmap8, h->tmap_vaug8, h->tmap_aaug8, up);
        else u8_scan_kernel<false, false><<<h->num_sms, U8_THREADS, u8_smem_bytes(false), st>>>(*sg.tmap_q, h->tmap8, h->tmap_vaug8, h->tmap_aaug8, up);
    } else
    if (tp.trace) tc_scan_kernel<false, true><<<h->num_sms, tc_threads(false), TC_SMEM_BYTES, st>>>(*sg.tmap_q, h->tmap16, h->tmap_vaug, h->tmap_aaug, tp);
    else LIRA_CUDA_OK(launch_pdl(tc_scan_kernel<false, false>, dim3(h->num_sms), dim3(tc_threads(false)), TC_SMEM_BYTES, st, *sg.tmap_q, h->tmap16,
                                 h->tmap_vaug, h->tmap_aaug, tp));
    LIRA_LAUNCH_CHECK();
    if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[1], st));
    // ---- refine ----
    RefineParams rp{ws.cand_key.as<unsigned long long>(), ws.cand_count.as<int>(), cap, sg.po, ws.probe_slot.as<int>(), h->ids, k,
                    (int)Q, sg.dedup, h->metric == LIRA_METRIC_IP, sg.d_D, sg.d_I, ws.redo.as<int>(), sg.redo_count,
                    h->vecs, (long long)h->ds, sg.d_q, sg.ldq, h->d, ws.qnorm.as<float>(), h->tc_sigma * h->tc_sigma, sg.margin_c, sg.margin_abs,
                    1.0f / (h->tc_sigma * h->tc_sigma)};
    const int warps = 8;
    if (parts == 1 && k <= 32) LIRA_CUDA_OK(launch_pdl(refine_topk_kernel<1, false, 1>, dim3((unsigned)((Q + warps - 1) / warps)), dim3(warps * 32), 0, st, rp));
    else if (parts == 1) LIRA_CUDA_OK(launch_pdl(refine_topk_kernel<4, false, 1>, dim3((unsigned)((Q + warps - 1) / warps)), dim3(warps * 32), 0, st, rp));
    else if (approx) LIRA_CUDA_OK(launch_pdl(refine_topk_kernel<1, true>, dim3((unsigned)((Q + warps - 1) / warps)), dim3(warps * 32), 0, st, rp));
    else if (k <= 32) LIRA_CUDA_OK(launch_pdl(refine_topk_kernel<1, false>, dim3((unsigned)((Q + warps - 1) / warps)), dim3(warps * 32), 0, st, rp));
    else LIRA_CUDA_OK(launch_pdl(refine_topk_kernel<4, false>, dim3((unsigned)((Q + warps - 1) / warps)), dim3(warps * 32), 0, st, rp));
    LIRA_LAUNCH_CHECK();
    if (h->timing) { LIRA_CUDA_OK(cudaEventRecord(h->ev[5], st)); h->scan_total_valid = true; }
    return 0;
}

static void tc_dump_trace(lira_index* h, Workspace& ws, const char* trace_path);
static void tc_debug_stats(lira_index* h, Workspace& ws, long long Q, long long P, int n_redo);

// The online query path on the tensor cores (tc_scan_kernels.cuh). *done = true when results were produced
// for every query whose ws.redo flag is 0; *n_redo counts the queries (flag 1) whose candidate buffer
// overflowed and that the caller must answer with the exact CUDA-core path. *done = false (no error) when
// the batch does not qualify at all.
static int tc_search(lira_index* h, const float* d_q, long long ldq, long long Q, const ProbeSpec& ps, int k, int dedup,
                     float* d_D, long long* d_I, int* d_nprobe, long long* d_cmp, bool* done, int* n_redo, bool* retry,
                     cudaStream_t st, bool allow_u8 = true, bool* u8_tried = nullptr) {
    *done = false;
    *n_redo = 0;
    *retry = false;
    if (!h->tc_ok || Q < 256) return 0;
    const bool approx = h->tc_mode == 2;
    if (approx && k > TC_KMAX_TIGHTEN) return 0;   // the margin logic lives in the compaction path (k <= 16)
    // byte-valued index: the integer tensor-core scan. Threshold / top-n / explicit probe sets need k <= 16 (per-list bounds of
    // the in-kernel seed); exhaustive probe sets (exact kNN over disjoint base segments) pool the seed candidates of the first
    // U8_SEED_LISTS lists, any k <= 128, as long as those lists hold enough candidates
    // (k > 16: twice the seed sample -- the tighter bound saves more in the filter pass than the longer seed pass costs: measured)
    const int u8_seed_lists = std::min(h->B, getenv("LIRA_U8_SEED_LISTS") ? std::max(1, std::min(32, atoi(getenv("LIRA_U8_SEED_LISTS"))))
                                                                          : (k > TC_KMAX_TIGHTEN ? 2 * U8_SEED_LISTS : U8_SEED_LISTS));
    // Measured (profiles/r2_u8_scan_notes.md): with its per-column norm subtraction the byte scan's epilogue costs ~3x the fp16
    // scan's, so for threshold / top-n probe sets it only matches the fp16 scan; it is the default for exhaustive probe sets
    // (no CUDA-core seed pass, half the operand bytes) and opt-in (LIRA_U8_SEARCH=1) for the rest.
    const bool use_u8 = h->u8_ok && allow_u8 &&
                        (ps.kind == 2 ? (u8_seed_lists * 16 >= k + 16 && h->E >= 2048) : (k <= TC_KMAX_TIGHTEN && getenv("LIRA_U8_SEARCH") != nullptr));
    if (u8_tried) *u8_tried = use_u8;
    if (use_u8) { if (int rc = ensure_u8_shadow(h)) return rc; }
    else if (int rc = ensure_fp16_shadow(h)) return rc;
    // exhaustive probe sets (exact kNN over base segments) with k > 16: no in-kernel tightening, so the static seed must be
    // good: the CUDA-core seed scans two whole segments (below); a segment shorter than that would make the bound loose
    if (!use_u8 && ps.kind == 2 && k > TC_KMAX_TIGHTEN && h->E < 4 * 8192) return 0;
    LIRA_REQUIRE((ldq % 4) == 0 && ((uintptr_t)d_q & 15) == 0, "queries must be 16-byte aligned with ld % 4 == 0");
    Workspace& ws = h->ws;
    if (int rc = ws.qnorm.ensure((size_t)Q * 4)) return rc;
    if (int rc = ws.flags.ensure(64)) return rc;
    if (int rc = ws.redo.ensure((size_t)Q * 4)) return rc;
    // [0] query batch exactly representable, [1] number of overflowed queries, [2] a query's probe set was truncated
    // [3] an explicit probe set named a list outside [0, B)
    static const int one_zero[4] = {1, 0, 0, 0};
    LIRA_CUDA_OK(cudaMemcpyAsync(ws.flags.p, one_zero, 16, cudaMemcpyHostToDevice, st));
    const bool optimistic = ps.kind == 0 && !h->tc_force_sync;   // no host round trip before the scan (checked at the end)
    h->tc_force_sync = false;
    // |q|^2; exact mode: flags[0] is cleared unless the batch is exact in fp16 (approximate mode: the gather below clears it
    // when a scaled query value does not fit fp16)
    row_norms_kernel<<<grid_for(Q, 128), 128, 0, st>>>(d_q, ldq, h->ds, Q, ws.qnorm.as<float>(), approx ? nullptr : ws.flags.as<int>(), nullptr);
    LIRA_LAUNCH_CHECK();
    long long P = 0;
    const long long* po = nullptr;
    int q_exact = 1;
    ProbeSpec psf = ps;
    psf.d_bad_flag = ws.flags.as<int>() + 3;
    // (byte scan: items of up to 512 queries x a segment of the list's rows; exhaustive probe sets keep whole lists, since
    //  their pooled seed addresses one set of values per (query, list) pair)
    const int u8_seg_rows = ps.kind == 2 ? (1 << 30) : U8_SEG_ROWS;
    if (int rc = prepare_groups(h, ws, Q, psf, use_u8 ? h->u8_item_q() : TC_M, d_cmp, &P, &po, &q_exact, ws.flags.as<int>(), nullptr, st,
                                optimistic ? ws.flags.as<int>() + 2 : nullptr, use_u8 ? u8_seg_rows : 0)) return rc;
    if (!q_exact || P == 0) return 0;  // not exact in fp16 (or nothing probed): exact CUDA-core path
    if (int rc = save_stats(h, ws, st)) return rc;
    if (ps.kind == 1) {
        first_probes_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ps.d_probe_offsets, ps.d_probe_ids, (int)Q, ws.top1.as<int>());
        LIRA_LAUNCH_CHECK();
    } else if (ps.kind == 2) {   // every query probes every list: seed with the first (two)
        seed_first_lists_kernel<<<grid_for(Q, 256), 256, 0, st>>>((int)Q, h->B, ws.top1.as<int>());
        LIRA_LAUNCH_CHECK();
    }
    if (d_nprobe) {
        copy_nprobe_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ws.nsel.as<int>(), d_nprobe, (int)Q);
        LIRA_LAUNCH_CHECK();
    }
    if (int rc = ws.thr.ensure((size_t)Q * 4)) return rc;
    const int nk = (h->d16 + TC_KH - 1) / TC_KH;
    // L2: the gathered fp16 query rows carry the factor 2 of  s = 2 q.v - |v|^2  (exact: integers of <= 11 bits, doubled)
    const bool is_ip = h->metric == LIRA_METRIC_IP;
    const float qscale = (is_ip ? 1.0f : 2.0f) * h->tc_sigma;
    float margin_c = 0.f, margin_abs = 0.f;
    tc_margins(h, &margin_c, &margin_abs);
    int* d_ok = approx ? ws.flags.as<int>() : nullptr;
    Workspace& sw = h->ws_seed;
    // ---- queries in group order (one TMA box per tile) ----
    CUtensorMap tmap_q;
    if (use_u8) {
        // (the flag is CLEARED when a query value is not an integer in [0, 255]: the batch then takes the fp16 route)
        // (exhaustive probe sets: the first Q slots are the batch in query order, and one copy of the rows serves every list)
        const long long Pg = ps.kind == 2 ? Q : P;
        if (int rc = ws.gq.ensure((size_t)(Pg + U8_ITEM_Q) * h->d8)) return rc;
        gather_group_queries_u8_kernel<<<grid_for(Pg * (h->d8 / 4), 256, 148 * 16), 256, 0, st>>>(
            d_q, ldq, h->ds, ws.group_queries.as<int>(), Pg, ws.group_offsets.as<long long>() + h->B, ws.gq.as<uint8_t>(), h->d8, ws.flags.as<int>());
        LIRA_LAUNCH_CHECK();
        if (int rc = make_tmap_u8(&tmap_q, ws.gq.p, Pg, h->d8, h->d8)) return rc;
    } else {
        if (int rc = ws.gq.ensure((size_t)(P + TC_M) * h->d16 * 2)) return rc;
        gather_group_queries_kernel<<<grid_for(P * (h->d16 / 4), 256, 148 * 16), 256, 0, st>>>(d_q, ldq, h->ds, ws.group_queries.as<int>(), P,
                                                                                             ws.group_offsets.as<long long>() + h->B, ws.gq.as<__half>(), h->d16, qscale, d_ok);
        LIRA_LAUNCH_CHECK();
        if (int rc = make_tmap_f16(&tmap_q, ws.gq.as<__half>(), P, h->d16, h->d16)) return rc;
    }
    // Seed, k <= 16 and probe sets that are a small part of the index: the seed pass runs on the SAME work items as the filter
    // pass (no second grouping, no second gather) and EVERY (query, list) pair takes part: the k-th smallest of the 64 group
    // minima of a query's row in a list bounds the final k-th score, the smallest over the query's lists wins (atomicMin).
    // Exhaustive probe sets (exact kNN) keep the two-segment seed below: a full extra pass would double their work.
    const bool seed_on_main = k <= TC_KMAX_TIGHTEN && ps.kind != 2 && (use_u8 || !getenv("LIRA_TC_SEED_SEPARATE"));
    if (seed_on_main) {
        fill_u32_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ws.thr.as<uint32_t>(), Q, 0xFF800000u /* f32_to_ordered(+inf) */);
        LIRA_LAUNCH_CHECK();
        TcStage ss;
        ss.k = k; ss.tmap_q = &tmap_q; ss.margin_c = margin_c; ss.margin_abs = margin_abs; ss.u8 = use_u8; ss.seg_rows = u8_seg_rows;
        ss.seed_counter = ws.n_items.as<int>() + 2;   // (prepare_groups zeroes the whole control block)
        if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[4], st));
        if (int rc = tc_seed_main(h, ws, ss, st)) return rc;
    } else if (use_u8) {
        // ---- exhaustive probe sets on the byte copy: seed pass over the first lists only, candidates pooled per query ----
        const int S = u8_seed_lists;
        const size_t max_items = (size_t)S * ((size_t)(Q + h->u8_item_q() - 1) / h->u8_item_q()) + 1;
        if (int rc = ws.seed_items.ensure(max_items * sizeof(ScanItem))) return rc;
        if (int rc = ws.seed_out.ensure((size_t)S * Q * U8_PARTS * 4 * 4)) return rc;
        int* ctl = ws.n_items.as<int>();   // [4] seed items, [2] seed tickets (zeroed by prepare_groups)
        if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[4], st));
        u8_seed_items_kernel<<<grid_for((long long)h->B * ((Q + h->u8_item_q() - 1) / h->u8_item_q()), 256), 256, 0, st>>>(
            ws.items.as<ScanItem>(), ctl, S, ws.seed_items.as<ScanItem>(), ctl + 4);
        LIRA_LAUNCH_CHECK();
        TcStage ss;
        ss.k = k; ss.u8 = true; ss.seg_rows = u8_seg_rows; ss.q_mod = (int)Q;
        U8Params up = u8_params(h, ws, ss);
        up.items = ws.seed_items.as<ScanItem>();
        up.n_items = ctl + 4;
        up.work_counter = ctl + 2;
        up.seed_out = ws.seed_out.as<int>();
        if (!getenv("LIRA_TC_NO_SEED")) {
            if (up.is_ip) u8_scan_kernel<true, true><<<h->num_sms, U8_THREADS, u8_smem_bytes(true), st>>>(tmap_q, h->tmap8, h->tmap_vaug8, h->tmap_aaug8, up);
            else u8_scan_kernel<true, false><<<h->num_sms, U8_THREADS, u8_smem_bytes(true), st>>>(tmap_q, h->tmap8, h->tmap_vaug8, h->tmap_aaug8, up);
            LIRA_LAUNCH_CHECK();
            u8_seed_select_kernel<<<(int)((Q + 7) / 8), 256, 0, st>>>(ws.seed_out.as<int>(), (int)Q, S, k, ws.thr.as<uint32_t>());
            LIRA_LAUNCH_CHECK();
        } else {
            fill_u32_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ws.thr.as<uint32_t>(), Q, 0xFF800000u);
            LIRA_LAUNCH_CHECK();
        }
    } else if (k <= TC_KMAX_TIGHTEN) {
        // ---- seed on the tensor cores: first rows of every query's best list, 16 group minima per row ----
        if (int rc = sw.probe_offsets.ensure((size_t)(Q + 1) * 8)) return rc;
        if (int rc = sw.probe_ids.ensure((size_t)(Q + 1) * 4)) return rc;
        // seed lists per query: the best one, or (LIRA_TC_SEED_LISTS=2) the two best -- every list alone yields k real
        // candidates, so the smaller of the per-list bounds (atomicMin in the kernel) is valid
        const int seed_lists = getenv("LIRA_TC_SEED_LISTS") ? std::max(1, std::min(2, atoi(getenv("LIRA_TC_SEED_LISTS")))) : 2;
        iota_offsets_kernel<<<grid_for(Q + 1, 256), 256, 0, st>>>(sw.probe_offsets.as<long long>(), Q, seed_lists);
        LIRA_LAUNCH_CHECK();
        if (seed_lists == 1) {
            first_of_pairs_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ws.top1.as<int>(), (int)Q, sw.probe_ids.as<int>());
            LIRA_LAUNCH_CHECK();
        }
        fill_u32_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ws.thr.as<uint32_t>(), Q, 0xFF800000u /* f32_to_ordered(+inf) */);
        LIRA_LAUNCH_CHECK();
        ProbeSpec seed;
        seed.kind = 1;
        seed.d_probe_offsets = sw.probe_offsets.as<long long>();
        seed.d_probe_ids = seed_lists == 1 ? sw.probe_ids.as<int>() : ws.top1.as<int>();
        seed.P = Q * seed_lists;
        long long Pseed = 0;
        const long long* po_seed = nullptr;
        if (int rc = prepare_groups(h, sw, Q, seed, TC_M, nullptr, &Pseed, &po_seed, nullptr, nullptr, nullptr, st)) return rc;
        if (int rc = sw.gq.ensure((size_t)(Pseed + TC_M) * h->d16 * 2)) return rc;
        gather_group_queries_kernel<<<grid_for(Pseed * (h->d16 / 4), 256, 148 * 16), 256, 0, st>>>(
            d_q, ldq, h->ds, sw.group_queries.as<int>(), Pseed, sw.group_offsets.as<long long>() + h->B, sw.gq.as<__half>(), h->d16, qscale, d_ok);
        LIRA_LAUNCH_CHECK();
        CUtensorMap tmap_sq;
        if (int rc = make_tmap_f16(&tmap_sq, sw.gq.as<__half>(), Pseed, h->d16, h->d16)) return rc;
        TcParams sp;
        sp.group_queries = sw.group_queries.as<int>();
        sp.list_offsets = h->d_offsets;
        sp.items = sw.items.as<ScanItem>();
        sp.n_items = sw.n_items.as<int>();
        sp.work_counter = sw.n_items.as<int>() + 1;
        sp.nk = nk;
        sp.max_rows = getenv("LIRA_TC_SEED_ROWS") ? atoi(getenv("LIRA_TC_SEED_ROWS")) : (ps.kind == 2 ? 8192 : TC_SEED_ROWS_TC);   // 0: whole lists
        sp.exp = 0;
        sp.margin_c = margin_c;
        sp.margin_abs = margin_abs;
        sp.qn_scale = h->tc_sigma * h->tc_sigma;
        sp.qnorm = ws.qnorm.as<float>();
        sp.thr = ws.thr.as<uint32_t>();
        sp.cand_key = nullptr;
        sp.cand_count = nullptr;
        sp.cap = 0;
        sp.trace = nullptr;
        sp.k = k;
        sp.is_ip = h->metric == LIRA_METRIC_IP;
        // (LIRA_TC_NO_SEED, tests only: every bound starts at +inf, so the in-kernel region compaction does all the work)
        if (!getenv("LIRA_TC_NO_SEED")) {
            tc_scan_kernel<true, false><<<h->num_sms, tc_threads(true), TC_SMEM_BYTES, st>>>(tmap_sq, h->tmap16, h->tmap_vaug, h->tmap_aaug, sp);
            LIRA_LAUNCH_CHECK();
        }
    } else {
        // ---- seed on the CUDA cores (k > 16): exact scan of the first rows of the two best lists ----
        if (int rc = sw.probe_offsets.ensure((size_t)(Q + 1) * 8)) return rc;
        iota_offsets_kernel<<<grid_for(Q + 1, 256), 256, 0, st>>>(sw.probe_offsets.as<long long>(), Q, 2);
        LIRA_LAUNCH_CHECK();
        ProbeSpec seed;
        seed.kind = 1;
        seed.d_probe_offsets = sw.probe_offsets.as<long long>();
        seed.d_probe_ids = ws.top1.as<int>();
        seed.P = 2 * Q;
        long long Pseed = 0;
        const long long* po_seed = nullptr;
        if (int rc = prepare_groups(h, sw, Q, seed, SCAN_TM_MAX, nullptr, &Pseed, &po_seed, nullptr, nullptr, nullptr, st)) return rc;
        // (exact kNN: the first two segments whole -- the k-th best of 16 384 base rows lets k N / 16 384 rows per query through)
        if (int rc = simt_scan(h, sw, d_q, ldq, Pseed, k, 0, ps.kind == 2 ? 8192 : TC_SEED_ROWS, false, st)) return rc;
        seed_threshold_kernel<<<grid_for(Q, 128), 128, 0, st>>>(sw.part_key.as<unsigned long long>(), sw.probe_slot.as<int>(),
                                                                ws.top1.as<int>(), (int)Q, k, ws.thr.as<uint32_t>());
        LIRA_LAUNCH_CHECK();
    }
    if (!seed_on_main && !use_u8 && h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[4], st));   // (the other seeds' own grouping is not part of the scan time)
    TcStage sg;
    sg.u8 = use_u8; sg.seg_rows = u8_seg_rows; sg.private_regions = use_u8 && ps.kind == 2; sg.q_mod = (use_u8 && ps.kind == 2) ? (int)Q : 0;
    sg.Q = Q; sg.P = P; sg.po = po; sg.k = k; sg.dedup = dedup; sg.d_q = d_q; sg.ldq = ldq; sg.d_D = d_D; sg.d_I = d_I;
    sg.redo_count = ws.flags.as<int>() + 1; sg.tmap_q = &tmap_q; sg.margin_c = margin_c; sg.margin_abs = margin_abs;
    sg.filter_counter = ws.n_items.as<int>() + 1;
    if (int rc = tc_filter_refine(h, ws, sg, st)) return rc;
    const char* trace_path = getenv("LIRA_TC_TRACE");
    if (!h->h_flags) LIRA_CUDA_OK(cudaHostAlloc((void**)&h->h_flags, 64, cudaHostAllocDefault));
    int* fl = h->h_flags;
    LIRA_CUDA_OK(cudaMemcpyAsync(fl, ws.flags.p, 16, cudaMemcpyDeviceToHost, st));
    LIRA_CUDA_OK(cudaStreamSynchronize(st));
    LIRA_REQUIRE(fl[3] == 0, "probed list id out of range");
    *n_redo = fl[1];
    if ((optimistic && (fl[0] == 0 || fl[2] != 0)) || ((approx || use_u8) && fl[0] == 0)) {
        // the optimistic run is void: the batch is not exact in fp16 (-> the caller's CUDA-core path), or a probe set was
        // truncated (-> once more with the exact pair count and a larger cap from now on)
        if (fl[2] != 0) { h->nprobe_cap = std::min(h->B, h->nprobe_cap * 4); *retry = fl[0] != 0; }
        h->tc_force_sync = true;
        *n_redo = 0;
        return 0;
    }
    if (trace_path) tc_dump_trace(h, ws, trace_path);
    if (getenv("LIRA_U8_TRACE") && use_u8) {
        std::vector<long long> tr(6 * 512);
        cudaMemcpy(tr.data(), ws.trace.p, tr.size() * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(getenv("LIRA_U8_TRACE"), "w")) {
            fprintf(f, "unit,mma_acc_free,mma_b_ready,mma_issued,epi_ready,epi_loaded,epi_done\n");
            for (int u = 0; u < 512; ++u)
                fprintf(f, "%d,%lld,%lld,%lld,%lld,%lld,%lld\n", u, tr[u], tr[512 + u], tr[1024 + u], tr[1536 + u], tr[2048 + u], tr[2560 + u]);
            fclose(f);
        }
    }
    if (getenv("LIRA_DEBUG")) tc_debug_stats(h, ws, Q, P, *n_redo);
    h->last_path = 1;
    h->last_redo = *n_redo;
    *done = true;
    return 0;
}

static void tc_dump_trace(lira_index* h, Workspace& ws, const char* trace_path) {
        std::vector<long long> tr((size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS + TC_TRACE_CTAS * 4 + (size_t)TC_TRACE_CTAS * TC_TRACE_ITEMS * 4);
        cudaMemcpy(tr.data(), ws.trace.p, tr.size() * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "chunk,prod_start,mma_acc_free,mma_b0,mma_blast,mma_issued,epi4_ready,epi4_done,epi8_ready,epi8_done\n");
            for (int c = 0; c < TC_TRACE_CHUNKS; ++c) {
                fprintf(f, "%d", c);
                for (int r = 0; r < TC_TRACE_ROLES; ++r) fprintf(f, ",%lld", tr[(size_t)r * TC_TRACE_CHUNKS + c]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        if (FILE* f = fopen((std::string(trace_path) + ".ctas").c_str(), "w")) {   // per-CTA span and work (load balance)
            fprintf(f, "cta,start_ns,end_ns,chunks,items\n");
            const long long* c = tr.data() + (size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS;
            for (int b = 0; b < std::min(h->num_sms, TC_TRACE_CTAS); ++b)
                fprintf(f, "%d,%lld,%lld,%lld,%lld\n", b, c[b * 4], c[b * 4 + 1], c[b * 4 + 2], c[b * 4 + 3]);
            fclose(f);
        }
        if (FILE* f = fopen((std::string(trace_path) + ".items").c_str(), "w")) {   // per-CTA item timeline
            fprintf(f, "cta,n,start_ns,list,rows,queries\n");
            const long long* c = tr.data() + (size_t)TC_TRACE_ROLES * TC_TRACE_CHUNKS + TC_TRACE_CTAS * 4;
            for (int b = 0; b < std::min(h->num_sms, TC_TRACE_CTAS); ++b)
                for (int n = 0; n < TC_TRACE_ITEMS; ++n) {
                    const long long* e = c + ((size_t)b * TC_TRACE_ITEMS + n) * 4;
                    if (e[0]) fprintf(f, "%d,%d,%lld,%lld,%lld,%lld\n", b, n, e[0], e[1], e[2], e[3]);
                }
            fclose(f);
        }
}

static void tc_debug_stats(lira_index* h, Workspace& ws, long long Q, long long P, int n_redo) {
    std::vector<uint32_t> th((size_t)Q);
    cudaMemcpy(th.data(), ws.thr.p, (size_t)Q * 4, cudaMemcpyDeviceToHost);
    long long n_inf = 0;
    for (uint32_t t : th) n_inf += (t == 0xFF800000u);
    fprintf(stderr, "[lira] tc batch: final bound still +inf for %lld queries\n", n_inf);
    int dbg[6] = {0, 0, 0, 0, 0, 0};
    cudaMemcpy(dbg, ws.n_items.as<int>() + 10, 24, cudaMemcpyDeviceToHost);
    if (dbg[1]) fprintf(stderr, "[lira] tc batch: %d of %d (warp, 32 columns) groups held a survivor; %d items, %d query tiles, %d chunks, %d units\n",
                        dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[5]);
    long long n_slots = P;   // P may be an upper bound (fused flow): the valid pairs are the first group_offsets[B] slots
    cudaMemcpy(&n_slots, ws.group_offsets.as<long long>() + h->B, 8, cudaMemcpyDeviceToHost);
    P = std::min(P, n_slots);
    if (P <= 0) return;
    const int parts = h->last_parts;
    std::vector<int> cc((size_t)P * parts);
    cudaMemcpy(cc.data(), ws.cand_count.p, (size_t)P * parts * 4, cudaMemcpyDeviceToHost);
    std::vector<int> sorted(cc);
    std::sort(sorted.begin(), sorted.end());
    long long tot = 0, over32 = 0, over64 = 0;
    for (int c : cc) { tot += c; over32 += c > 32; over64 += c > 64; }
    fprintf(stderr, "[lira] tc batch Q=%lld P=%lld: survivors/query mean %.1f; per (pair, part) p50 %d p99 %d max %d; regions > 32: %lld, > 64: %lld; redo %d\n",
            Q, P, (double)tot / Q, sorted[P * parts / 2], sorted[(size_t)(P * parts * 0.99)], sorted[P * parts - 1], over32, over64, n_redo);
}

// exact CUDA-core path: the whole batch (done = false), or only the queries the tensor-core pass flagged in ws.redo
static int exact_fallback(lira_index* h, const float* d_q, long long ldq, long long Q, const ProbeSpec& ps, int k, int dedup,
                          float* d_D, long long* d_I, int* d_nprobe, long long* d_cmp, cudaStream_t st, bool done) {
    const int* mask = done ? h->ws.redo.as<int>() : nullptr;
    if (!done) h->last_path = 0;
    const long long* po = nullptr;
    long long P = 0;
    if (int rc = run_grouped_scan(h, d_q, ldq, Q, ps, k, /*store_local=*/0, done ? nullptr : d_cmp, &P, &po, st, mask, !done)) return rc;
    if (Q == 0) return 0;
    Workspace& ws = h->ws;
    MergeParams mp;
    mp.part_key = ws.part_key.as<unsigned long long>();
    mp.probe_offsets = po;
    mp.probe_slot = ws.probe_slot.as<int>();
    mp.k = k;
    mp.Q = (int)Q;
    mp.dedup = dedup;
    mp.out_dist = d_D;
    mp.out_ids = d_I;
    mp.is_ip = h->metric == LIRA_METRIC_IP;
    mp.mask = mask;
    if (int rc = launch_merge(mp, st)) return rc;
    if (d_nprobe && !done) {
        copy_nprobe_kernel<<<grid_for(Q, 256), 256, 0, st>>>(ws.nsel.as<int>(), d_nprobe, (int)Q);
        LIRA_LAUNCH_CHECK();
    }
    return 0;
}

static int search_core(lira_index* h, const float* d_q, long long ldq, long long Q, const ProbeSpec& ps, int k,
                       int dedup, float* d_D, long long* d_I, int* d_nprobe, long long* d_cmp, cudaStream_t st) {
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    h->last_Q = Q;
    h->last_k = k;
    h->last_redo = 0;
    if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[2], st));
    bool done = false;
    int n_redo = 0;
    if (h->use_tc) {
        bool retry = false, u8_tried = false;
        if (int rc = tc_search(h, d_q, ldq, Q, ps, k, dedup, d_D, d_I, d_nprobe, d_cmp, &done, &n_redo, &retry, st, true, &u8_tried)) return rc;
        if (!done && retry)
            if (int rc = tc_search(h, d_q, ldq, Q, ps, k, dedup, d_D, d_I, d_nprobe, d_cmp, &done, &n_redo, &retry, st, true, &u8_tried)) return rc;
        if (!done && u8_tried) {   // e.g. a query batch with values outside [0, 255]: the fp16 route may still serve it
            h->tc_force_sync = true;
            if (int rc = tc_search(h, d_q, ldq, Q, ps, k, dedup, d_D, d_I, d_nprobe, d_cmp, &done, &n_redo, &retry, st, false)) return rc;
            if (!done && retry)
                if (int rc = tc_search(h, d_q, ldq, Q, ps, k, dedup, d_D, d_I, d_nprobe, d_cmp, &done, &n_redo, &retry, st, false)) return rc;
        }
    }
    if (!done || n_redo > 0)
        if (int rc = exact_fallback(h, d_q, ldq, Q, ps, k, dedup, d_D, d_I, d_nprobe, d_cmp, st, done)) return rc;
    if (h->timing) {
        LIRA_CUDA_OK(cudaEventRecord(h->ev[3], st));
        h->timing_pending = true;
    }
    return 0;
}

// The whole query phase of one batch with the front end fused (fused_probe_kernels.cuh): prep -> model forward whose last
// layer selects -> finish_select -> scatter_queries -> seed -> filter -> refine; no host round trip before the final
// status read. Takes the batch only when the tensor-core scan with in-kernel tightening serves it (k <= 16, >= 256
// queries, threshold selection); *done = false (and nothing written) otherwise, or when the optimistic assumptions did
// not hold (batch not exact in fp16, or more than nprobe_cap partitions selected for a query even after raising the cap).
// On success ws.probe_offsets / ws.probe_ids hold the probe sets as a CSR (the exact redo of flagged queries uses it).
// a batch that ran with `ran_with` partitions per query at most saw a query selecting `seen_max`: the next ones get room for it
// (several truncated batches may be in flight: only the first to be checked moves the cap)
static void raise_nprobe_cap(lira_index* h, int ran_with, int seen_max) {
    if (ran_with < h->nprobe_cap) return;
    const int want = std::max(ran_with + 32, (seen_max + 31) / 32 * 32);
    h->nprobe_cap = std::min(h->B, want);
}

static bool fused_eligible(const lira_index* h, const lira_model* m, long long Q, int mode, int k) {
    return h->use_tc && m->use_tc && h->tc_ok && Q >= 256 && k <= TC_KMAX_TIGHTEN && mode != LIRA_SELECT_TOPN && !getenv("LIRA_NO_FUSED") &&
           Q * (long long)std::min(h->B, h->nprobe_cap) <= (4ll << 20);
}

// Enqueues the fused flow of one batch on `st` (every launch, no host wait) and the copy of its four status words into
// `h_flags` (pinned). The caller decides when to wait: fused_status() reads the words once the stream has passed them.
static int fused_enqueue(lira_index* h, lira_model* m, const float* d_q, long long ldq, long long Q, int mode, double value,
                         int k, int dedup, float* d_D, long long* d_I, int* d_nprobe, long long* d_cmp, cudaStream_t st, int* h_flags) {
    LIRA_REQUIRE(mode == LIRA_SELECT_GT || mode == LIRA_SELECT_GE_ARGMAX, "unknown selection mode");
    LIRA_REQUIRE((ldq % 4) == 0 && ((uintptr_t)d_q & 15) == 0, "queries must be 16-byte aligned with ld % 4 == 0");
    const int B = h->B;
    const bool approx = h->tc_mode == 2;
    const bool u8 = h->u8_ok && getenv("LIRA_U8_SEARCH") != nullptr;
    if (u8) { if (int rc = ensure_u8_shadow(h)) return rc; }
    else if (int rc = ensure_fp16_shadow(h)) return rc;
    Workspace& ws = h->ws;
    const int cap = std::min(B, h->nprobe_cap);
    const long long P = Q * (long long)cap;   // bound: sizes the per-pair buffers
    if (int rc = model_ensure_tc(m, Q)) return rc;
    for (DevBuf* b : {&ws.qnorm, &ws.thr, &ws.redo})
        if (int rc = b->ensure((size_t)Q * 4)) return rc;
    if (int rc = ws.nsel.ensure((size_t)(Q + 1) * 4)) return rc;
    if (int rc = ws.top1.ensure((size_t)(Q + 1) * 8)) return rc;        // per-query argmax keys of the selection
    if (int rc = ws.sel.ensure((size_t)P * 4)) return rc;
    if (int rc = ws.probe_ids.ensure((size_t)P * 4 + 64)) return rc;
    if (int rc = ws.probe_offsets.ensure((size_t)(Q + 1) * 8)) return rc;
    if (int rc = ws.group_offsets.ensure((size_t)(B + 1) * 8)) return rc;
    if (int rc = ws.list_count.ensure((size_t)(B + 1) * 4)) return rc;
    if (int rc = ws.cursor.ensure((size_t)(B + 1) * 4)) return rc;
    if (int rc = ws.n_items.ensure(128)) return rc;
    if (int rc = ws.group_queries.ensure((size_t)P * 4)) return rc;
    if (int rc = ws.probe_slot.ensure((size_t)P * 4)) return rc;
    size_t max_items = (size_t)P / 8 + B + 1;
    if (u8) max_items = std::max(max_items, ((size_t)P / h->u8_item_q() + 1) * (size_t)h->u8_max_nseg + (size_t)h->u8_nseg_total + B + 1);
    if (int rc = ws.items.ensure(max_items * sizeof(ScanItem))) return rc;
    if (u8)
        if (int rc = ws.cand_count.ensure((size_t)P * 4)) return rc;
    if (int rc = ws.gq.ensure(u8 ? (size_t)(P + U8_ITEM_Q) * h->d8 : (size_t)(P + TC_M) * h->d16 * 2)) return rc;
    // control block: [0] work items, [1] filter tickets, [2] seed tickets, +64 B {E_p, pairs}, +96 B status words:
    // [24] batch not exact in fp16, [25] queries left for the exact path, [26] a probe set was truncated, [27] largest selection of a query
    LIRA_CUDA_OK(cudaMemsetAsync(ws.n_items.p, 0, 128, st));
    int* ctl = ws.n_items.as<int>();
    int* fl_dev = ctl + 24;
    PrepParams pp{d_q, (long)ldq, m->d, m->ds, Q, m->mu, m->qch.as<float>(), m->qcl.as<float>(), m->qn.as<float>(),
                  m->qrh.as<float>(), m->qrl.as<float>(), ws.qnorm.as<float>(), approx ? nullptr : fl_dev, ws.thr.as<uint32_t>(),
                  ws.nsel.as<int>(), ws.top1.as<unsigned long long>(), ws.list_count.as<int>(), ws.cursor.as<int>(), B};
    const int warps = 8;
    const int qgrid = (int)((Q + warps - 1) / warps);
    prep_queries_kernel<<<qgrid, warps * 32, 0, st>>>(pp);
    LIRA_LAUNCH_CHECK();
    FusedSelect fs{ws.sel.as<int>(), ws.nsel.as<int>(), ws.list_count.as<int>(), ws.top1.as<unsigned long long>(), cap,
                   mode == LIRA_SELECT_GE_ARGMAX ? 1 : 0, (float)value};
    if (int rc = model_forward_tc(m, d_q, ldq, Q, nullptr, 0, nullptr, 0, st, /*prepped=*/true, &fs)) return rc;
    FinishSelectParams fp{ws.nsel.as<int>(), ws.sel.as<int>(), cap, fs.mode, ws.top1.as<unsigned long long>(), ws.list_count.as<int>(),
                          (int)Q, B, ws.probe_offsets.as<long long>(), ws.group_offsets.as<long long>(), fl_dev + 2, h->d_list_order,
                          h->d_offsets, u8 ? h->u8_item_q() : TC_M, u8 ? U8_SEG_ROWS : 0, ws.items.as<ScanItem>(), ctl, (unsigned long long*)((char*)ws.n_items.p + 64)};
    LIRA_CUDA_OK(launch_pdl(finish_select_kernel, dim3(fs.mode == 0 ? 2 : 1), dim3(1024), 0, st, fp));
    LIRA_LAUNCH_CHECK();
    if (int rc = save_stats(h, ws, st)) return rc;
    const bool is_ip = h->metric == LIRA_METRIC_IP;
    const float qscale = (is_ip ? 1.0f : 2.0f) * h->tc_sigma;
    ScatterQueriesParams sc{d_q, (long)ldq, h->ds, ws.sel.as<int>(), ws.nsel.as<int>(), cap, ws.probe_offsets.as<long long>(),
                            ws.group_offsets.as<long long>(), h->d_offsets, ws.cursor.as<int>(), ws.group_queries.as<int>(),
                            ws.probe_slot.as<int>(), ws.probe_ids.as<int>(), ws.gq.as<__half>(), h->d16, qscale, u8 ? ws.cand_count.as<int>() : nullptr, u8 ? ws.gq.as<uint8_t>() : nullptr, h->d8,
                            (approx || u8) ? fl_dev : nullptr,
                            d_nprobe, d_cmp, (int)Q};
    LIRA_CUDA_OK(launch_pdl(scatter_queries_kernel, dim3(qgrid), dim3(warps * 32), 0, st, sc));
    LIRA_LAUNCH_CHECK();
    CUtensorMap tmap_q;
    if (u8) { if (int rc = make_tmap_u8(&tmap_q, ws.gq.p, P, h->d8, h->d8)) return rc; }
    else if (int rc = make_tmap_f16(&tmap_q, ws.gq.as<__half>(), P, h->d16, h->d16)) return rc;
    TcStage sg;
    sg.Q = Q; sg.P = P; sg.po = ws.probe_offsets.as<long long>(); sg.k = k; sg.dedup = dedup; sg.d_q = d_q; sg.ldq = ldq;
    sg.d_D = d_D; sg.d_I = d_I; sg.redo_count = fl_dev + 1; sg.tmap_q = &tmap_q; sg.u8 = u8; sg.counts_zeroed = u8;
    tc_margins(h, &sg.margin_c, &sg.margin_abs);
    sg.seed_counter = ctl + 2;
    sg.filter_counter = ctl + 1;
    if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[4], st));
    if (int rc = tc_seed_main(h, ws, sg, st)) return rc;
    if (int rc = tc_filter_refine(h, ws, sg, st)) return rc;
    LIRA_CUDA_OK(cudaMemcpyAsync(h_flags, fl_dev, 16, cudaMemcpyDeviceToHost, st));
    return 0;
}

// The whole query phase of one batch with the front end fused (fused_probe_kernels.cuh): prep -> model forward whose last
// layer selects -> finish_select -> scatter_queries -> seed -> filter -> refine, then ONE host wait for the status words.
// Takes the batch only when the tensor-core scan with in-kernel tightening serves it (k <= 16, >= 256 queries, threshold
// selection); *done = false otherwise, or when the optimistic assumptions did not hold (batch not exact in fp16, or more
// than nprobe_cap partitions selected for a query even after raising the cap: the caller's unfused flow answers).
// On success ws.probe_offsets / ws.probe_ids hold the probe sets as a CSR (the exact redo of flagged queries uses it).
static int fused_probe_search(lira_index* h, lira_model* m, const float* d_q, long long ldq, long long Q, int mode, double value,
                              int k, int dedup, float* d_D, long long* d_I, int* d_nprobe, long long* d_cmp, cudaStream_t st,
                              bool* done, int* n_redo) {
    *done = false;
    *n_redo = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (!fused_eligible(h, m, Q, mode, k)) return 0;
        if (!h->h_flags) LIRA_CUDA_OK(cudaHostAlloc((void**)&h->h_flags, 64, cudaHostAllocDefault));
        int* fl = h->h_flags;
        if (int rc = fused_enqueue(h, m, d_q, ldq, Q, mode, value, k, dedup, d_D, d_I, d_nprobe, d_cmp, st, fl)) return rc;
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        if (getenv("LIRA_TC_TRACE")) tc_dump_trace(h, h->ws, getenv("LIRA_TC_TRACE"));
        if (fl[0] != 0) return 0;                       // batch not exact in fp16 (or out of fp16 range): the unfused flow decides
        if (fl[2] != 0) {                               // a query selected more than the cap: raise it, once more
            if (h->nprobe_cap >= h->B) return 0;
            raise_nprobe_cap(h, h->nprobe_cap, fl[3]);
            continue;
        }
        if (getenv("LIRA_DEBUG")) tc_debug_stats(h, h->ws, Q, Q * (long long)std::min(h->B, h->nprobe_cap), fl[1]);
        *n_redo = fl[1];
        h->last_path = 1;
        h->last_redo = fl[1];
        *done = true;
        return 0;
    }
    return 0;
}

static int finish_timing(lira_index* h) {
    if (!h->timing || !h->timing_pending) return 0;
    h->timing_pending = false;
    LIRA_CUDA_OK(cudaEventSynchronize(h->ev[3]));
    LIRA_CUDA_OK(cudaEventElapsedTime(&h->last_scan_ms, h->ev[0], h->ev[1]));
    h->last_scan_total_ms = h->last_scan_ms;
    if (h->scan_total_valid) LIRA_CUDA_OK(cudaEventElapsedTime(&h->last_scan_total_ms, h->ev[4], h->ev[5]));
    h->scan_total_valid = false;
    LIRA_CUDA_OK(cudaEventElapsedTime(&h->last_total_ms, h->ev[2], h->ev[3]));
    unsigned long long stats[2] = {0, 0};
    if (h->stats.p) LIRA_CUDA_OK(cudaMemcpy(stats, h->stats.p, 16, cudaMemcpyDeviceToHost));
    // bytes_alg = E_p (4 d + 4) + Q 4 d + Q k 8
    h->last_scan_bytes = (long long)stats[0] * (4ll * h->d + 4) + h->last_Q * (4ll * h->d + 8ll * h->last_k);
    h->last_scan_pairs = (long long)stats[1];
    return 0;
}

// upload a host [n, d] fp32 matrix into a device buffer with row stride ds (zero padded)
// Large pageable host buffers (a base of hundreds of MB): the driver's own staging of a pageable cudaMemcpy runs at a few
// GB/s; here the bytes go through three pinned 16 MiB buffers filled by four host threads while the previous chunk is on the
// wire, so the copy runs at about the PCIe rate. Synchronous (returns when the data is on the device).
static int upload_big(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    constexpr size_t CH = 16u << 20;
    constexpr int NB = 3, NT = 4;
    static std::mutex mu;
    static void* pin[NB] = {nullptr, nullptr, nullptr};
    std::lock_guard<std::mutex> lk(mu);
    for (int i = 0; i < NB; ++i)
        if (!pin[i]) LIRA_CUDA_OK(cudaHostAlloc(&pin[i], CH, cudaHostAllocDefault));
    cudaEvent_t ev[NB];
    for (auto& e : ev) LIRA_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = 0;
    size_t i = 0;
    for (size_t off = 0; off < bytes && !rc; off += CH, ++i) {
        const int b = (int)(i % NB);
        const size_t len = std::min(CH, bytes - off);
        if (i >= NB && cudaEventSynchronize(ev[b]) != cudaSuccess) { set_error("upload: event wait failed"); rc = 2; break; }
        const size_t part = (len / NT + 63) & ~(size_t)63;
        std::thread th[NT - 1];
        for (int t = 1; t < NT; ++t) {
            const size_t a = std::min(len, part * t), e = std::min(len, part * (t + 1));
            th[t - 1] = std::thread([=]() { if (e > a) memcpy((char*)pin[b] + a, (const char*)src + off + a, e - a); });
        }
        memcpy(pin[b], (const char*)src + off, std::min(len, part));
        for (auto& t : th) t.join();
        if (cudaMemcpyAsync((char*)dst + off, pin[b], len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaEventRecord(ev[b], st) != cudaSuccess) { set_error("upload: cudaMemcpyAsync failed"); rc = 2; }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && !rc) { set_error("upload: stream synchronisation failed"); rc = 2; }
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

// contiguous host -> device copy on `st`: staged through upload_big when the source is large and pageable, asynchronous otherwise
static int upload_bytes(void* dst, const void* host, size_t bytes, cudaStream_t st) {
    if (bytes >= ((size_t)64 << 20)) {
        cudaPointerAttributes attr;
        const bool pageable = cudaPointerGetAttributes(&attr, host) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        if (pageable) return upload_big(dst, host, bytes, st);
    }
    LIRA_CUDA_OK(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

static int upload_rows(DevBuf& buf, const float* host, long long n, int d, int ds, cudaStream_t st) {
    if (int rc = buf.ensure((size_t)std::max<long long>(n, 1) * ds * 4)) return rc;
    if (n == 0) return 0;
    if (d == ds) {
        if (int rc = upload_bytes(buf.p, host, (size_t)n * d * 4, st)) return rc;
    } else {
        LIRA_CUDA_OK(cudaMemsetAsync(buf.p, 0, (size_t)n * ds * 4, st));
        LIRA_CUDA_OK(cudaMemcpy2DAsync(buf.p, (size_t)ds * 4, host, (size_t)d * 4, (size_t)d * 4, (size_t)n,
                                       cudaMemcpyHostToDevice, st));
    }
    return 0;
}

// the fp16 shadow copy of the list rows + their augmented-K blocks (tc_scan_kernels.cuh); built at create time, or on the
// first batch that needs it when the index also has the byte copy
static int ensure_fp16_shadow(lira_index* h) {
    if (h->has16 || !h->tc_ok) return 0;
    LIRA_CUDA_OK(cudaMalloc(&h->vaug, (size_t)std::max<long long>(h->E, 1) * 32));
    LIRA_CUDA_OK(cudaMalloc(&h->aaug, 128 * 32));
    LIRA_CUDA_OK(cudaMalloc(&h->vecs16, (size_t)std::max<long long>(h->E, 1) * h->d16 * 2));
    std::vector<__half> a(128 * 16, __float2half(0.0f));
    for (int r = 0; r < 128; ++r) {
        a[r * 16] = __float2half(h->tc_mode == 1 ? -2048.0f : -1.0f);
        a[r * 16 + 1] = __float2half(-1.0f);
    }
    LIRA_CUDA_OK(cudaMemcpy(h->aaug, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    shadow_rows_kernel<<<grid_for(h->E, 128, 148 * 16), 128, 0, h->stream>>>(h->vecs, h->ds, h->ds, h->E, h->vnorm, h->tc_sigma,
                                                                            h->tc_mode == 1, h->vecs16, h->d16, h->vaug);
    g_launches.fetch_add(1);
    LIRA_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (int rc = make_tmap_aug(&h->tmap_vaug, h->vaug, h->E)) return rc;
    if (int rc = make_tmap_aug(&h->tmap_aaug, h->aaug, 128)) return rc;
    if (int rc = make_tmap_f16(&h->tmap16, h->vecs16, h->E, h->d16, h->d16)) return rc;
    h->has16 = true;
    return 0;
}

// the byte shadow copy of the list rows + their |v|^2 as int32 (u8_scan_kernels.cuh)
static int ensure_u8_shadow(lira_index* h) {
    if (h->has8 || !h->u8_ok) return 0;
    h->u8_max_nseg = 1;
    h->u8_nseg_total = 0;
    for (int b = 0; b < h->B; ++b) {
        const long long nseg = std::max<long long>(1, (h->h_offsets[b + 1] - h->h_offsets[b] + U8_SEG_ROWS - 1) / U8_SEG_ROWS);
        h->u8_max_nseg = (int)std::max<long long>(h->u8_max_nseg, nseg);
        h->u8_nseg_total += nseg;
    }
    const size_t rows = (size_t)h->E + U8_NS;   // (a box may reach past the last entry)
    LIRA_CUDA_OK(cudaMalloc(&h->vecs8, rows * h->d8));
    LIRA_CUDA_OK(cudaMalloc(&h->nv_i, rows * 4));
    LIRA_CUDA_OK(cudaMalloc(&h->vaug8, rows * U8_AUG));
    LIRA_CUDA_OK(cudaMalloc(&h->aaug8, U8_AUG_BOX));
    LIRA_CUDA_OK(cudaMemsetAsync(h->vecs8 + (size_t)h->E * h->d8, 0, (size_t)U8_NS * h->d8, h->stream));
    LIRA_CUDA_OK(cudaMemsetAsync(h->nv_i + h->E, 0, (size_t)U8_NS * 4, h->stream));
    LIRA_CUDA_OK(cudaMemsetAsync(h->vaug8 + (size_t)h->E * U8_AUG, 0, (size_t)U8_NS * U8_AUG, h->stream));
    {
        std::vector<int8_t> a((size_t)U8_AUG_BOX);
        for (int r = 0; r < 128; ++r)
            for (int j = 0; j < U8_AUG; ++j) a[(size_t)r * U8_AUG + j] = j == U8_AUG - 1 ? (int8_t)-1 : (int8_t)-128;
        LIRA_CUDA_OK(cudaMemcpy(h->aaug8, a.data(), a.size(), cudaMemcpyHostToDevice));
    }
    if (h->E > 0) {
        shadow_rows_u8_kernel<<<grid_for(h->E * (h->d8 / 4), 256, 148 * 16), 256, 0, h->stream>>>(
            h->vecs, h->ds, h->ds, h->E, h->vnorm, h->metric == LIRA_METRIC_IP, h->vecs8, h->d8, h->nv_i);
        g_launches.fetch_add(1);
        if (h->metric != LIRA_METRIC_IP) {
            shadow_digits_u8_kernel<<<grid_for(h->E * 16, 256, 148 * 16), 256, 0, h->stream>>>(h->vnorm, h->E, h->vaug8);
            g_launches.fetch_add(1);
        } else {
            LIRA_CUDA_OK(cudaMemsetAsync(h->vaug8, 0, (size_t)h->E * U8_AUG, h->stream));
        }
    }
    LIRA_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (int rc = make_tmap_u8(&h->tmap8, h->vecs8, (long long)rows, h->d8, h->d8)) return rc;
    if (int rc = make_tmap_u8_aug(&h->tmap_vaug8, h->vaug8, (long long)rows, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return rc;
    if (int rc = make_tmap_u8_aug(&h->tmap_aaug8, h->aaug8, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return rc;
    h->has8 = true;
    return 0;
}

static int index_finish_create(lira_index* h, const long long* offsets) {
    h->h_offsets.assign(offsets, offsets + h->B + 1);
    LIRA_CUDA_OK(cudaMalloc(&h->d_offsets, (size_t)(h->B + 1) * 8));
    LIRA_CUDA_OK(cudaMemcpy(h->d_offsets, offsets, (size_t)(h->B + 1) * 8, cudaMemcpyHostToDevice));
    std::vector<int> order(h->B);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (offsets[a + 1] - offsets[a]) > (offsets[b + 1] - offsets[b]);
    });
    LIRA_CUDA_OK(cudaMalloc(&h->d_list_order, (size_t)std::max(h->B, 1) * 4));
    LIRA_CUDA_OK(cudaMemcpy(h->d_list_order, order.data(), (size_t)h->B * 4, cudaMemcpyHostToDevice));
    if (int rc = make_tmap(&h->tmap, h->vecs, h->E, h->ds, h->ds)) return rc;
    for (auto& e : h->ev) LIRA_CUDA_OK(cudaEventCreate(&e));
    // |v|^2 per entry, the exactness flag and the largest norm: they decide how the tensor-core scan may be used
    //   mode 1 (exact):       every value is an integer of <= 11 bits -> fp16 shadow copy is exact, results bit-identical
    //   mode 2 (approximate): any other finite data with d <= 256 -> fp16(sigma v) shadow copy, rigorous error margin in the
    //                         filter, every surviving candidate scored again exactly from the fp32 rows
    LIRA_CUDA_OK(cudaMalloc(&h->vnorm, (size_t)std::max<long long>(h->E, 1) * 4));
    int* d_flag = nullptr;
    LIRA_CUDA_OK(cudaMalloc(&d_flag, 8));
    int init[2] = {1, 0};
    LIRA_CUDA_OK(cudaMemcpy(d_flag, init, 8, cudaMemcpyHostToDevice));
    h->d16 = (h->ds + 7) / 8 * 8;
    const bool want16 = h->d16 <= TC_MAX_KB_STREAM * TC_KH && h->E > 0;
    if (h->E > 0) {
        row_norms_kernel<<<grid_for(h->E, 128, 148 * 16), 128, 0, h->stream>>>(h->vecs, h->ds, h->ds, h->E, h->vnorm, d_flag,
                                                                              (uint32_t*)(d_flag + 1));
        g_launches.fetch_add(1);
    }
    LIRA_CUDA_OK(cudaStreamSynchronize(h->stream));
    LIRA_CUDA_OK(cudaMemcpy(init, d_flag, 8, cudaMemcpyDeviceToHost));
    cudaFree(d_flag);
    float max_norm2;
    memcpy(&max_norm2, &init[1], 4);
    h->tc_mode = 0;
    h->tc_sigma = 1.0f;
    if (want16 && init[0] == 1) h->tc_mode = 1;
    else if (want16 && max_norm2 > 0.f && max_norm2 < 1e30f && !getenv("LIRA_NO_APPROX_TC")) {
        // sigma = 2^e with sigma * max|v| in (48, 96]: fp16 never overflows (|sigma v_i| <= 96, sigma^2 |v|^2 <= 9216) and
        // ordinary components stay far above the fp16 subnormal range
        const float vmax = std::sqrt(max_norm2);
        int e = (int)std::floor(std::log2(96.0f / vmax));
        e = std::max(-60, std::min(60, e));
        h->tc_sigma = std::ldexp(1.0f, e);
        h->tc_vmax = vmax;
        h->tc_mode = 2;
    }
    h->tc_ok = h->tc_mode != 0;
    // byte-valued data: the u8 shadow copy (and no fp16 copy until a batch needs one)
    h->d8 = (h->d + 15) / 16 * 16;
    if (h->tc_mode == 1 && h->d8 <= U8_MAX_D && max_norm2 <= (float)U8_MAX_NORM && !getenv("LIRA_NO_U8")) {
        int* d_u8 = nullptr;
        LIRA_CUDA_OK(cudaMalloc(&d_u8, 4));
        int one = 1;
        LIRA_CUDA_OK(cudaMemcpy(d_u8, &one, 4, cudaMemcpyHostToDevice));
        check_u8_kernel<<<grid_for(h->E * ((h->d + 3) / 4), 256, 148 * 16), 256, 0, h->stream>>>(h->vecs, h->ds, h->d, h->E, d_u8);
        g_launches.fetch_add(1);
        LIRA_CUDA_OK(cudaStreamSynchronize(h->stream));
        LIRA_CUDA_OK(cudaMemcpy(&one, d_u8, 4, cudaMemcpyDeviceToHost));
        cudaFree(d_u8);
        h->u8_ok = one == 1;
    }
    // which shadow copy is built now: the byte copy for exact-kNN handles (their exhaustive probe sets run the byte scan),
    // the fp16 copy for everything else; the other one is built by the first batch that needs it
    if (h->u8_ok && h->prefer_u8) {
        if (int rc = ensure_u8_shadow(h)) return rc;
    } else if (h->tc_ok) {
        if (int rc = ensure_fp16_shadow(h)) return rc;
    }
    cudaDeviceProp prop;
    LIRA_CUDA_OK(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    return 0;
}

static int validate_offsets(const int64_t* off, int B, long long* E) {
    LIRA_REQUIRE(off[0] == 0, "list_offsets[0] must be 0");
    for (int b = 0; b < B; ++b) LIRA_REQUIRE(off[b + 1] >= off[b], "list_offsets must be non-decreasing");
    *E = off[B];
    LIRA_REQUIRE(*E < (1ll << 31), "more than 2^31 list entries per index: shard the lists across GPUs");
    return 0;
}

}  // namespace lira

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* lira_last_error(void) { return g_err.c_str(); }
int lira_version(void) { return 100; }
int lira_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int64_t lira_launch_count(void) { return g_launches.load(); }

int lira_index_create(const float* base, int64_t N, int d, const int64_t* list_offsets, const int32_t* list_ids,
                      int B, int metric, int device, lira_index_t** out) {
    LIRA_REQUIRE(out && base && list_offsets && (list_ids || list_offsets[B] == 0), "null argument");
    LIRA_REQUIRE(N >= 0 && d >= 1 && B >= 1, "bad shape");
    LIRA_REQUIRE(metric == LIRA_METRIC_L2 || metric == LIRA_METRIC_IP, "metric must be LIRA_METRIC_L2 or LIRA_METRIC_IP");
    if (int rc = check_device(device)) return rc;
    long long E = 0;
    if (int rc = validate_offsets(list_offsets, B, &E)) return rc;
    for (long long e = 0; e < E; ++e)
        LIRA_REQUIRE(list_ids[e] >= 0 && list_ids[e] < N, "list id out of range");
    lira_index* h = new lira_index();
    h->device = device; h->B = B; h->d = d; h->ds = round_up(d, 4); h->metric = metric; h->E = E; h->owns = true;
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); rc = 2; break; }
        // base goes up in row chunks and is gathered into list order on the device
        if (cudaMalloc(&h->vecs, (size_t)std::max<long long>(E, 1) * h->ds * 4) != cudaSuccess) { set_error("cudaMalloc(list vectors) failed"); rc = 2; break; }
        if (cudaMalloc(&h->ids, (size_t)std::max<long long>(E, 1) * 4) != cudaSuccess) { set_error("cudaMalloc(list ids) failed"); rc = 2; break; }
        float* d_base = nullptr;
        if (cudaMalloc(&d_base, (size_t)std::max<long long>(N, 1) * d * 4) != cudaSuccess) { set_error("cudaMalloc(base) failed"); rc = 2; break; }
        if (upload_bytes(d_base, base, (size_t)N * d * 4, h->stream)) { cudaFree(d_base); rc = 2; break; }
        cudaMemcpyAsync(h->ids, list_ids, (size_t)E * 4, cudaMemcpyHostToDevice, h->stream);
        if (E > 0) {
            gather_rows_kernel<<<grid_for(E * (h->ds / 4), 256, 148 * 16), 256, 0, h->stream>>>(d_base, d, d, h->ids, E, h->vecs, h->ds);
            g_launches.fetch_add(1);
        }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        cudaFree(d_base);
        if (e != cudaSuccess) { set_error(std::string("index build failed: ") + cudaGetErrorString(e)); rc = 2; break; }
        std::vector<long long> off(list_offsets, list_offsets + B + 1);
        rc = index_finish_create(h, off.data());
    } while (0);
    if (rc) { lira_index_free(h); return rc; }
    *out = h;
    return 0;
}

int lira_index_create_from_assign(const float* base, int64_t N, int d, const int32_t* data_2_bkt, int n_mul, int B,
                                  int metric, int device, lira_index_t** out) {
    LIRA_REQUIRE(out && base && data_2_bkt && n_mul >= 1 && B >= 1, "bad argument");
    // search.cpp:368-386: ids enter each bucket in increasing order, so "sort + unique" reduces to
    // skipping an id equal to the bucket's last one.
    std::vector<int64_t> cnt(B + 1, 0);
    std::vector<int32_t> last(B, -1);
    for (int64_t i = 0; i < N; ++i)
        for (int j = 0; j < n_mul; ++j) {
            const int b = data_2_bkt[i * n_mul + j];
            if (b < 0) continue;
            LIRA_REQUIRE(b < B, "bucket id out of range.");
            if (last[b] == (int32_t)i) continue;
            last[b] = (int32_t)i;
            cnt[b + 1]++;
        }
    for (int b = 0; b < B; ++b) cnt[b + 1] += cnt[b];
    std::vector<int32_t> ids((size_t)std::max<int64_t>(cnt[B], 1));
    std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
    std::fill(last.begin(), last.end(), -1);
    for (int64_t i = 0; i < N; ++i)
        for (int j = 0; j < n_mul; ++j) {
            const int b = data_2_bkt[i * n_mul + j];
            if (b < 0 || last[b] == (int32_t)i) continue;
            last[b] = (int32_t)i;
            ids[cur[b]++] = (int32_t)i;
        }
    return lira_index_create(base, N, d, cnt.data(), ids.data(), B, metric, device, out);
}

static int index_create_dev_impl(const float* d_vecs, int64_t ld, int d, const int64_t* list_offsets, const int32_t* d_ids,
                                 int B, int metric, int device, lira_index_t** out, bool prefer_u8);
int lira_index_create_dev(const float* d_vecs, int64_t ld, int d, const int64_t* list_offsets, const int32_t* d_ids,
                          int B, int metric, int device, lira_index_t** out) {
    return index_create_dev_impl(d_vecs, ld, d, list_offsets, d_ids, B, metric, device, out, false);
}
static int index_create_dev_impl(const float* d_vecs, int64_t ld, int d, const int64_t* list_offsets, const int32_t* d_ids,
                                 int B, int metric, int device, lira_index_t** out, bool prefer_u8) {
    LIRA_REQUIRE(out && d_vecs && list_offsets && d_ids && B >= 1 && d >= 1, "bad argument");
    LIRA_REQUIRE(ld >= d && (ld % 4) == 0 && ((uintptr_t)d_vecs & 15) == 0, "device vectors need ld % 4 == 0 and 16-byte alignment");
    LIRA_REQUIRE(metric == LIRA_METRIC_L2 || metric == LIRA_METRIC_IP, "metric must be LIRA_METRIC_L2 or LIRA_METRIC_IP");
    if (int rc = check_device(device)) return rc;
    long long E = 0;
    if (int rc = validate_offsets(list_offsets, B, &E)) return rc;
    lira_index* h = new lira_index();
    h->device = device; h->B = B; h->d = d; h->ds = (int)ld; h->metric = metric; h->E = E; h->owns = false;
    h->prefer_u8 = prefer_u8;
    h->vecs = const_cast<float*>(d_vecs);
    h->ids = const_cast<int*>(d_ids);
    int rc = 0;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); rc = 2; }
    if (!rc) {
        std::vector<long long> off(list_offsets, list_offsets + B + 1);
        rc = index_finish_create(h, off.data());
    }
    if (rc) { lira_index_free(h); return rc; }
    *out = h;
    return 0;
}

int lira_index_free(lira_index_t* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->owns) { cudaFree(h->vecs); cudaFree(h->ids); }
    cudaFree(h->d_offsets);
    cudaFree(h->d_list_order);
    cudaFree(h->vnorm);
    cudaFree(h->vaug);
    cudaFree(h->vecs16);
    cudaFree(h->vecs8);
    cudaFree(h->nv_i);
    cudaFree(h->vaug8);
    cudaFree(h->aaug8);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    for (PendingBatch& pb : h->pending) { if (pb.h_flags) cudaFreeHost(pb.h_flags); if (pb.ev) cudaEventDestroy(pb.ev); }
    for (int* p : h->flag_pool) cudaFreeHost(p);
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    for (auto& sl : h->slots) {
        for (DevBuf* b : {&sl.q, &sl.D, &sl.I, &sl.nprobe, &sl.cmp}) b->release();
        if (sl.pin_in) cudaFreeHost(sl.pin_in);
        if (sl.pin_out) cudaFreeHost(sl.pin_out);
        if (sl.pb.enqueued) { cudaFreeHost(sl.pb.h_flags); cudaEventDestroy(sl.pb.ev); }
        for (cudaEvent_t e : {sl.ev_in, sl.ev_done, sl.ev_out, sl.ev_t[0], sl.ev_t[1], sl.ev_t[2], sl.ev_t[3]}) if (e) cudaEventDestroy(e);
    }
    if (h->st_in) cudaStreamDestroy(h->st_in);
    if (h->st_out) cudaStreamDestroy(h->st_out);
    cudaFree(h->aaug);
    h->ws.release();
    h->ws_seed.release();
    h->stats.release();
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int64_t lira_index_ntotal(const lira_index_t* h, int list) {
    if (!h) return -1;
    if (list < 0) return h->E;
    if (list >= h->B) return -1;
    return h->h_offsets[list + 1] - h->h_offsets[list];
}
int lira_index_nlist(const lira_index_t* h) { return h ? h->B : -1; }
int lira_index_dim(const lira_index_t* h) { return h ? h->d : -1; }

int lira_index_set_use_tensor_cores(lira_index_t* h, int enable) {
    LIRA_REQUIRE(h, "null index");
    h->use_tc = enable != 0;
    return 0;
}
int lira_index_last_path(const lira_index_t* h) { return h ? h->last_path : -1; }
int lira_index_last_redo(const lira_index_t* h) { return h ? h->last_redo : -1; }
int lira_index_tensor_core_eligible(const lira_index_t* h) { return h ? (h->tc_ok ? 1 : 0) : -1; }
int lira_index_tensor_core_mode(const lira_index_t* h) { return h ? h->tc_mode : -1; }
int lira_index_byte_scan_eligible(const lira_index_t* h) { return h ? (h->u8_ok ? 1 : 0) : -1; }
int lira_index_last_scan_kind(const lira_index_t* h) { return h ? (h->last_path == 0 ? 0 : (h->last_u8 ? 2 : 1)) : -1; }

int lira_index_set_timing(lira_index_t* h, int enable) {
    LIRA_REQUIRE(h, "null index");
    h->timing = enable != 0;
    return 0;
}
int lira_index_last_timing(const lira_index_t* hc, float* scan_ms, float* total_ms, int64_t* scan_bytes, int64_t* scan_pairs) {
    LIRA_REQUIRE(hc, "null index");
    lira_index_t* h = const_cast<lira_index_t*>(hc);
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    if (int rc = finish_timing(h)) return rc;
    if (scan_ms) *scan_ms = h->last_scan_ms;
    if (total_ms) *total_ms = h->last_total_ms;
    if (scan_bytes) *scan_bytes = h->last_scan_bytes;
    if (scan_pairs) *scan_pairs = h->last_scan_pairs;
    return 0;
}

int lira_index_last_scan_total_ms(const lira_index_t* hc, float* scan_total_ms) {
    LIRA_REQUIRE(hc && scan_total_ms, "null argument");
    lira_index_t* h = const_cast<lira_index_t*>(hc);
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    if (int rc = finish_timing(h)) return rc;
    *scan_total_ms = h->last_scan_total_ms;
    return 0;
}

int lira_index_list_search(lira_index_t* h, int list, const float* q, int64_t nq, int k, float* D, int64_t* I) {
    LIRA_REQUIRE(h && q && D && I, "null argument");
    LIRA_REQUIRE(list >= 0 && list < h->B, "list out of range");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Workspace& ws = h->ws;
    if (int rc = upload_rows(ws.q, q, nq, h->d, h->ds, st)) return rc;
    // probe sets: every query probes exactly `list`
    std::vector<long long> po(nq + 1);
    std::iota(po.begin(), po.end(), 0ll);
    std::vector<int> pi((size_t)std::max<int64_t>(nq, 1), list);
    if (int rc = ws.probe_ids.ensure((size_t)(nq + 1) * 4 + (size_t)(nq + 1) * 8 + 64)) return rc;
    long long* d_po = (long long*)ws.probe_ids.p;
    int* d_pi = (int*)(d_po + nq + 1);
    LIRA_CUDA_OK(cudaMemcpyAsync(d_po, po.data(), (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, st));
    LIRA_CUDA_OK(cudaMemcpyAsync(d_pi, pi.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    ProbeSpec ps;
    ps.kind = 1; ps.d_probe_offsets = d_po; ps.d_probe_ids = d_pi; ps.P = nq;
    long long P = 0;
    if (int rc = run_grouped_scan(h, ws.q.as<float>(), h->ds, nq, ps, k, /*store_local=*/1, nullptr, &P, nullptr, st)) return rc;
    if (nq == 0) return 0;
    if (int rc = ws.D.ensure((size_t)nq * k * 4)) return rc;
    if (int rc = ws.I.ensure((size_t)nq * k * 8)) return rc;
    MergeParams mp{ws.part_key.as<unsigned long long>(), d_po, ws.probe_slot.as<int>(), k, (int)nq, 0,
                   ws.D.as<float>(), ws.I.as<long long>(), h->metric == LIRA_METRIC_IP, nullptr};
    if (int rc = launch_merge(mp, st)) return rc;
    LIRA_CUDA_OK(cudaMemcpyAsync(D, ws.D.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    LIRA_CUDA_OK(cudaMemcpyAsync(I, ws.I.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    LIRA_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int lira_scan_all_pairs(lira_index_t* h, const float* q, int64_t Q, int k, int64_t* found, int64_t* cmp) {
    LIRA_REQUIRE(h && q && found, "null argument");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Workspace& ws = h->ws;
    const int B = h->B;
    // bound the partial-result workspace: process the queries in batches
    // at most 64 Mi (query, list) pairs per batch, and at most ~6 GB for the two P * k * 8-byte workspaces (partial keys + found ids)
    const long long max_pairs = std::min<long long>(64ll << 20, (6ll << 30) / (16ll * k));
    long long qb = std::max<long long>(1, std::min<long long>(Q, max_pairs / std::max(B, 1)));
    for (long long q0 = 0; q0 < Q; q0 += qb) {
        const long long nq = std::min<long long>(qb, Q - q0);
        if (int rc = upload_rows(ws.q, q + q0 * h->d, nq, h->d, h->ds, st)) return rc;
        ProbeSpec ps;
        ps.kind = 2;
        long long P = 0;
        if (int rc = run_grouped_scan(h, ws.q.as<float>(), h->ds, nq, ps, k, 0, nullptr, &P, nullptr, st)) return rc;
        if (int rc = ws.I.ensure((size_t)P * k * 8)) return rc;
        if (int rc = ws.cmp.ensure((size_t)P * 8)) return rc;
        found_from_partials_kernel<<<grid_for(P * k, 256, 148 * 16), 256, 0, st>>>(
            ws.part_key.as<unsigned long long>(), ws.probe_slot.as<int>(), h->d_offsets, h->ids, (int)nq, B, k,
            ws.I.as<long long>(), ws.cmp.as<long long>());
        LIRA_LAUNCH_CHECK();
        LIRA_CUDA_OK(cudaMemcpyAsync(found + (size_t)q0 * B * k, ws.I.p, (size_t)P * k * 8, cudaMemcpyDeviceToHost, st));
        if (cmp) LIRA_CUDA_OK(cudaMemcpyAsync(cmp + (size_t)q0 * B, ws.cmp.p, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}

int lira_search_dev(lira_index_t* h, const float* d_q, int64_t ldq, int64_t Q, const int64_t* d_probe_offsets,
                    const int32_t* d_probe_ids, int64_t P, int k, int dedup, float* d_D, int64_t* d_I, int64_t* d_cmp,
                    void* stream) {
    LIRA_REQUIRE(h && d_q && d_probe_offsets && (d_probe_ids || P == 0) && d_D && d_I, "null argument");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    ProbeSpec ps;
    ps.kind = 1; ps.d_probe_offsets = (const long long*)d_probe_offsets; ps.d_probe_ids = d_probe_ids; ps.P = P;
    return search_core(h, d_q, ldq, Q, ps, k, dedup, d_D, (long long*)d_I, nullptr, (long long*)d_cmp, st);
}

int lira_search(lira_index_t* h, const float* q, int64_t Q, const int64_t* probe_offsets, const int32_t* probe_ids,
                int k, int dedup, float* D, int64_t* I, int64_t* cmp) {
    LIRA_REQUIRE(h && q && probe_offsets && D && I, "null argument");
    LIRA_REQUIRE(probe_offsets[0] == 0, "probe_offsets[0] must be 0");
    const int64_t P = probe_offsets[Q];
    for (int64_t j = 0; j < P; ++j) LIRA_REQUIRE(probe_ids[j] >= 0 && probe_ids[j] < h->B, "probed list id out of range");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Workspace& ws = h->ws;
    if (int rc = upload_rows(ws.q, q, Q, h->d, h->ds, st)) return rc;
    if (int rc = ws.probe_ids.ensure((size_t)(Q + 1) * 8 + (size_t)(P + 1) * 4 + 64)) return rc;
    long long* d_po = (long long*)ws.probe_ids.p;
    int* d_pi = (int*)(d_po + Q + 1);
    LIRA_CUDA_OK(cudaMemcpyAsync(d_po, probe_offsets, (size_t)(Q + 1) * 8, cudaMemcpyHostToDevice, st));
    if (P) LIRA_CUDA_OK(cudaMemcpyAsync(d_pi, probe_ids, (size_t)P * 4, cudaMemcpyHostToDevice, st));
    if (int rc = ws.D.ensure((size_t)std::max<int64_t>(Q, 1) * k * 4)) return rc;
    if (int rc = ws.I.ensure((size_t)std::max<int64_t>(Q, 1) * k * 8)) return rc;
    if (int rc = ws.cmp.ensure((size_t)std::max<int64_t>(Q, 1) * 8)) return rc;
    if (int rc = lira_search_dev(h, ws.q.as<float>(), h->ds, Q, (const int64_t*)d_po, d_pi, P, k, dedup, ws.D.as<float>(),
                                 (int64_t*)ws.I.p, (int64_t*)ws.cmp.p, st)) return rc;
    if (Q) {
        LIRA_CUDA_OK(cudaMemcpyAsync(D, ws.D.p, (size_t)Q * k * 4, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaMemcpyAsync(I, ws.I.p, (size_t)Q * k * 8, cudaMemcpyDeviceToHost, st));
        if (cmp) LIRA_CUDA_OK(cudaMemcpyAsync(cmp, ws.cmp.p, (size_t)Q * 8, cudaMemcpyDeviceToHost, st));
    }
    LIRA_CUDA_OK(cudaStreamSynchronize(st));
    return finish_timing(h);
}

// ---- probing model ---------------------------------------------------------------------------
int lira_model_create(const float* centroids, const float* scaler_mean, const float* scaler_scale, int B, int d,
                      const float* const weights[12], int device, lira_model_t** out) {
    LIRA_REQUIRE(out && centroids && weights && B >= 1 && d >= 1, "bad argument");
    for (int i = 0; i < 12; ++i) LIRA_REQUIRE(weights[i], "null weight tensor");
    if (int rc = check_device(device)) return rc;
    lira_model* m = new lira_model();
    m->device = device; m->B = B; m->Bp = round_up(B, 4); m->d = d; m->ds = round_up(d, 4);
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->num_sms = prop.multiProcessorCount;
    }
    const int od[6] = {128, 64, 128, 64, 128, B};
    const int id[6] = {B, 128, d, 128, 128, 128};
    int rc = 0;
    auto up = [&](float** dst, const float* src, int rows, int cols, int ld) -> int {
        LIRA_CUDA_OK(cudaMalloc(dst, (size_t)rows * ld * 4));
        LIRA_CUDA_OK(cudaMemset(*dst, 0, (size_t)rows * ld * 4));
        LIRA_CUDA_OK(cudaMemcpy2D(*dst, (size_t)ld * 4, src, (size_t)cols * 4, (size_t)cols * 4, rows, cudaMemcpyHostToDevice));
        return 0;
    };
    do {
        if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); rc = 2; break; }
        if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) { set_error("cudaStreamCreate failed"); rc = 2; break; }
        if ((rc = up(&m->centroids, centroids, B, d, m->ds))) break;
        if ((rc = make_tmap(&m->tm_cent, m->centroids, B, m->ds, m->ds))) break;
        if (scaler_mean && scaler_scale) {
            if ((rc = up(&m->mean, scaler_mean, 1, B, m->Bp))) break;
            if ((rc = up(&m->scale, scaler_scale, 1, B, m->Bp))) break;
        }
        for (int l = 0; l < 6 && !rc; ++l) {
            m->out_dim[l] = od[l];
            m->in_dim[l] = id[l];
            m->in_ld[l] = round_up(id[l], 4);
            if ((rc = up(&m->W[l], weights[2 * l], od[l], id[l], m->in_ld[l]))) break;
            if ((rc = up(&m->bias[l], weights[2 * l + 1], 1, od[l], round_up(od[l], 4)))) break;
            rc = make_tmap(&m->tm_w[l], m->W[l], od[l], m->in_ld[l], m->in_ld[l]);
        }
        if (rc) break;
        // tensor-core front end: error-free hi / lo splits of the static operands (tc_dense_kernels.cuh)
        auto split_up = [&](float** dh, float** dl, const std::vector<float>& x, int rows, int cols, int ld) -> int {
            std::vector<float> hi(x.size()), lo(x.size());
            for (size_t i = 0; i < x.size(); ++i) {
                uint32_t b;
                memcpy(&b, &x[i], 4);
                b &= 0xFFFFE000u;
                memcpy(&hi[i], &b, 4);
                lo[i] = x[i] - hi[i];
            }
            if (int r = up(dh, hi.data(), rows, cols, ld)) return r;
            return up(dl, lo.data(), rows, cols, ld);
        };
        {
            // centre by the centroid mean (double accumulation), |c'|^2 in double
            std::vector<double> mu(d, 0.0);
            for (int b = 0; b < B; ++b)
                for (int j = 0; j < d; ++j) mu[j] += centroids[(size_t)b * d + j];
            std::vector<float> muf(d), cc((size_t)B * d), cn(B);
            for (int j = 0; j < d; ++j) muf[j] = (float)(mu[j] / B);
            for (int b = 0; b < B; ++b) {
                double s = 0.0;
                for (int j = 0; j < d; ++j) {
                    const float v = centroids[(size_t)b * d + j] - muf[j];
                    cc[(size_t)b * d + j] = v;
                    s += (double)v * v;
                }
                cn[b] = (float)s;
            }
            if ((rc = up(&m->mu, muf.data(), 1, d, m->ds))) break;
            if ((rc = up(&m->cn, cn.data(), 1, B, m->Bp))) break;
            if ((rc = split_up(&m->cent_h, &m->cent_l, cc, B, d, m->ds))) break;
            if ((rc = make_tmap(&m->tm_ch, m->cent_h, B, d, m->ds))) break;
            if ((rc = make_tmap(&m->tm_cl, m->cent_l, B, d, m->ds))) break;
        }
        for (int l = 0; l < 6 && !rc; ++l) {
            std::vector<float> w(weights[2 * l], weights[2 * l] + (size_t)od[l] * id[l]);
            if ((rc = split_up(&m->W_h[l], &m->W_l[l], w, od[l], id[l], m->in_ld[l]))) break;
            if ((rc = make_tmap(&m->tm_wh[l], m->W_h[l], od[l], id[l], m->in_ld[l]))) break;
            rc = make_tmap(&m->tm_wl[l], m->W_l[l], od[l], id[l], m->in_ld[l]);
        }
    } while (0);
    if (rc) { lira_model_free(m); return rc; }
    *out = m;
    return 0;
}

int lira_model_set_use_tensor_cores(lira_model_t* m, int enable) {
    LIRA_REQUIRE(m, "null model");
    m->use_tc = enable != 0;
    return 0;
}

int lira_model_free(lira_model_t* m) {
    if (!m) return 0;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->side) { cudaStreamSynchronize(m->side); cudaStreamDestroy(m->side); }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    cudaFree(m->centroids); cudaFree(m->mean); cudaFree(m->scale);
    for (int l = 0; l < 6; ++l) { cudaFree(m->W[l]); cudaFree(m->bias[l]); cudaFree(m->W_h[l]); cudaFree(m->W_l[l]); }
    cudaFree(m->mu); cudaFree(m->cent_h); cudaFree(m->cent_l); cudaFree(m->cn);
    for (DevBuf* b : {&m->feats, &m->h1, &m->cat, &m->h2, &m->h5, &m->scores, &m->q, &m->qch, &m->qcl, &m->qn, &m->qrh, &m->qrl,
                      &m->fh, &m->fl, &m->h1h, &m->h1l, &m->cath, &m->catl, &m->h2h, &m->h2l, &m->h5h, &m->h5l}) b->release();
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return 0;
}

int lira_centroid_features(const float* q, int64_t Q, const float* centroids, int B, int d, const float* mean,
                           const float* scale, int device, float* out) {
    LIRA_REQUIRE(q && centroids && out && B >= 1 && d >= 1 && Q >= 0, "bad argument");
    LIRA_REQUIRE((mean == nullptr) == (scale == nullptr), "mean and scale must be given together");
    if (int rc = check_device(device)) return rc;
    const int ds = round_up(d, 4), Bp = round_up(B, 4);
    DevBuf cent, mq, ms, dq, dout;
    cudaStream_t st = nullptr;
    int rc = 0;
    auto body = [&]() -> int {
        LIRA_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (int r = upload_rows(cent, centroids, B, d, ds, st)) return r;
        if (mean) {
            if (int r = upload_rows(mq, mean, 1, B, Bp, st)) return r;
            if (int r = upload_rows(ms, scale, 1, B, Bp, st)) return r;
        }
        CUtensorMap tm;
        if (int r = make_tmap(&tm, cent.as<float>(), B, ds, ds)) return r;
        const long long chunk = 65536;
        for (long long q0 = 0; q0 < Q; q0 += chunk) {
            const long long nq = std::min<long long>(chunk, Q - q0);
            if (int r = upload_rows(dq, q + q0 * d, nq, d, ds, st)) return r;
            if (int r = dout.ensure((size_t)nq * Bp * 4)) return r;
            DenseParams p{dq.as<float>(), ds, (int)nq, ds, B, dout.as<float>(), Bp, 0, mean ? mq.as<float>() : nullptr,
                          mean ? ms.as<float>() : nullptr};
            if (int r = launch_dense<64, OP_L2, EPI_FEATURE>(tm, p, st)) return r;
            LIRA_CUDA_OK(cudaMemcpy2DAsync(out + (size_t)q0 * B, (size_t)B * 4, dout.p, (size_t)Bp * 4, (size_t)B * 4,
                                           (size_t)nq, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
        }
        return 0;
    };
    rc = body();
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (DevBuf* b : {&cent, &mq, &ms, &dq, &dout}) b->release();
    return rc;
}

int lira_model_scores(lira_model_t* m, const float* q, int64_t Q, float* scores, float* feats) {
    LIRA_REQUIRE(m && q && scores && Q >= 0, "bad argument");
    LIRA_CUDA_OK(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const long long chunk = 32768;
    for (long long q0 = 0; q0 < Q; q0 += chunk) {
        const long long nq = std::min<long long>(chunk, Q - q0);
        if (int rc = upload_rows(m->q, q + q0 * m->d, nq, m->d, m->ds, st)) return rc;
        if (int rc = m->scores.ensure((size_t)nq * m->Bp * 4)) return rc;
        if (int rc = m->feats.ensure((size_t)nq * m->Bp * 4)) return rc;
        if (int rc = model_forward(m, m->q.as<float>(), m->ds, nq, m->scores.as<float>(), m->Bp, m->feats.as<float>(), m->Bp, st)) return rc;
        LIRA_CUDA_OK(cudaMemcpy2DAsync(scores + (size_t)q0 * m->B, (size_t)m->B * 4, m->scores.p, (size_t)m->Bp * 4,
                                       (size_t)m->B * 4, (size_t)nq, cudaMemcpyDeviceToHost, st));
        if (feats)
            LIRA_CUDA_OK(cudaMemcpy2DAsync(feats + (size_t)q0 * m->B, (size_t)m->B * 4, m->feats.p, (size_t)m->Bp * 4,
                                           (size_t)m->B * 4, (size_t)nq, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}

int lira_select_search_dev(lira_index_t* h, const float* d_scores, int64_t lds, const float* d_q, int64_t ldq, int64_t Q,
                           int mode, double value, int k, int dedup, float* d_D, int64_t* d_I, int32_t* d_nprobe,
                           int64_t* d_cmp, void* stream) {
    LIRA_REQUIRE(h && d_scores && d_q && d_D && d_I, "null argument");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    ProbeSpec ps;
    ps.kind = 0; ps.d_scores = d_scores; ps.lds = lds; ps.mode = mode; ps.value = value;
    return search_core(h, d_q, ldq, Q, ps, k, dedup, d_D, (long long*)d_I, d_nprobe, (long long*)d_cmp, st);
}

int lira_probe_search_dev(lira_index_t* h, lira_model_t* m, const float* d_q, int64_t ldq, int64_t Q, int mode,
                          double value, int k, int dedup, float* d_D, int64_t* d_I, int32_t* d_nprobe, int64_t* d_cmp,
                          void* stream) {
    LIRA_REQUIRE(h && m && d_q && d_D && d_I, "null argument");
    LIRA_REQUIRE(h->device == m->device && h->B == m->B && h->d == m->d, "index and model disagree on device / B / d");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    // fused front end (selection inside the last layer, no scores in HBM) where the tensor-core scan serves the batch
    h->last_Q = Q;
    h->last_k = k;
    h->last_redo = 0;
    if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[2], st));
    bool done = false;
    int n_redo = 0;
    if (int rc = fused_probe_search(h, m, d_q, ldq, Q, mode, value, k, dedup, d_D, (long long*)d_I, d_nprobe, (long long*)d_cmp, st, &done, &n_redo)) return rc;
    if (done) {
        if (n_redo > 0) {   // queries whose candidate regions overflowed: exact CUDA-core scan of their probe sets
            ProbeSpec ps;
            ps.kind = 1; ps.d_probe_offsets = h->ws.probe_offsets.as<long long>(); ps.d_probe_ids = h->ws.probe_ids.as<int>();
            ps.P = Q * (long long)std::min(h->B, h->nprobe_cap);
            if (int rc = exact_fallback(h, d_q, ldq, Q, ps, k, dedup, d_D, (long long*)d_I, nullptr, nullptr, st, true)) return rc;
        }
        if (h->timing) {
            LIRA_CUDA_OK(cudaEventRecord(h->ev[3], st));
            h->timing_pending = true;
        }
        return 0;
    }
    if (int rc = h->ws.scores.ensure((size_t)std::max<int64_t>(Q, 1) * m->Bp * 4)) return rc;
    if (int rc = model_forward(m, d_q, ldq, Q, h->ws.scores.as<float>(), m->Bp, nullptr, 0, st)) return rc;
    return lira_select_search_dev(h, h->ws.scores.as<float>(), m->Bp, d_q, ldq, Q, mode, value, k, dedup, d_D, d_I,
                                  d_nprobe, d_cmp, st);
}

int lira_probe_search(lira_index_t* h, lira_model_t* m, const float* q, int64_t Q, int mode, double value, int k,
                      int dedup, float* D, int64_t* I, int32_t* nprobe, int64_t* cmp) {
    LIRA_REQUIRE(h && m && q && D && I, "null argument");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Workspace& ws = h->ws;
    const size_t Qs = (size_t)std::max<int64_t>(Q, 1);
    if (int rc = upload_rows(ws.q, q, Q, h->d, h->ds, st)) return rc;
    if (int rc = ws.D.ensure(Qs * k * 4)) return rc;
    if (int rc = ws.I.ensure(Qs * k * 8)) return rc;
    if (int rc = ws.cmp.ensure(Qs * 8)) return rc;
    if (int rc = ws.nprobe.ensure(Qs * 4)) return rc;
    if (int rc = lira_probe_search_dev(h, m, ws.q.as<float>(), h->ds, Q, mode, value, k, dedup, ws.D.as<float>(),
                                       (int64_t*)ws.I.p, ws.nprobe.as<int>(), (int64_t*)ws.cmp.p, st)) return rc;
    if (Q) {
        // results: device -> pinned staging (asynchronous, full PCIe rate) -> the caller's buffers
        const size_t bD = (size_t)Q * k * 4, bI = (size_t)Q * k * 8, bC = cmp ? (size_t)Q * 8 : 0, bN = nprobe ? (size_t)Q * 4 : 0;
        const size_t oI = (bD + 255) & ~(size_t)255, oC = oI + ((bI + 255) & ~(size_t)255), oN = oC + ((bC + 255) & ~(size_t)255);
        const size_t total = oN + bN;
        if (total > h->h_stage_cap) {
            if (h->h_stage) cudaFreeHost(h->h_stage);
            h->h_stage = nullptr;
            h->h_stage_cap = 0;
            LIRA_CUDA_OK(cudaHostAlloc(&h->h_stage, total + total / 4, cudaHostAllocDefault));
            h->h_stage_cap = total + total / 4;
        }
        char* sg = (char*)h->h_stage;
        LIRA_CUDA_OK(cudaMemcpyAsync(sg, ws.D.p, bD, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaMemcpyAsync(sg + oI, ws.I.p, bI, cudaMemcpyDeviceToHost, st));
        if (cmp) LIRA_CUDA_OK(cudaMemcpyAsync(sg + oC, ws.cmp.p, bC, cudaMemcpyDeviceToHost, st));
        if (nprobe) LIRA_CUDA_OK(cudaMemcpyAsync(sg + oN, ws.nprobe.p, bN, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        memcpy(D, sg, bD);
        memcpy(I, sg + oI, bI);
        if (cmp) memcpy(cmp, sg + oC, bC);
        if (nprobe) memcpy(nprobe, sg + oN, bN);
    }
    LIRA_CUDA_OK(cudaStreamSynchronize(st));
    return finish_timing(h);
}

// ---- asynchronous forms: enqueue now, check later ------------------------------------------------------------------------
namespace lira {
static int take_flags_and_event(lira_index* h, int** fl, cudaEvent_t* ev) {
    if (h->flag_pool.empty()) {
        int* p = nullptr;
        LIRA_CUDA_OK(cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault));
        h->flag_pool.push_back(p);
    }
    if (h->event_pool.empty()) {
        cudaEvent_t e;
        LIRA_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->event_pool.push_back(e);
    }
    *fl = h->flag_pool.back(); h->flag_pool.pop_back();
    *ev = h->event_pool.back(); h->event_pool.pop_back();
    return 0;
}
// waits for one enqueued batch and reads its status words: *ok = false when the optimistic run is void (not exact in
// fp16, a truncated probe set, or queries left for the exact path) and the batch has to be answered again
static int settle(lira_index* h, PendingBatch& pb, bool* ok) {
    *ok = true;
    if (!pb.enqueued) return 0;
    LIRA_CUDA_OK(cudaEventSynchronize(pb.ev));
    const int* fl = pb.h_flags;
    if (fl[2] != 0) raise_nprobe_cap(h, pb.cap, fl[3]);
    *ok = fl[0] == 0 && fl[1] == 0 && fl[2] == 0;
    if (*ok) { h->last_path = 1; h->last_redo = 0; }
    h->flag_pool.push_back(pb.h_flags);
    h->event_pool.push_back(pb.ev);
    pb.enqueued = false;
    return 0;
}
}  // namespace lira

int lira_probe_search_enqueue_dev(lira_index_t* h, lira_model_t* m, const float* d_q, int64_t ldq, int64_t Q, int mode,
                                  double value, int k, int dedup, float* d_D, int64_t* d_I, int32_t* d_nprobe, int64_t* d_cmp,
                                  void* stream) {
    LIRA_REQUIRE(h && m && d_q && d_D && d_I, "null argument");
    LIRA_REQUIRE(h->device == m->device && h->B == m->B && h->d == m->d, "index and model disagree on device / B / d");
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (!fused_eligible(h, m, Q, mode, k))   // nothing to defer: the synchronous call answers
        return lira_probe_search_dev(h, m, d_q, ldq, Q, mode, value, k, dedup, d_D, d_I, d_nprobe, d_cmp, stream);
    PendingBatch pb;
    if (int rc = take_flags_and_event(h, &pb.h_flags, &pb.ev)) return rc;
    pb.m = m; pb.d_q = d_q; pb.ldq = ldq; pb.Q = Q; pb.mode = mode; pb.value = value; pb.k = k; pb.dedup = dedup;
    pb.d_D = d_D; pb.d_I = (long long*)d_I; pb.d_nprobe = d_nprobe; pb.d_cmp = (long long*)d_cmp; pb.st = st;
    pb.cap = std::min(h->B, h->nprobe_cap);
    h->last_Q = Q;
    h->last_k = k;
    if (h->timing) LIRA_CUDA_OK(cudaEventRecord(h->ev[2], st));
    if (int rc = fused_enqueue(h, m, d_q, ldq, Q, mode, value, k, dedup, d_D, (long long*)d_I, d_nprobe, (long long*)d_cmp, st, pb.h_flags)) return rc;
    LIRA_CUDA_OK(cudaEventRecord(pb.ev, st));
    if (h->timing) {
        LIRA_CUDA_OK(cudaEventRecord(h->ev[3], st));
        h->timing_pending = true;
    }
    pb.enqueued = true;
    h->pending.push_back(pb);
    return 0;
}

int lira_index_finish(lira_index_t* h) {
    LIRA_REQUIRE(h, "null index");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    while (!h->pending.empty()) {
        PendingBatch pb = h->pending.front();
        h->pending.pop_front();
        bool ok = true;
        if (int rc = settle(h, pb, &ok)) return rc;
        if (!ok)   // rare: answer the batch again through the checked path (its inputs and outputs are the caller's, still valid)
            if (int rc = lira_probe_search_dev(h, pb.m, pb.d_q, pb.ldq, pb.Q, pb.mode, pb.value, pb.k, pb.dedup, pb.d_D, (int64_t*)pb.d_I,
                                               pb.d_nprobe, (int64_t*)pb.d_cmp, pb.st)) return rc;
    }
    return 0;
}

int lira_probe_search_submit(lira_index_t* h, lira_model_t* m, const float* q, int64_t Q, int mode, double value, int k,
                             int dedup, int slot) {
    LIRA_REQUIRE(h && m && q && Q >= 0, "null argument");
    LIRA_REQUIRE(slot >= 0 && slot < 4, "slot must be in [0, 4)");
    LIRA_REQUIRE(h->device == m->device && h->B == m->B && h->d == m->d, "index and model disagree on device / B / d");
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    lira_index::HostSlot& s = h->slots[slot];
    LIRA_REQUIRE(!s.busy, "slot still holds a batch: call lira_probe_search_wait first");
    if (!h->st_in) {
        LIRA_CUDA_OK(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
        LIRA_CUDA_OK(cudaStreamCreateWithFlags(&h->st_out, cudaStreamNonBlocking));
    }
    if (!s.ev_in) {
        LIRA_CUDA_OK(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        LIRA_CUDA_OK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        LIRA_CUDA_OK(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    const size_t Qs = (size_t)std::max<int64_t>(Q, 1);
    s.pb = PendingBatch();
    s.pb.m = m; s.pb.Q = Q; s.pb.mode = mode; s.pb.value = value; s.pb.k = k; s.pb.dedup = dedup; s.pb.slot = slot;
    s.pb.cap = std::min(h->B, h->nprobe_cap);
    s.user_q = q;
    s.busy = true;
    s.sync_done = false;
    if (h->ds != h->d || Q == 0 || !fused_eligible(h, m, Q, mode, k)) return 0;   // answered synchronously in wait()
    if (int rc = s.q.ensure(Qs * h->ds * 4)) return rc;
    if (int rc = s.D.ensure(Qs * k * 4)) return rc;
    if (int rc = s.I.ensure(Qs * k * 8)) return rc;
    if (int rc = s.nprobe.ensure(Qs * 4)) return rc;
    if (int rc = s.cmp.ensure(Qs * 8)) return rc;
    const size_t bq = (size_t)Q * h->d * 4;
    const size_t bD = (size_t)Q * k * 4, bI = (size_t)Q * k * 8, bC = (size_t)Q * 8, bN = (size_t)Q * 4;
    s.oI = (bD + 255) & ~(size_t)255; s.oC = s.oI + ((bI + 255) & ~(size_t)255); s.oN = s.oC + ((bC + 255) & ~(size_t)255);
    const size_t total = s.oN + bN;
    if (total > s.pin_out_cap) {
        if (s.pin_out) cudaFreeHost(s.pin_out);
        s.pin_out = nullptr; s.pin_out_cap = 0;
        LIRA_CUDA_OK(cudaHostAlloc(&s.pin_out, total + total / 4, cudaHostAllocDefault));
        s.pin_out_cap = total + total / 4;
    }
    // queries: a caller's pinned array is copied from directly (the copy engine does the work while this thread goes on to
    // enqueue the batch); a pageable one goes through a pinned staging buffer of this handle -- cudaMemcpyAsync from pageable
    // memory would be staged by the driver synchronously and in small pieces (that memcpy is ~0.3 ms of this thread's time
    // per 5 MB batch, the largest host-side cost of a batch: callers that can should hand over pinned arrays).
    // LIRA_STAGE_PINNED=1 stages pinned arrays too (on this pool, pages pinned by another allocator upload at about half
    // the rate of this buffer: worth it only when the upload, not the host thread, is the bottleneck).
    cudaPointerAttributes attr;
    const bool pageable = cudaPointerGetAttributes(&attr, q) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    const void* src = q;
    if (h->stage_pinned < 0 && getenv("LIRA_STAGE_PINNED")) h->stage_pinned = atoi(getenv("LIRA_STAGE_PINNED")) ? 1 : 0;
    s.timed_direct = !pageable && h->stage_pinned < 0;
    if (s.timed_direct && !s.ev_t[0])
        for (int i = 0; i < 4; ++i) LIRA_CUDA_OK(cudaEventCreate(&s.ev_t[i]));
    if (pageable || h->stage_pinned == 1) {
        if (bq > s.pin_in_cap) {
            if (s.pin_in) cudaFreeHost(s.pin_in);
            s.pin_in = nullptr; s.pin_in_cap = 0;
            LIRA_CUDA_OK(cudaHostAlloc(&s.pin_in, bq + bq / 4, cudaHostAllocDefault));
            s.pin_in_cap = bq + bq / 4;
        }
        memcpy(s.pin_in, q, bq);   // (spreading this over helper threads was measured slower: creating them costs more than it saves)
        src = s.pin_in;
    }
    if (s.timed_direct) LIRA_CUDA_OK(cudaEventRecord(s.ev_t[0], h->st_in));
    LIRA_CUDA_OK(cudaMemcpyAsync(s.q.p, src, bq, cudaMemcpyHostToDevice, h->st_in));
    if (s.timed_direct) LIRA_CUDA_OK(cudaEventRecord(s.ev_t[1], h->st_in));
    LIRA_CUDA_OK(cudaEventRecord(s.ev_in, h->st_in));
    cudaStream_t st = h->stream;
    LIRA_CUDA_OK(cudaStreamWaitEvent(st, s.ev_in, 0));
    if (s.timed_direct) LIRA_CUDA_OK(cudaEventRecord(s.ev_t[2], st));
    if (int rc = take_flags_and_event(h, &s.pb.h_flags, &s.pb.ev)) return rc;
    h->last_Q = Q;
    h->last_k = k;
    if (int rc = fused_enqueue(h, m, s.q.as<float>(), h->ds, Q, mode, value, k, dedup, s.D.as<float>(), s.I.as<long long>(), s.nprobe.as<int>(),
                               s.cmp.as<long long>(), st, s.pb.h_flags)) return rc;
    if (s.timed_direct) LIRA_CUDA_OK(cudaEventRecord(s.ev_t[3], st));
    LIRA_CUDA_OK(cudaEventRecord(s.ev_done, st));
    LIRA_CUDA_OK(cudaStreamWaitEvent(h->st_out, s.ev_done, 0));
    char* sg = (char*)s.pin_out;
    LIRA_CUDA_OK(cudaMemcpyAsync(sg, s.D.p, bD, cudaMemcpyDeviceToHost, h->st_out));
    LIRA_CUDA_OK(cudaMemcpyAsync(sg + s.oI, s.I.p, bI, cudaMemcpyDeviceToHost, h->st_out));
    LIRA_CUDA_OK(cudaMemcpyAsync(sg + s.oC, s.cmp.p, bC, cudaMemcpyDeviceToHost, h->st_out));
    LIRA_CUDA_OK(cudaMemcpyAsync(sg + s.oN, s.nprobe.p, bN, cudaMemcpyDeviceToHost, h->st_out));
    LIRA_CUDA_OK(cudaEventRecord(s.pb.ev, h->st_out));
    s.pb.enqueued = true;
    return 0;
}

int lira_probe_search_wait(lira_index_t* h, int slot, float* D, int64_t* I, int32_t* nprobe, int64_t* cmp) {
    LIRA_REQUIRE(h && D && I, "null argument");
    LIRA_REQUIRE(slot >= 0 && slot < 4, "slot must be in [0, 4)");
    LIRA_CUDA_OK(cudaSetDevice(h->device));
    lira_index::HostSlot& s = h->slots[slot];
    LIRA_REQUIRE(s.busy, "no batch was submitted to this slot");
    s.busy = false;
    bool ok = false;
    if (s.pb.enqueued) {
        if (int rc = settle(h, s.pb, &ok)) return rc;
        if (s.timed_direct && h->stage_pinned < 0) {   // (the batch is complete: all four events have fired)
            float up = 0.f, run = 0.f;
            if (cudaEventElapsedTime(&up, s.ev_t[0], s.ev_t[1]) == cudaSuccess && cudaEventElapsedTime(&run, s.ev_t[2], s.ev_t[3]) == cudaSuccess) {
                if (h->direct_samples > 0) { h->direct_h2d_ms += up; h->direct_compute_ms += run; }   // the first one warms up
                if (++h->direct_samples >= 4) {
                    // (staged as soon as the upload comes within 30 % of the kernels' time: measured on a host whose direct upload
                    //  took 0.41 ms against 0.49 ms of kernels, staging was the faster pipeline, 21.2 M against 20.2 M queries/s)
                    h->stage_pinned = h->direct_h2d_ms > 0.7f * h->direct_compute_ms ? 1 : 0;
                    if (h->stage_pinned == 1) {   // every slot's staging buffer now, not in the middle of the caller's steady state
                        const size_t bq = (size_t)s.pb.Q * h->d * 4;
                        for (auto& sl : h->slots)
                            if (bq > sl.pin_in_cap && !sl.busy) {
                                if (sl.pin_in) cudaFreeHost(sl.pin_in);
                                sl.pin_in = nullptr; sl.pin_in_cap = 0;
                                if (cudaHostAlloc(&sl.pin_in, bq + bq / 4, cudaHostAllocDefault) == cudaSuccess) sl.pin_in_cap = bq + bq / 4;
                                else { sl.pin_in = nullptr; cudaGetLastError(); }
                            }
                    }
                    if (getenv("LIRA_DEBUG_STAGE"))
                        fprintf(stderr, "[lira] pinned query arrays: upload %.3f ms vs kernels %.3f ms per batch -> %s\n", h->direct_h2d_ms / 3,
                                h->direct_compute_ms / 3, h->stage_pinned ? "staged through the handle's own pinned buffer" : "uploaded directly");
                }
            }
            cudaGetLastError();
        }
        s.timed_direct = false;
    }
    const PendingBatch& pb = s.pb;
    if (!ok)   // not eligible for the fused flow, or its optimistic run was void: the synchronous call answers
        return lira_probe_search(h, pb.m, s.user_q, pb.Q, pb.mode, pb.value, pb.k, pb.dedup, D, I, nprobe, cmp);
    const char* sg = (const char*)s.pin_out;
    memcpy(D, sg, (size_t)pb.Q * pb.k * 4);
    memcpy(I, sg + s.oI, (size_t)pb.Q * pb.k * 8);
    if (cmp) memcpy(cmp, sg + s.oC, (size_t)pb.Q * 8);
    if (nprobe) memcpy(nprobe, sg + s.oN, (size_t)pb.Q * 4);
    return 0;
}

// ---- exact kNN: the base is resident behind a handle, cut into segments that play the role of lists --------
}  // extern "C"

struct lira_knn_index {
    int device = 0, d = 0, ds = 0, metric = 0;
    long long N = 0, seg = 0;
    int last_path = 0;           // 0 = CUDA cores, 1 = tensor cores, 2 = mixed (some batches of the last search on each)
    const float* base = nullptr; // [N, ds] on the device: dbase (uploaded by lira_knn_create) or the caller's memory (lira_knn_create_dev)
    DevBuf dbase, dids;
    lira_index_t* index = nullptr;
};

namespace lira {
// (re)cut the base into segments of `seg` rows: only the offsets change, the fp16 shadow copy and the norms are per row
static int knn_set_segments(lira_knn_index* kn, long long seg) {
    if (kn->seg == seg) return 0;
    lira_index* h = kn->index;
    const int nseg = (int)((kn->N + seg - 1) / seg);
    std::vector<long long> off(nseg + 1);
    for (int s = 0; s <= nseg; ++s) off[s] = std::min<long long>((long long)s * seg, kn->N);
    std::vector<int> order(nseg);
    std::iota(order.begin(), order.end(), 0);   // equal sizes (the last one may be shorter): already size-descending
    LIRA_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (h->d_offsets) LIRA_CUDA_OK(cudaFree(h->d_offsets));
    if (h->d_list_order) LIRA_CUDA_OK(cudaFree(h->d_list_order));
    h->d_offsets = nullptr;
    h->d_list_order = nullptr;
    LIRA_CUDA_OK(cudaMalloc(&h->d_offsets, (size_t)(nseg + 1) * 8));
    LIRA_CUDA_OK(cudaMalloc(&h->d_list_order, (size_t)nseg * 4));
    LIRA_CUDA_OK(cudaMemcpy(h->d_offsets, off.data(), (size_t)(nseg + 1) * 8, cudaMemcpyHostToDevice));
    LIRA_CUDA_OK(cudaMemcpy(h->d_list_order, order.data(), (size_t)nseg * 4, cudaMemcpyHostToDevice));
    h->h_offsets = off;
    h->B = nseg;
    kn->seg = seg;
    return 0;
}
}  // namespace lira

extern "C" {

static int knn_finish_create(lira_knn_index* kn) {
    const long long N = kn->N;
    cudaStream_t st0 = nullptr;
    LIRA_CUDA_OK(cudaStreamCreateWithFlags(&st0, cudaStreamNonBlocking));
    int r = kn->dids.ensure((size_t)N * 4);
    if (!r) {
        iota_i32_kernel<<<grid_for(N, 256), 256, 0, st0>>>(kn->dids.as<int>(), N);
        g_launches.fetch_add(1);
    }
    cudaError_t e = cudaStreamSynchronize(st0);
    cudaStreamDestroy(st0);
    if (r) return r;
    LIRA_CUDA_OK(e);
    const long long seg = 8192;
    const int nseg = (int)((N + seg - 1) / seg);
    std::vector<int64_t> off(nseg + 1);
    for (int s = 0; s <= nseg; ++s) off[s] = std::min<long long>((long long)s * seg, N);
    if (int r2 = index_create_dev_impl(kn->base, kn->ds, kn->d, off.data(), kn->dids.as<int>(), nseg, kn->metric, kn->device, &kn->index, true)) return r2;
    kn->seg = seg;
    return 0;
}

int lira_knn_create(const float* base, int64_t N, int d, int metric, int device, lira_knn_t** out) {
    LIRA_REQUIRE(out && base && N >= 1 && d >= 1, "bad argument");
    LIRA_REQUIRE(N < (1ll << 31), "N must be below 2^31 per handle");
    LIRA_REQUIRE(metric == LIRA_METRIC_L2 || metric == LIRA_METRIC_IP, "metric must be LIRA_METRIC_L2 or LIRA_METRIC_IP");
    if (int rc = check_device(device)) return rc;
    lira_knn_index* kn = new lira_knn_index();
    kn->device = device; kn->d = d; kn->ds = round_up(d, 4); kn->metric = metric; kn->N = N;
    auto body = [&]() -> int {
        cudaStream_t st0 = nullptr;
        LIRA_CUDA_OK(cudaStreamCreateWithFlags(&st0, cudaStreamNonBlocking));
        int r = upload_rows(kn->dbase, base, N, d, kn->ds, st0);
        cudaError_t e = cudaStreamSynchronize(st0);
        cudaStreamDestroy(st0);
        if (r) return r;
        LIRA_CUDA_OK(e);
        kn->base = kn->dbase.as<float>();
        return knn_finish_create(kn);
    };
    const int rc = body();
    if (rc) { lira_knn_free(kn); return rc; }
    *out = kn;
    return 0;
}

int lira_knn_create_dev(const float* d_base, int64_t ld, int64_t N, int d, int metric, int device, lira_knn_t** out) {
    LIRA_REQUIRE(out && d_base && N >= 1 && d >= 1, "bad argument");
    LIRA_REQUIRE(N < (1ll << 31), "N must be below 2^31 per handle");
    LIRA_REQUIRE(ld >= d && (ld % 4) == 0 && ((uintptr_t)d_base & 15) == 0, "device base needs ld % 4 == 0 and 16-byte alignment");
    LIRA_REQUIRE(metric == LIRA_METRIC_L2 || metric == LIRA_METRIC_IP, "metric must be LIRA_METRIC_L2 or LIRA_METRIC_IP");
    if (int rc = check_device(device)) return rc;
    lira_knn_index* kn = new lira_knn_index();
    kn->device = device; kn->d = d; kn->ds = (int)ld; kn->metric = metric; kn->N = N;
    kn->base = d_base;
    const int rc = knn_finish_create(kn);
    if (rc) { lira_knn_free(kn); return rc; }
    *out = kn;
    return 0;
}

int lira_knn_free(lira_knn_t* kn) {
    if (!kn) return 0;
    cudaSetDevice(kn->device);
    if (kn->index) lira_index_free(kn->index);
    kn->dbase.release();
    kn->dids.release();
    delete kn;
    return 0;
}

int64_t lira_knn_ntotal(const lira_knn_t* kn) { return kn ? kn->N : -1; }
int lira_knn_last_path(const lira_knn_t* kn) { return kn ? kn->last_path : -1; }
int lira_knn_last_redo(const lira_knn_t* kn) { return (kn && kn->index) ? kn->index->last_redo : -1; }
int lira_knn_last_scan_kind(const lira_knn_t* kn) { return (kn && kn->index) ? lira_index_last_scan_kind(kn->index) : -1; }
int lira_knn_set_use_tensor_cores(lira_knn_t* kn, int enable) {
    LIRA_REQUIRE(kn && kn->index, "null handle");
    kn->index->use_tc = enable != 0;
    return 0;
}

// one search over device queries; host_q != nullptr: the batches are uploaded from there (d_q ignored) and the results copied back
static int knn_search_impl(lira_knn_t* kn, const float* host_q, const float* d_q, long long ldq, int64_t Q, int k, float* D, int64_t* I,
                           cudaStream_t st_user) {
    lira_index* h = kn->index;
    const int d = kn->d, ds = kn->ds;
    // k <= 16 on a large base: long segments (the in-kernel compaction keeps their candidate regions small), i.e. 8x fewer
    // (query, segment) pairs, regions and refine work, and larger query batches
    long long seg = (k <= 16 && kn->N >= 8 * 65536 && !getenv("LIRA_KNN_SHORT_SEGMENTS")) ? 65536 : 8192;
    if (h->u8_ok) {
        // byte copy: the seed pass scans the first U8_SEED_LISTS segments (a sample of the base), so short segments keep it
        // cheap; at most ~512 segments so that a batch of (query, segment) pairs still holds thousands of queries
        seg = 8192;
        while (kn->N / seg > 512) seg *= 2;
    }
    if (int rc = knn_set_segments(kn, seg)) return rc;
    const int nseg = h->B;
    cudaStream_t st = st_user ? st_user : h->stream;
    Workspace& ws = h->ws;
    int redo_total = 0, n_tc = 0, n_batches = 0;
    const long long max_pairs = 4ll << 20;   // (query, segment) pairs per batch: bounds the per-pair workspaces (1 KiB each on the tensor-core path)
    const long long qb = std::max<long long>(1, std::min<long long>(std::max<int64_t>(Q, 1), max_pairs / nseg));
    for (long long q0 = 0; q0 < Q; q0 += qb) {
        const long long nq = std::min<long long>(qb, Q - q0);
        ProbeSpec ps;
        ps.kind = 2;
        if (host_q) {
            if (int r3 = upload_rows(ws.q, host_q + q0 * d, nq, d, ds, st)) return r3;
            if (int r3 = ws.D.ensure((size_t)nq * k * 4)) return r3;
            if (int r3 = ws.I.ensure((size_t)nq * k * 8)) return r3;
            if (int r3 = search_core(h, ws.q.as<float>(), ds, nq, ps, k, 1, ws.D.as<float>(), ws.I.as<long long>(), nullptr, nullptr, st)) return r3;
            LIRA_CUDA_OK(cudaMemcpyAsync(D + (size_t)q0 * k, ws.D.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaMemcpyAsync(I + (size_t)q0 * k, ws.I.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
        } else {
            if (int r3 = search_core(h, d_q + q0 * ldq, ldq, nq, ps, k, 1, D + (size_t)q0 * k, (long long*)I + (size_t)q0 * k, nullptr, nullptr, st)) return r3;
        }
        redo_total += h->last_redo;
        n_tc += h->last_path == 1;
        ++n_batches;
    }
    h->last_redo = redo_total;
    kn->last_path = n_tc == 0 ? 0 : (n_tc == n_batches ? 1 : 2);
    return 0;
}

int lira_knn_search(lira_knn_t* kn, const float* query, int64_t Q, int k, float* D, int64_t* I) {
    LIRA_REQUIRE(kn && kn->index && query && D && I && Q >= 0, "bad argument");
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    LIRA_CUDA_OK(cudaSetDevice(kn->device));
    return knn_search_impl(kn, query, nullptr, 0, Q, k, D, I, nullptr);
}

int lira_knn_search_dev(lira_knn_t* kn, const float* d_query, int64_t ldq, int64_t Q, int k, float* d_D, int64_t* d_I, void* stream) {
    LIRA_REQUIRE(kn && kn->index && d_query && d_D && d_I && Q >= 0, "bad argument");
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    LIRA_REQUIRE(ldq >= kn->d && (ldq % 4) == 0 && ((uintptr_t)d_query & 15) == 0, "device queries need ld % 4 == 0 and 16-byte alignment");
    LIRA_CUDA_OK(cudaSetDevice(kn->device));
    return knn_search_impl(kn, nullptr, d_query, ldq, Q, k, d_D, d_I, (cudaStream_t)stream);
}

int lira_knn(const float* base, int64_t N, const float* query, int64_t Q, int d, int k, int metric, int device,
             float* D, int64_t* I) {
    LIRA_REQUIRE(base && query && D && I && N >= 1 && Q >= 0 && d >= 1, "bad argument");
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    lira_knn_t* kn = nullptr;
    if (int rc = lira_knn_create(base, N, d, metric, device, &kn)) return rc;
    const int rc = lira_knn_search(kn, query, Q, k, D, I);
    lira_knn_free(kn);
    return rc;
}

// ---- build side: K-Means, scaler statistics, redundancy rule --------------------------------------------------
int lira_kmeans_train_dev(const float* d_x, int64_t ld, int64_t n, int d, int B, int niter, int init_given, float* d_centroids,
                          int64_t ldc, const int64_t* d_init_rows, int device, void* stream) {
    LIRA_REQUIRE(d_x && d_centroids && n >= 1 && d >= 1 && B >= 1 && niter >= 0, "bad argument");
    LIRA_REQUIRE(n >= B, "K-Means needs at least as many points as centroids");
    LIRA_REQUIRE(ld >= d && (ld % 4) == 0 && ((uintptr_t)d_x & 15) == 0, "device rows need ld % 4 == 0 and 16-byte alignment");
    LIRA_REQUIRE(ldc >= d && (ldc % 4) == 0 && ((uintptr_t)d_centroids & 15) == 0, "device centroids need ld % 4 == 0 and 16-byte alignment");
    LIRA_REQUIRE(init_given || d_init_rows, "initial centroids or the rows to take them from are needed");
    if (int rc = check_device(device)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf sums, counts, assign, dist;
    const int lds = round_up(d, 4);
    int rc = 0;
    auto body = [&]() -> int {
        if (int r = sums.ensure((size_t)B * lds * 4)) return r;
        if (int r = counts.ensure((size_t)B * 4)) return r;
        if (int r = assign.ensure((size_t)n * 8)) return r;
        if (int r = dist.ensure((size_t)n * 4)) return r;
        if (!init_given) {   // centroids = the given rows of x (faiss: k points of a random permutation)
            LIRA_CUDA_OK(cudaMemsetAsync(d_centroids, 0, (size_t)B * ldc * 4, st));
            gather_rows64_kernel<<<grid_for((long long)B * (ldc / 4), 256), 256, 0, st>>>(d_x, ld, d, (const long long*)d_init_rows, B, d_centroids, (int)ldc);
            LIRA_LAUNCH_CHECK();
        }
        std::vector<int> h_counts(B);
        std::vector<float> h_cent;
        for (int it = 0; it < niter; ++it) {
            // assignment: exact nearest centroid (ties to the lower id), the kNN path with the centroid table as base
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
            lira_knn_t* kn = nullptr;
            if (int r = lira_knn_create_dev(d_centroids, ldc, B, d, LIRA_METRIC_L2, device, &kn)) return r;
            int r = lira_knn_search_dev(kn, d_x, ld, n, 1, dist.as<float>(), (int64_t*)assign.p, st);
            if (!r && cudaStreamSynchronize(st) != cudaSuccess) { set_error("K-Means assignment failed"); r = 2; }
            lira_knn_free(kn);
            if (r) return r;
            // update: segmented mean
            LIRA_CUDA_OK(cudaMemsetAsync(sums.p, 0, (size_t)B * lds * 4, st));
            LIRA_CUDA_OK(cudaMemsetAsync(counts.p, 0, (size_t)B * 4, st));
            kmeans_accumulate_kernel<<<(int)((n + 7) / 8), 256, 0, st>>>(d_x, ld, d, n, assign.as<long long>(), sums.as<float>(), lds, counts.as<int>());
            LIRA_LAUNCH_CHECK();
            kmeans_finalize_kernel<<<grid_for((long long)B * d, 256), 256, 0, st>>>(sums.as<float>(), lds, counts.as<int>(), B, d, d_centroids, ldc);
            LIRA_LAUNCH_CHECK();
            LIRA_CUDA_OK(cudaMemcpyAsync(h_counts.data(), counts.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
            // empty clusters: split the currently largest one (faiss clustering: both halves get the centroid, perturbed by
            // +-1/1024 in alternating dimensions, and the donor's count is halved)
            bool any_empty = false;
            for (int b = 0; b < B; ++b) any_empty |= h_counts[b] == 0;
            if (any_empty) {
                h_cent.resize((size_t)B * ldc);
                LIRA_CUDA_OK(cudaMemcpy(h_cent.data(), d_centroids, (size_t)B * ldc * 4, cudaMemcpyDeviceToHost));
                for (int b = 0; b < B; ++b) {
                    if (h_counts[b] != 0) continue;
                    const int donor = (int)(std::max_element(h_counts.begin(), h_counts.end()) - h_counts.begin());
                    if (h_counts[donor] < 2) break;
                    const float eps = 1.0f / 1024.0f;
                    for (int j = 0; j < d; ++j) {
                        const float c = h_cent[(size_t)donor * ldc + j];
                        h_cent[(size_t)b * ldc + j] = c * ((j & 1) ? 1.0f - eps : 1.0f + eps);
                        h_cent[(size_t)donor * ldc + j] = c * ((j & 1) ? 1.0f + eps : 1.0f - eps);
                    }
                    h_counts[b] = h_counts[donor] / 2;
                    h_counts[donor] -= h_counts[b];
                }
                LIRA_CUDA_OK(cudaMemcpy(d_centroids, h_cent.data(), (size_t)B * ldc * 4, cudaMemcpyHostToDevice));
            }
        }
        return 0;
    };
    rc = body();
    cudaStreamSynchronize(st);
    for (DevBuf* b : {&sums, &counts, &assign, &dist}) b->release();
    return rc;
}

int lira_kmeans_train(const float* x, int64_t n, int d, int B, int niter, uint64_t seed, const float* init_centroids, int device,
                      float* centroids_out) {
    LIRA_REQUIRE(x && centroids_out && n >= 1 && d >= 1 && B >= 1 && niter >= 0, "bad argument");
    LIRA_REQUIRE(n >= B, "K-Means needs at least as many points as centroids");
    if (int rc = check_device(device)) return rc;
    // faiss clustering: at most 256 points per centroid take part (a random subset, selection sampling: one pass, no O(n) memory)
    std::mt19937_64 rng(seed);
    const int64_t cap = 256ll * B;
    const int64_t ns = std::min<int64_t>(n, cap);
    const int ds = round_up(d, 4);
    std::vector<float> sample;
    const float* src = x;
    if (ns < n || ds != d) {
        sample.assign((size_t)ns * ds, 0.0f);
        int64_t need = ns, w = 0;
        for (int64_t i = 0; i < n && need > 0; ++i) {
            const uint64_t remaining = (uint64_t)(n - i);
            if (ns == n || (rng() % remaining) < (uint64_t)need) {
                memcpy(sample.data() + (size_t)w * ds, x + (size_t)i * d, (size_t)d * 4);
                ++w;
                --need;
            }
        }
        src = sample.data();
    }
    // initial centroids: B distinct sample rows
    std::vector<int64_t> rows(ns);
    std::iota(rows.begin(), rows.end(), 0ll);
    for (int b = 0; b < B; ++b) std::swap(rows[b], rows[b + (int64_t)(rng() % (uint64_t)(ns - b))]);
    DevBuf dx, dc, dr;
    cudaStream_t st = nullptr;
    auto body = [&]() -> int {
        LIRA_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (int r = dx.ensure((size_t)ns * ds * 4)) return r;
        if (int r = dc.ensure((size_t)B * ds * 4)) return r;
        if (int r = dr.ensure((size_t)B * 8)) return r;
        if (int r = upload_bytes(dx.p, src, (size_t)ns * ds * 4, st)) return r;
        LIRA_CUDA_OK(cudaMemcpyAsync(dr.p, rows.data(), (size_t)B * 8, cudaMemcpyHostToDevice, st));
        if (init_centroids) {
            LIRA_CUDA_OK(cudaMemsetAsync(dc.p, 0, (size_t)B * ds * 4, st));
            LIRA_CUDA_OK(cudaMemcpy2DAsync(dc.p, (size_t)ds * 4, init_centroids, (size_t)d * 4, (size_t)d * 4, (size_t)B, cudaMemcpyHostToDevice, st));
        }
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        if (int r = lira_kmeans_train_dev(dx.as<float>(), ds, ns, d, B, niter, init_centroids ? 1 : 0, dc.as<float>(), ds, (const int64_t*)dr.p, device, st)) return r;
        LIRA_CUDA_OK(cudaMemcpy2DAsync(centroids_out, (size_t)d * 4, dc.p, (size_t)ds * 4, (size_t)d * 4, (size_t)B, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    };
    const int rc = body();
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (DevBuf* b : {&dx, &dc, &dr}) b->release();
    return rc;
}

int lira_centroid_features_dev(const float* d_q, int64_t ldq, int64_t Q, const float* d_centroids, int64_t ldc, int B, int d,
                               const float* d_mean, const float* d_scale, float* d_out, int64_t ldo, int device, void* stream) {
    LIRA_REQUIRE(d_q && d_centroids && d_out && B >= 1 && d >= 1 && Q >= 0, "bad argument");
    LIRA_REQUIRE((d_mean == nullptr) == (d_scale == nullptr), "mean and scale must be given together");
    LIRA_REQUIRE((ldq % 4) == 0 && (ldc % 4) == 0 && (ldo % 4) == 0 && ((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_centroids & 15) == 0 &&
                 ((uintptr_t)d_out & 15) == 0, "device matrices need row strides that are multiples of 4 and 16-byte alignment");
    LIRA_REQUIRE(ldq >= d && ldc >= d && ldo >= B, "row strides too small");
    if (int rc = check_device(device)) return rc;
    if (Q == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tm;
    if (int r = make_tmap(&tm, d_centroids, B, round_up(d, 4), ldc)) return r;
    const long long chunk = 1ll << 20;
    for (long long q0 = 0; q0 < Q; q0 += chunk) {
        const long long nq = std::min<long long>(chunk, Q - q0);
        DenseParams p{d_q + q0 * ldq, (long)ldq, (int)nq, round_up(d, 4), B, d_out + q0 * ldo, (long)ldo, 0, d_mean, d_scale};
        if (int r = launch_dense<64, OP_L2, EPI_FEATURE>(tm, p, st)) return r;
    }
    return 0;
}

int lira_feature_stats_dev(const float* d_x, int64_t ld, int64_t n, const float* d_centroids, int64_t ldc, int B, int d,
                           double* mean_out, double* var_out, int device, void* stream) {
    LIRA_REQUIRE(d_x && d_centroids && mean_out && var_out && n >= 1 && B >= 1 && d >= 1, "bad argument");
    if (int rc = check_device(device)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int Bp = round_up(B, 4);
    const long long chunk = std::min<long long>(n, 65536);
    DevBuf feats, acc;
    auto body = [&]() -> int {
        if (int r = feats.ensure((size_t)chunk * Bp * 4)) return r;
        if (int r = acc.ensure((size_t)2 * B * 8)) return r;
        LIRA_CUDA_OK(cudaMemsetAsync(acc.p, 0, (size_t)2 * B * 8, st));
        for (long long a = 0; a < n; a += chunk) {
            const long long m = std::min<long long>(chunk, n - a);
            if (int r = lira_centroid_features_dev(d_x + a * ld, ld, m, d_centroids, ldc, B, d, nullptr, nullptr, feats.as<float>(), Bp, device, st)) return r;
            dim3 grid((B + 31) / 32, (unsigned)std::min<long long>(256, (m + 7) / 8));
            feature_stats_kernel<<<grid, 256, 0, st>>>(feats.as<float>(), Bp, m, B, acc.as<double>(), acc.as<double>() + B);
            LIRA_LAUNCH_CHECK();
        }
        std::vector<double> h(2 * (size_t)B);
        LIRA_CUDA_OK(cudaMemcpyAsync(h.data(), acc.p, (size_t)2 * B * 8, cudaMemcpyDeviceToHost, st));
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        for (int b = 0; b < B; ++b) {
            const double mu = h[b] / (double)n;
            mean_out[b] = mu;
            var_out[b] = std::max(h[B + b] / (double)n - mu * mu, 0.0);
        }
        return 0;
    };
    const int rc = body();
    cudaStreamSynchronize(st);
    feats.release();
    acc.release();
    return rc;
}

int lira_mul_partition_dev(const float* d_score, int64_t lds, int64_t n_rows, int B, float sigma, const int64_t* d_points, int64_t first,
                           int n_mul, int32_t* d_data_2_bkt, int32_t* d_added, int device, void* stream) {
    LIRA_REQUIRE(d_score && d_data_2_bkt && d_added && n_rows >= 0 && B >= 1, "bad argument");
    LIRA_REQUIRE(n_mul >= 2 && n_mul <= 8, "n_mul must be in [2, 8]");
    LIRA_REQUIRE(lds >= B, "score row stride too small");
    if (int rc = check_device(device)) return rc;
    if (n_rows == 0) return 0;
    mul_partition_kernel<<<(int)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(d_score, (long)lds, n_rows, B, sigma, (const long long*)d_points, first,
                                                                                 n_mul, d_data_2_bkt, d_added);
    LIRA_LAUNCH_CHECK();
    return 0;
}

// ---- a12: IVF-approximate self-kNN (compute_knn.cpp:158-203: IndexIVFFlat(quantizer, d, nlist), train, add, nprobe) ----------
namespace {
__global__ void narrow_probe_ids_kernel(const long long* __restrict__ in, long long n, int* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (int)in[i];
}
}  // namespace

int lira_knn_ivf(const float* base, int64_t N, int d, int k, int nlist, int nprobe, uint64_t seed, int device, float* D, int64_t* I) {
    LIRA_REQUIRE(base && D && I && N >= 1 && d >= 1, "bad argument");
    LIRA_REQUIRE(k >= 1 && k <= 128, "k must be in [1, 128]");
    LIRA_REQUIRE(nlist >= 1 && nlist <= N && nprobe >= 1, "nlist must be in [1, N] and nprobe >= 1");
    LIRA_REQUIRE(N < (1ll << 31), "N must be below 2^31");
    if (int rc = check_device(device)) return rc;
    nprobe = std::min(nprobe, nlist);
    const int ds = round_up(d, 4);
    DevBuf dbase, dcent, dI, dD, dvecs, dids, dpo, dpi, dDo, dIo;
    lira_knn_t* kc = nullptr;
    lira_index_t* index = nullptr;
    cudaStream_t st = nullptr;
    auto body = [&]() -> int {
        LIRA_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        // train: K-Means on (a sample of) the base, as IndexIVFFlat::train hands its quantizer to faiss clustering
        std::vector<float> cent((size_t)nlist * d);
        if (int r = lira_kmeans_train(base, N, d, nlist, 20, seed, nullptr, device, cent.data())) return r;
        if (int r = upload_rows(dbase, base, N, d, ds, st)) return r;
        if (int r = upload_rows(dcent, cent.data(), nlist, d, ds, st)) return r;
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        if (int r = lira_knn_create_dev(dcent.as<float>(), ds, nlist, d, LIRA_METRIC_L2, device, &kc)) return r;
        // add: every vector goes to the list of its nearest centroid
        const long long chunk = 1ll << 20;
        if (int r = dI.ensure((size_t)std::min<long long>(chunk, N) * std::max(nprobe, 1) * 8)) return r;
        if (int r = dD.ensure((size_t)std::min<long long>(chunk, N) * std::max(nprobe, 1) * 4)) return r;
        std::vector<long long> a1((size_t)N);
        for (long long a = 0; a < N; a += chunk) {
            const long long m = std::min<long long>(chunk, N - a);
            if (int r = lira_knn_search_dev(kc, dbase.as<float>() + a * ds, ds, m, 1, dD.as<float>(), (int64_t*)dI.p, st)) return r;
            LIRA_CUDA_OK(cudaMemcpyAsync(a1.data() + a, dI.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
        }
        std::vector<int64_t> off((size_t)nlist + 1, 0);
        for (long long i = 0; i < N; ++i) off[(size_t)a1[i] + 1]++;
        for (int b = 0; b < nlist; ++b) off[b + 1] += off[b];
        std::vector<int32_t> ids((size_t)N);
        {
            std::vector<int64_t> cur(off.begin(), off.end() - 1);
            for (long long i = 0; i < N; ++i) ids[(size_t)cur[(size_t)a1[i]]++] = (int32_t)i;   // ids ascending inside a list (add order)
        }
        if (int r = dids.ensure((size_t)N * 4)) return r;
        if (int r = dvecs.ensure((size_t)N * ds * 4)) return r;
        LIRA_CUDA_OK(cudaMemcpyAsync(dids.p, ids.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
        gather_rows_kernel<<<grid_for(N * (ds / 4), 256, 148 * 16), 256, 0, st>>>(dbase.as<float>(), ds, ds, dids.as<int>(), N, dvecs.as<float>(), ds);
        LIRA_LAUNCH_CHECK();
        LIRA_CUDA_OK(cudaStreamSynchronize(st));
        if (int r = lira_index_create_dev(dvecs.as<float>(), ds, d, off.data(), dids.as<int>(), nlist, LIRA_METRIC_L2, device, &index)) return r;
        // search: the nprobe nearest lists of every vector, exact scan inside them
        const long long qb = std::max<long long>(256, std::min<long long>(32768, (2ll << 20) / nprobe));
        if (int r = dI.ensure((size_t)qb * nprobe * 8)) return r;
        if (int r = dD.ensure((size_t)qb * nprobe * 4)) return r;
        if (int r = dpo.ensure((size_t)(qb + 1) * 8)) return r;
        if (int r = dpi.ensure((size_t)qb * nprobe * 4)) return r;
        if (int r = dDo.ensure((size_t)qb * k * 4)) return r;
        if (int r = dIo.ensure((size_t)qb * k * 8)) return r;
        iota_offsets_kernel<<<grid_for(qb + 1, 256), 256, 0, st>>>(dpo.as<long long>(), qb, nprobe);
        LIRA_LAUNCH_CHECK();
        for (long long a = 0; a < N; a += qb) {
            const long long m = std::min<long long>(qb, N - a);
            const float* q = dbase.as<float>() + a * ds;
            if (int r = lira_knn_search_dev(kc, q, ds, m, nprobe, dD.as<float>(), (int64_t*)dI.p, st)) return r;
            narrow_probe_ids_kernel<<<grid_for(m * nprobe, 256), 256, 0, st>>>(dI.as<long long>(), m * nprobe, dpi.as<int>());
            LIRA_LAUNCH_CHECK();
            if (int r = lira_search_dev(index, q, ds, m, (const int64_t*)dpo.p, dpi.as<int>(), m * nprobe, k, 0, dDo.as<float>(), (int64_t*)dIo.p, nullptr, st)) return r;
            LIRA_CUDA_OK(cudaMemcpyAsync(D + (size_t)a * k, dDo.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaMemcpyAsync(I + (size_t)a * k, dIo.p, (size_t)m * k * 8, cudaMemcpyDeviceToHost, st));
            LIRA_CUDA_OK(cudaStreamSynchronize(st));
        }
        return 0;
    };
    const int rc = body();
    if (st) cudaStreamSynchronize(st);
    if (index) lira_index_free(index);
    if (kc) lira_knn_free(kc);
    if (st) cudaStreamDestroy(st);
    for (DevBuf* b : {&dbase, &dcent, &dI, &dD, &dvecs, &dids, &dpo, &dpi, &dDo, &dIo}) b->release();
    return rc;
}

// ---- multi-GPU merge ---------------------------------------------------------------------------
namespace {
__global__ void pack_keys_kernel(const float* D, const long long* I, long long n, int is_ip, unsigned long long* keys) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long id = I[i];
        keys[i] = id < 0 ? KEY_INF : make_key(is_ip ? -D[i] : D[i], (uint32_t)id);
    }
}
}  // namespace

int lira_pack_keys_dev(const float* d_D, const int64_t* d_I, int64_t n, int metric, uint64_t* d_keys, int device,
                       void* stream) {
    LIRA_REQUIRE(d_D && d_I && d_keys && n >= 0, "bad argument");
    if (int rc = check_device(device)) return rc;
    if (n == 0) return 0;
    pack_keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_D, (const long long*)d_I, n, metric == LIRA_METRIC_IP,
                                                                         (unsigned long long*)d_keys);
    LIRA_LAUNCH_CHECK();
    return 0;
}

int lira_merge_ranks_dev(const uint64_t* d_keys_in, int R, int64_t Q, int k, int metric, int dedup, float* d_D,
                         int64_t* d_I, int device, void* stream) {
    LIRA_REQUIRE(d_keys_in && d_D && d_I && R >= 1 && Q >= 0 && k >= 1 && k <= 128, "bad argument");
    LIRA_REQUIRE(Q * (long long)R < (1ll << 31), "Q * R too large");
    if (int rc = check_device(device)) return rc;
    if (Q == 0) return 0;
    MergeParams mp{(const unsigned long long*)d_keys_in, nullptr, nullptr, k, (int)Q, dedup, d_D, (long long*)d_I,
                   metric == LIRA_METRIC_IP, nullptr, R};
    return launch_merge(mp, (cudaStream_t)stream);
}

}  // extern "C"
