// compute_knn: self-kNN ground truth of a dataset, the command line and file formats of the reference's
// compute_knn.cpp (argv :100-104, input :111-137, exact branch :208-259, outputs :262-290), with the brute-force
// search on the GPU through liblira_b200 (lira_knn) instead of faiss::IndexFlatL2.
//
//   compute_knn <dataset> <data_path> <k> [nprobe] [n_threads]
//
// reads  {data_path}/{dataset}/{dataset}_base.fvecs (or .bvecs)
// writes {data_path}/{dataset}/knn_cache/{dataset}-data_self_knn{k}-n{n}.bin   (raw int32 [n, k]) and .bin.meta
//
// nprobe as in the reference (compute_knn.cpp:103, 150-203): 0 = exact brute force (lira_knn); any other value = the IVF
// approximation (lira_knn_ivf: K-Means with the reference's nlist rule, every vector searched in its nprobe nearest lists;
// negative or absent = the reference's automatic nprobe) and the `_ivf_nprobe{p}` suffix utils.compute_data_knn looks for
// first (utils.py:245-266). On a B200 the exact search is affordable at any size the reference targets, so `0` is the
// recommended value. n_threads is accepted and ignored.
//
// Not in the reference (opt-in):  compute_knn <dataset> <data_path> <k> --queries
//   exact ground truth of the QUERY set against the base: reads {dataset}_query.fvecs (or .bvecs) and writes
//   {data_path}/{dataset}/{dataset}_groundtruth.ivecs (per row: int32 k, then the k base ids, best first) -- the file
//   utils.load_data (utils.py:54-76) and search.cpp:341-343 read. The reference ships it with the datasets.
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "lira_b200.h"

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count();
}

// .fvecs / .bvecs: per vector a little-endian int32 dimension followed by `dim` elements
template <class T>
bool read_vecs(const std::string& path, std::vector<float>& out, int64_t& n, int& dim) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    int32_t d = 0;
    if (std::fread(&d, 4, 1, f) != 1 || d <= 0) { std::fclose(f); return false; }
    std::fseek(f, 0, SEEK_END);
    const long long bytes = std::ftell(f);
    const long long rec = 4 + (long long)d * sizeof(T);
    if (bytes % rec != 0) { std::fclose(f); std::cerr << "Error: " << path << " is not a whole number of records" << std::endl; return false; }
    n = bytes / rec;
    dim = d;
    std::fseek(f, 0, SEEK_SET);
    out.resize((size_t)n * d);
    std::vector<unsigned char> buf(rec);
    for (int64_t i = 0; i < n; ++i) {
        if (std::fread(buf.data(), 1, rec, f) != (size_t)rec) { std::fclose(f); return false; }
        const T* src = reinterpret_cast<const T*>(buf.data() + 4);
        for (int j = 0; j < d; ++j) out[(size_t)i * d + j] = (float)src[j];
    }
    std::fclose(f);
    return true;
}

bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cout << "Usage: " << argv[0] << " <dataset> <data_path> <k> [nprobe] [n_threads]   |   <dataset> <data_path> <k> --queries" << std::endl;
        std::cout << "  nprobe: 0 = exact search on the GPU (recommended), > 0 = IVF approximation, absent / negative = automatic" << std::endl;
        return 1;
    }
    const std::string dataset = argv[1], data_path = argv[2];
    const int k = std::atoi(argv[3]);
    bool query_mode = false;
    for (int a = 4; a < argc; ++a) query_mode |= std::string(argv[a]) == "--queries";
    const int nprobe_arg = (argc > 4 && !query_mode) ? std::atoi(argv[4]) : -1;   // -1: automatic (compute_knn.cpp:103)
    if (k < 1 || k > (query_mode ? 128 : 127)) { std::cerr << "Error: k must be in [1, " << (query_mode ? 128 : 127) << "]" << std::endl; return 1; }

    std::cout << "=== GPU KNN Computation (liblira_b200) ===" << std::endl;
    std::cout << "Dataset: " << dataset << std::endl;
    std::cout << "K: " << k << std::endl;

    const std::string dir = data_path + "/" + dataset;
    std::string base_file = dir + "/" + dataset + "_base.fvecs";
    bool bvecs = false;
    if (!exists(base_file)) {
        base_file = dir + "/" + dataset + "_base.bvecs";
        bvecs = true;
        if (!exists(base_file)) {
            std::cerr << "Error: Cannot find base file for dataset " << dataset << std::endl;
            return 1;
        }
    }
    int64_t n = 0;
    int dim = 0;
    std::vector<float> data;
    const double t0 = now_s();
    std::cout << "Reading " << (bvecs ? ".bvecs" : ".fvecs") << " file..." << std::endl;
    const bool ok = bvecs ? read_vecs<uint8_t>(base_file, data, n, dim) : read_vecs<float>(base_file, data, n, dim);
    if (!ok) { std::cerr << "Error: Cannot open file " << base_file << std::endl; return 1; }
    const double read_time = now_s() - t0;
    std::cout << "Loaded " << n << " vectors of dimension " << dim << std::endl;
    std::cout << "Read time: " << read_time << "s" << std::endl;

    if (query_mode) {
        // ---- ground truth of the query set: k exact neighbours of every query, written as .ivecs ----
        std::string qfile = dir + "/" + dataset + "_query.fvecs";
        bool qb = false;
        if (!exists(qfile)) { qfile = dir + "/" + dataset + "_query.bvecs"; qb = true; }
        if (!exists(qfile)) { std::cerr << "Error: Cannot find query file for dataset " << dataset << std::endl; return 1; }
        int64_t nq = 0;
        int qdim = 0;
        std::vector<float> queries;
        if (!(qb ? read_vecs<uint8_t>(qfile, queries, nq, qdim) : read_vecs<float>(qfile, queries, nq, qdim))) {
            std::cerr << "Error: Cannot open file " << qfile << std::endl;
            return 1;
        }
        if (qdim != dim) { std::cerr << "Error: query dimension " << qdim << " != base dimension " << dim << std::endl; return 1; }
        std::cout << "Loaded " << nq << " queries" << std::endl;
        std::vector<float> Dq((size_t)nq * k);
        std::vector<int64_t> Iq((size_t)nq * k);
        const double tq = now_s();
        if (lira_knn(data.data(), n, queries.data(), nq, dim, k, LIRA_METRIC_L2, 0, Dq.data(), Iq.data()) != 0) {
            std::cerr << "Error: " << lira_last_error() << std::endl;
            return 1;
        }
        const double search_time = now_s() - tq;
        std::cout << "Search time: " << search_time << "s" << std::endl;
        const std::string out = dir + "/" + dataset + "_groundtruth.ivecs";
        std::ofstream f(out, std::ios::binary);
        if (!f) { std::cerr << "Error: Cannot write " << out << std::endl; return 1; }
        std::vector<int32_t> row(k + 1);
        row[0] = k;
        for (int64_t i = 0; i < nq; ++i) {
            for (int j = 0; j < k; ++j) row[j + 1] = (int32_t)Iq[(size_t)i * k + j];
            f.write(reinterpret_cast<const char*>(row.data()), (std::streamsize)(row.size() * sizeof(int32_t)));
        }
        std::cout << std::endl << "=== Summary ===" << std::endl;
        std::cout << "Total time: " << (read_time + search_time) << "s" << std::endl;
        std::cout << "Output file: " << out << std::endl;
        return 0;
    }

    // k + 1 neighbours of every base vector, column 0 (the vector itself) dropped: compute_knn.cpp:237, 254-259
    std::vector<float> D((size_t)n * (k + 1));
    std::vector<int64_t> I((size_t)n * (k + 1));
    const bool approximate = nprobe_arg != 0;
    int n_list = 0, actual_nprobe = 0;
    double build_time = 0.0;
    double t1 = now_s();
    if (approximate) {
        // compute_knn.cpp:160-168 (number of lists) and :190-199 (automatic nprobe)
        const int root = (int)std::sqrt((double)n);
        n_list = std::max(1, n < 50000 ? std::min(root, 256) : n < 1000000 ? std::min(root, 1024) : std::min(root, 4096));
        if (nprobe_arg < 0) actual_nprobe = n < 100000 ? std::min(std::max(n_list / 4, 16), 64) : std::min(std::max(n_list / 8, 32), 128);
        else actual_nprobe = nprobe_arg;
        std::cout << "Using IVF index with " << n_list << " clusters" << std::endl;
        std::cout << "Set nprobe = " << actual_nprobe << " (out of " << n_list << " clusters)" << std::endl;
        std::cout << "Method: Approximate IVF search" << std::endl;
        if (lira_knn_ivf(data.data(), n, dim, k + 1, n_list, actual_nprobe, 1234, 0, D.data(), I.data()) != 0) {
            std::cerr << "Error: " << lira_last_error() << std::endl;
            return 1;
        }
    } else {
        // compute_knn.cpp:208-246: IndexFlatL2.add (here: the base goes to the device behind a handle, shadow copies included) is
        // the index build time, the batched search of every base vector the search time
        std::cout << "Method: Exact FLAT search" << std::endl;
        lira_knn_t* kn = nullptr;
        if (lira_knn_create(data.data(), n, dim, LIRA_METRIC_L2, 0, &kn) != 0) {
            std::cerr << "Error: " << lira_last_error() << std::endl;
            return 1;
        }
        build_time = now_s() - t1;
        std::cout << "Index build time: " << build_time << "s" << std::endl;
        t1 = now_s();
        const int rc = lira_knn_search(kn, data.data(), n, k + 1, D.data(), I.data());
        if (rc != 0) std::cerr << "Error: " << lira_last_error() << std::endl;
        lira_knn_free(kn);
        if (rc != 0) return 1;
    }
    const double search_time = now_s() - t1;
    std::cout << "Search time: " << search_time << "s" << std::endl;
    std::cout << "Average: " << (search_time / n * 1000) << " ms/query" << std::endl;

    std::vector<int32_t> knn((size_t)n * k);
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) knn[(size_t)i * k + j] = (int32_t)I[(size_t)i * (k + 1) + j + 1];

    const std::string cache_dir = dir + "/knn_cache";
    mkdir(cache_dir.c_str(), 0755);
    const std::string suffix = approximate ? ("_ivf_nprobe" + std::to_string(actual_nprobe)) : "";
    const std::string out = cache_dir + "/" + dataset + "-data_self_knn" + std::to_string(k) + "-n" + std::to_string(n) + suffix + ".bin";
    {
        std::ofstream f(out, std::ios::binary);
        if (!f) { std::cerr << "Error: Cannot write " << out << std::endl; return 1; }
        f.write(reinterpret_cast<const char*>(knn.data()), (std::streamsize)(knn.size() * sizeof(int32_t)));
    }
    {
        std::ofstream meta(out + ".meta");
        meta << "dataset: " << dataset << std::endl;
        meta << "n: " << n << std::endl;
        meta << "dim: " << dim << std::endl;
        meta << "k: " << k << std::endl;
        meta << "method: " << (approximate ? "ivf_approximate" : "flat_exact") << std::endl;
        if (approximate) {
            meta << "n_clusters: " << n_list << std::endl;
            meta << "nprobe: " << actual_nprobe << std::endl;
            meta << "probe_ratio: " << (100.0 * actual_nprobe / n_list) << "%" << std::endl;
        }
        meta << "read_time: " << read_time << "s" << std::endl;
        meta << "build_time: " << build_time << "s" << std::endl;
        meta << "search_time: " << search_time << "s" << std::endl;
        meta << "total_time: " << (read_time + build_time + search_time) << "s" << std::endl;
    }
    std::cout << std::endl << "=== Summary ===" << std::endl;
    std::cout << "Total time: " << (read_time + build_time + search_time) << "s" << std::endl;
    std::cout << "Output file: " << out << std::endl;
    std::cout << "To load in Python:" << std::endl;
    std::cout << "  knn_data = np.fromfile('" << out << "', dtype=np.int32).reshape(" << n << ", " << k << ")" << std::endl;
    return 0;
}
