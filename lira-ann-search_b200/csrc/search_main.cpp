// search: end-to-end LIRA query phase over the artifacts written by the reference's index.py, with the command line
// and the printout of the reference's search.cpp (argv :32-84, artifacts :302-338, sweep :413-549, stdout keys
// :542-547) and the whole batch answered on the GPU through liblira_b200 (lira_probe_search).
//
//   search --dataset <name> --data_path <root> --artifacts_dir <dir> --prefix <cfg.file_name> --k <K>
//          --metric <L2|inner_product> [--num_threads N] [--t_min v --t_max v --t_step v] [--dedup 0|1]
//
// --dedup 0 (default) keeps search.cpp's select-k-then-collapse-duplicates behaviour (:499-513); --dedup 1 removes
// ids stored in several probed partitions BEFORE selection (the recall definition of LIRA_smallscale.py:210-214).
// LibTorch is used only to read the TorchScript checkpoint `{prefix}_mlp_2_input.pt` into the 12 weight tensors.
#include <torch/script.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lira_b200.h"

namespace {

struct Args {
    std::string dataset, data_path = "/data/vector_datasets", artifacts_dir = ".", prefix, metric = "L2";
    int k = 10, num_threads = 32, dedup = 0;
    float t_min = 0.02f, t_max = 0.80f, t_step = 0.02f;
};

void usage() {
    std::cout << "Usage:\n  search --dataset <name> --data_path <path_to_datasets_root> --artifacts_dir <path_to_python_artifacts>\n"
              << "         --prefix <file_prefix_same_as_cfg.file_name> --k <K_eval> --metric <L2|inner_product>\n"
              << "         [--num_threads N] [--t_min v --t_max v --t_step v] [--dedup 0|1]\n";
}

bool parse(int argc, char** argv, Args& a) {
    for (int i = 1; i < argc; ++i) {
        const std::string f = argv[i];
        if (i + 1 >= argc) { std::cerr << "Unknown or incomplete arg: " << f << "\n"; return false; }
        const char* v = argv[++i];
        if (f == "--dataset") a.dataset = v;
        else if (f == "--data_path") a.data_path = v;
        else if (f == "--artifacts_dir") a.artifacts_dir = v;
        else if (f == "--prefix") a.prefix = v;
        else if (f == "--k") a.k = std::stoi(v);
        else if (f == "--metric") a.metric = v;
        else if (f == "--num_threads") a.num_threads = std::stoi(v);
        else if (f == "--t_min") a.t_min = std::stof(v);
        else if (f == "--t_max") a.t_max = std::stof(v);
        else if (f == "--t_step") a.t_step = std::stof(v);
        else if (f == "--dedup") a.dedup = std::stoi(v);
        else { std::cerr << "Unknown or incomplete arg: " << f << "\n"; return false; }
    }
    if (a.dataset.empty() || a.prefix.empty()) { std::cerr << "Error: --dataset and --prefix are required.\n"; return false; }
    return true;
}

// ---- .npy (format 1.0-3.0, C order, little endian) ------------------------------------------------
struct Npy {
    std::string descr;
    std::vector<size_t> shape;
    std::vector<char> data;
};

Npy load_npy(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open " + path);
    unsigned char m[12];
    if (std::fread(m, 1, 8, f) != 8 || std::memcmp(m, "\x93NUMPY", 6) != 0) { std::fclose(f); throw std::runtime_error("Not an .npy file: " + path); }
    size_t hlen = 0;
    if (m[6] == 1) { if (std::fread(m, 1, 2, f) != 2) throw std::runtime_error("Short .npy header: " + path); hlen = m[0] | (m[1] << 8); }
    else { if (std::fread(m, 1, 4, f) != 4) throw std::runtime_error("Short .npy header: " + path); hlen = m[0] | (m[1] << 8) | (m[2] << 16) | ((size_t)m[3] << 24); }
    std::string h(hlen, ' ');
    if (std::fread(&h[0], 1, hlen, f) != hlen) { std::fclose(f); throw std::runtime_error("Short .npy header: " + path); }
    Npy a;
    auto field = [&](const char* key) {
        const size_t p = h.find(std::string("'") + key + "'");
        if (p == std::string::npos) throw std::runtime_error(std::string(".npy header lacks ") + key + ": " + path);
        return h.find(':', p) + 1;
    };
    { size_t p = h.find('\'', field("descr")); a.descr = h.substr(p + 1, h.find('\'', p + 1) - p - 1); }
    if (h.compare(h.find_first_not_of(' ', field("fortran_order")), 4, "True") == 0) throw std::runtime_error("Fortran-ordered .npy not supported: " + path);
    {
        size_t p = h.find('(', field("shape")), e = h.find(')', p);
        std::string t = h.substr(p + 1, e - p - 1);
        size_t i = 0;
        while (i < t.size()) {
            while (i < t.size() && (t[i] == ' ' || t[i] == ',')) ++i;
            if (i >= t.size()) break;
            a.shape.push_back(std::stoull(t.substr(i)));
            while (i < t.size() && t[i] != ',') ++i;
        }
    }
    size_t count = 1;
    for (size_t s : a.shape) count *= s;
    const size_t isz = (size_t)std::stoi(a.descr.substr(2));
    a.data.resize(count * isz);
    if (std::fread(a.data.data(), 1, a.data.size(), f) != a.data.size()) { std::fclose(f); throw std::runtime_error("Truncated .npy: " + path); }
    std::fclose(f);
    return a;
}

std::vector<float> as_f32(const Npy& a, const std::string& what) {
    const size_t n = a.data.size() / (size_t)std::stoi(a.descr.substr(2));
    std::vector<float> out(n);
    if (a.descr == "<f4") std::memcpy(out.data(), a.data.data(), n * 4);
    else if (a.descr == "<f8") for (size_t i = 0; i < n; ++i) out[i] = (float)reinterpret_cast<const double*>(a.data.data())[i];
    else throw std::runtime_error(what + ": expected a float array, got dtype " + a.descr);
    return out;
}
std::vector<int32_t> as_i32(const Npy& a, const std::string& what) {
    const size_t n = a.data.size() / (size_t)std::stoi(a.descr.substr(2));
    std::vector<int32_t> out(n);
    if (a.descr == "<i4") std::memcpy(out.data(), a.data.data(), n * 4);
    else if (a.descr == "<i8") for (size_t i = 0; i < n; ++i) out[i] = (int32_t)reinterpret_cast<const int64_t*>(a.data.data())[i];
    else throw std::runtime_error(what + ": expected an integer array, got dtype " + a.descr);
    return out;
}

// ---- .fvecs / .ivecs ------------------------------------------------------------------------------
template <class T>
std::vector<T> read_xvecs(const std::string& path, size_t& n, size_t& d) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open " + path);
    int32_t dim = 0;
    if (std::fread(&dim, 4, 1, f) != 1 || dim <= 0) { std::fclose(f); throw std::runtime_error("Bad vecs header: " + path); }
    std::fseek(f, 0, SEEK_END);
    const long long bytes = std::ftell(f), rec = 4 + 4ll * dim;
    if (bytes % rec) { std::fclose(f); throw std::runtime_error("Bad vecs size: " + path); }
    n = (size_t)(bytes / rec);
    d = (size_t)dim;
    std::vector<T> out(n * d);
    std::fseek(f, 0, SEEK_SET);
    for (size_t i = 0; i < n; ++i) {
        int32_t di;
        if (std::fread(&di, 4, 1, f) != 1 || di != dim || std::fread(out.data() + i * d, 4, d, f) != d) { std::fclose(f); throw std::runtime_error("Truncated vecs file: " + path); }
    }
    std::fclose(f);
    return out;
}

}  // namespace

int main(int argc, char** argv) {
    Args args;
    if (!parse(argc, argv, args)) { usage(); return 1; }
    try {
        std::string base = args.artifacts_dir;
        if (!base.empty() && base.back() != '/') base += "/";
        const std::string prefix = base + args.prefix;
        std::cout << "Dataset      : " << args.dataset << "\nArtifacts dir: " << base << "\nPrefix       : " << prefix
                  << "\nMetric       : " << args.metric << "\nK            : " << args.k << "\n";

        // ---- artifacts of index.py (search.cpp:302-329) ----
        const Npy cen = load_npy(prefix + "_centroids.npy"), d2b = load_npy(prefix + "_data_2_bkt.npy"), xd = load_npy(prefix + "_x_d.npy");
        if (cen.shape.size() != 2 || d2b.shape.size() != 2 || xd.shape.size() != 2) throw std::runtime_error("centroids / data_2_bkt / x_d must be 2-D");
        const size_t B = cen.shape[0], dim = cen.shape[1], N = d2b.shape[0], n_mul = d2b.shape[1];
        if (xd.shape[0] != N) throw std::runtime_error("x_d.npy and data_2_bkt.npy mismatch in N.");
        if (xd.shape[1] != dim) throw std::runtime_error("centroids dim and x_d dim mismatch.");
        const std::vector<float> centroids = as_f32(cen, "centroids"), x_d = as_f32(xd, "x_d");
        const std::vector<int32_t> data_2_bkt = as_i32(d2b, "data_2_bkt");
        const std::vector<float> mean = as_f32(load_npy(prefix + "_scaler_mean.npy"), "scaler_mean"),
                                 scale = as_f32(load_npy(prefix + "_scaler_scale.npy"), "scaler_scale");
        if (mean.size() != B || scale.size() != B) throw std::runtime_error("Scaler length must equal n_bkt.");
        std::cout << "Loaded centroids: " << B << " x " << dim << "\nLoaded data_2_bkt: " << N << " x " << n_mul << "\nLoaded x_d: " << N
                  << " x " << dim << "\nLoaded scaler parameters, len = " << B << "\n";

        // ---- the probing model's weights out of the TorchScript checkpoint (search.cpp:331-338) ----
        const std::string model_path = prefix + "_mlp_2_input.pt";
        std::cout << "Loading TorchScript model from: " << model_path << "\n";
        torch::jit::script::Module module = torch::jit::load(model_path, torch::kCPU);
        const char* names[12] = {"distance_net.0.weight", "distance_net.0.bias", "distance_net.2.weight", "distance_net.2.bias",
                                 "vector_net.0.weight",   "vector_net.0.bias",   "vector_net.2.weight",   "vector_net.2.bias",
                                 "fc.0.weight",           "fc.0.bias",           "fc.2.weight",           "fc.2.bias"};
        std::vector<torch::Tensor> keep(12);
        const float* weights[12];
        for (int i = 0; i < 12; ++i) {
            bool found = false;
            for (const auto& p : module.named_parameters())
                if (p.name == names[i]) { keep[i] = p.value.detach().to(torch::kFloat32).contiguous().cpu(); found = true; }
            if (!found) throw std::runtime_error(std::string("TorchScript model lacks parameter ") + names[i]);
            weights[i] = keep[i].data_ptr<float>();
        }

        // ---- queries + ground truth (search.cpp:340-366) ----
        const std::string ds_dir = args.data_path + "/" + args.dataset;
        size_t n_q, d_q, n_gt, d_gt;
        const std::vector<float> x_q = read_xvecs<float>(ds_dir + "/" + args.dataset + "_query.fvecs", n_q, d_q);
        const std::vector<int32_t> gt = read_xvecs<int32_t>(ds_dir + "/" + args.dataset + "_groundtruth.ivecs", n_gt, d_gt);
        if (d_q != dim) throw std::runtime_error("query dim and x_d dim mismatch.");
        if (n_gt != n_q) throw std::runtime_error("groundtruth and query count mismatch.");
        if ((size_t)args.k > d_gt) throw std::runtime_error("k exceeds the groundtruth depth.");
        std::cout << "Loaded queries: " << n_q << " x " << d_q << "\nLoaded groundtruth: " << n_gt << " x " << d_gt << "\n";

        // ---- device-resident index (buckets of search.cpp:368-403) and model ----
        const int metric = (args.metric == "inner_product" || args.metric == "ip" || args.metric == "IP") ? LIRA_METRIC_IP : LIRA_METRIC_L2;
        lira_index_t* index = nullptr;
        lira_model_t* model = nullptr;
        if (lira_index_create_from_assign(x_d.data(), (int64_t)N, (int)dim, data_2_bkt.data(), (int)n_mul, (int)B, metric, 0, &index) ||
            lira_model_create(centroids.data(), mean.data(), scale.data(), (int)B, (int)dim, weights, 0, &model))
            throw std::runtime_error(lira_last_error());

        std::cout << "Start end-to-end search (outer loop = threshold, one GPU batch of " << n_q << " queries per threshold)\n";
        std::cout << "Threshold range: [" << args.t_min << ", " << args.t_max << "] step " << args.t_step << "\n\n";
        std::vector<float> D(n_q * args.k);
        std::vector<int64_t> I(n_q * args.k), cmp(n_q);
        std::vector<int32_t> nprobe(n_q);
        for (float thr = args.t_min; thr <= args.t_max + 1e-6f; thr += args.t_step) {   // float accumulation as search.cpp:413
            std::cout << "=== Threshold = " << thr << " ===\n";
            const auto t0 = std::chrono::high_resolution_clock::now();
            if (lira_probe_search(index, model, x_q.data(), (int64_t)n_q, LIRA_SELECT_GE_ARGMAX, (double)thr, args.k, args.dedup, D.data(),
                                  I.data(), nprobe.data(), cmp.data()))
                throw std::runtime_error(lira_last_error());
            const double total_time = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
            double sum_recall = 0, sum_nprobe = 0, sum_cmp = 0;
            for (size_t q = 0; q < n_q; ++q) {   // recall@k against gt[:k] (search.cpp:520-528)
                int hit = 0;
                for (int j = 0; j < args.k; ++j) {
                    const int64_t g = gt[q * d_gt + j];
                    for (int e = 0; e < args.k; ++e)
                        if (I[q * args.k + e] == g) { ++hit; break; }
                }
                sum_recall += (double)hit / args.k;
                sum_nprobe += nprobe[q];
                sum_cmp += (double)cmp[q];
            }
            std::cout << "Threshold    : " << thr << "\n";
            std::cout << "avg_recall   : " << sum_recall / n_q << "\n";
            std::cout << "avg_nprobe   : " << sum_nprobe / n_q << "\n";
            std::cout << "avg_cmp      : " << sum_cmp / n_q << "\n";
            std::cout << "avg_time(q)  : " << total_time / n_q << " s\n";
            std::cout << "QPS          : " << n_q / total_time << " q/s\n";
            std::cout << "----------------------------------------\n";
        }
        lira_model_free(model);
        lira_index_free(index);
        std::cout << "Done.\n";
    } catch (const std::exception& e) {
        std::cerr << "[Error] " << e.what() << "\n";
        return 1;
    }
    return 0;
}
