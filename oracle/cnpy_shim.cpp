// TEST INFRASTRUCTURE (oracle/): not part of the product path.
//
// The reference's search.cpp (/root/reference/search.cpp:170-215) reads its artifacts
// through cnpy::npy_load, declared in the reference's vendored cnpy.h:73. The matching
// cnpy.cpp is NOT vendored in the reference and libcnpy is not installed in this image,
// so this file supplies that one function so that the UNMODIFIED reference search.cpp
// links here (oracle/Makefile -> oracle/_ref/search_ref). It parses .npy format v1/v2/v3
// headers ("\x93NUMPY", version, header length, python-dict literal with 'descr',
// 'fortran_order', 'shape') as published in the NumPy format spec (NEP 1).
#include <cnpy.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

std::string dict_value(const std::string& hdr, const std::string& key) {
    size_t p = hdr.find("'" + key + "'");
    if (p == std::string::npos) throw std::runtime_error("npy header: missing key " + key);
    p = hdr.find(':', p);
    if (p == std::string::npos) throw std::runtime_error("npy header: malformed at " + key);
    ++p;
    while (p < hdr.size() && hdr[p] == ' ') ++p;
    size_t e = p;
    if (hdr[p] == '(') {
        e = hdr.find(')', p);
        return hdr.substr(p, e - p + 1);
    }
    if (hdr[p] == '\'') {
        e = hdr.find('\'', p + 1);
        return hdr.substr(p + 1, e - p - 1);
    }
    while (e < hdr.size() && hdr[e] != ',' && hdr[e] != '}') ++e;
    return hdr.substr(p, e - p);
}

}  // namespace

namespace cnpy {

NpyArray npy_load(std::string fname) {
    FILE* fp = std::fopen(fname.c_str(), "rb");
    if (!fp) throw std::runtime_error("npy_load: Unable to open file " + fname);
    unsigned char pre[10];
    if (std::fread(pre, 1, 8, fp) != 8 || std::memcmp(pre, "\x93NUMPY", 6) != 0) {
        std::fclose(fp);
        throw std::runtime_error("npy_load: bad magic in " + fname);
    }
    const int major = pre[6];
    size_t hlen = 0;
    if (major == 1) {
        if (std::fread(pre, 1, 2, fp) != 2) { std::fclose(fp); throw std::runtime_error("npy_load: short header"); }
        hlen = size_t(pre[0]) | (size_t(pre[1]) << 8);
    } else {
        if (std::fread(pre, 1, 4, fp) != 4) { std::fclose(fp); throw std::runtime_error("npy_load: short header"); }
        hlen = size_t(pre[0]) | (size_t(pre[1]) << 8) | (size_t(pre[2]) << 16) | (size_t(pre[3]) << 24);
    }
    std::string hdr(hlen, '\0');
    if (std::fread(&hdr[0], 1, hlen, fp) != hlen) { std::fclose(fp); throw std::runtime_error("npy_load: short header"); }

    const std::string descr = dict_value(hdr, "descr");
    const std::string forder = dict_value(hdr, "fortran_order");
    const std::string shp = dict_value(hdr, "shape");
    if (descr.size() < 3) { std::fclose(fp); throw std::runtime_error("npy_load: bad descr"); }
    if (descr[0] == '>') { std::fclose(fp); throw std::runtime_error("npy_load: big-endian data unsupported"); }
    const size_t word_size = std::strtoul(descr.c_str() + 2, nullptr, 10);

    std::vector<size_t> shape;
    for (size_t i = 0; i < shp.size();) {
        if (shp[i] >= '0' && shp[i] <= '9') {
            char* end = nullptr;
            shape.push_back(std::strtoull(shp.c_str() + i, &end, 10));
            i = size_t(end - shp.c_str());
        } else {
            ++i;
        }
    }
    NpyArray arr(shape, word_size, forder.find("True") != std::string::npos);
    const size_t nread = std::fread(arr.data<char>(), 1, arr.num_bytes(), fp);
    std::fclose(fp);
    if (nread != arr.num_bytes()) throw std::runtime_error("npy_load: failed fread of " + fname);
    return arr;
}

}  // namespace cnpy
