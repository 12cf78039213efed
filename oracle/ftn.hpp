// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use it.
//
// Small helpers shared by the CPU restatements: a table-blob reader and 1-based,
// column-major ("Fortran order") array views, so that the restatements can index
// exactly like the Fortran they follow.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// Blob layout (little endian), written by climt_b200/tables.py:write_blob():
//   char magic[8] = "CB2TBL01"; int64 n;
//   n x { char name[56]; int64 ndim; int64 shape[6]; int64 offset; int64 count }
//   double data[]           (each array in Fortran / column-major element order)
struct BlobEntry {
  std::vector<int64_t> shape;
  const double* p = nullptr;
  int64_t count = 0;
};

class Blob {
 public:
  explicit Blob(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open table blob " + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    raw_.resize((size_t)sz);
    if (std::fread(raw_.data(), 1, (size_t)sz, f) != (size_t)sz) {
      std::fclose(f);
      throw std::runtime_error("short read on " + path);
    }
    std::fclose(f);
    if (sz < 16 || std::memcmp(raw_.data(), "CB2TBL01", 8) != 0)
      throw std::runtime_error("bad magic in " + path);
    int64_t n;
    std::memcpy(&n, raw_.data() + 8, 8);
    const size_t esz = 56 + 8 + 48 + 8 + 8;
    const char* data0 = raw_.data() + 16 + (size_t)n * esz;
    for (int64_t i = 0; i < n; ++i) {
      const char* e = raw_.data() + 16 + (size_t)i * esz;
      char name[57];
      std::memcpy(name, e, 56);
      name[56] = 0;
      int64_t ndim, shp[6], off, cnt;
      std::memcpy(&ndim, e + 56, 8);
      std::memcpy(shp, e + 64, 48);
      std::memcpy(&off, e + 112, 8);
      std::memcpy(&cnt, e + 120, 8);
      BlobEntry be;
      be.shape.assign(shp, shp + ndim);
      be.p = reinterpret_cast<const double*>(data0) + off;
      be.count = cnt;
      map_[name] = be;
    }
  }
  const BlobEntry& get(const std::string& k) const {
    auto it = map_.find(k);
    if (it == map_.end()) throw std::runtime_error("table blob has no entry " + k);
    return it->second;
  }
  bool has(const std::string& k) const { return map_.count(k) != 0; }

 private:
  std::vector<char> raw_;
  std::map<std::string, BlobEntry> map_;
};

// 1-based column-major owning arrays (optionally with explicit lower bounds).
struct A1 {
  std::vector<double> d;
  int lb = 1;
  A1() {}
  A1(int n, int lb_ = 1) : d((size_t)n, 0.0), lb(lb_) {}
  double& operator()(int i) { return d[(size_t)(i - lb)]; }
  const double& operator()(int i) const { return d[(size_t)(i - lb)]; }
};
struct A2 {
  std::vector<double> d;
  int n1 = 0, lb1 = 1, lb2 = 1;
  A2() {}
  A2(int n1_, int n2_, int lb1_ = 1, int lb2_ = 1) : d((size_t)n1_ * n2_, 0.0), n1(n1_), lb1(lb1_), lb2(lb2_) {}
  double& operator()(int i, int j) { return d[(size_t)(i - lb1) + (size_t)n1 * (j - lb2)]; }
  const double& operator()(int i, int j) const { return d[(size_t)(i - lb1) + (size_t)n1 * (j - lb2)]; }
};
struct A3 {
  std::vector<double> d;
  int n1 = 0, n2 = 0;
  A3() {}
  A3(int n1_, int n2_, int n3_) : d((size_t)n1_ * n2_ * n3_, 0.0), n1(n1_), n2(n2_) {}
  double& operator()(int i, int j, int k) { return d[(size_t)(i - 1) + (size_t)n1 * ((j - 1) + (size_t)n2 * (k - 1))]; }
  const double& operator()(int i, int j, int k) const {
    return d[(size_t)(i - 1) + (size_t)n1 * ((j - 1) + (size_t)n2 * (k - 1))];
  }
};

}  // namespace orc
