// climt_b200 -- bits shared by the engine translation units (host only).
#pragma once
#include <mutex>
#include <string>

namespace cb {
inline std::string g_error;
inline std::mutex g_mu;
inline void set_global_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_error = s;
}
constexpr int kBlock = 128;
}  // namespace cb
