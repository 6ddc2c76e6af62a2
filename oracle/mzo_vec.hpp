// oracle/mzo_vec.hpp — TEST INFRASTRUCTURE.  Growable byte vector handed across
// the oracle's C API (mzo_vec_*).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
namespace mzo {
struct Vec {
  uint32_t row_bytes;
  std::vector<unsigned char> bytes;
};
template <class R>
inline void vec_append(Vec* v, const std::vector<R>& rows) {
  size_t off = v->bytes.size();
  v->bytes.resize(off + rows.size() * sizeof(R));
  if (!rows.empty()) std::memcpy(v->bytes.data() + off, rows.data(), rows.size() * sizeof(R));
}
}  // namespace mzo
