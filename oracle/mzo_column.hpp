// oracle/mzo_column.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the columnar wire format of SURVEY 8(f)-4:
//   * `Column<C>` = Typed | Bytes | Align (src/timely-util/src/columnar.rs:54-222): the
//     serialized form is `columnar::bytes::indexed::{encode, decode, length_in_words}` over the
//     container's `as_bytes()` slices.  The `columnar` crate (0.12.1, Cargo.lock:1954-1979) is NOT
//     vendored; its published layout is restated here and PINNED by the reference-held bytes of
//     `raw_columnar_bytes()` (columnar.rs:247-258: [16, 28, 1i32, 2i32, 3i32, pad]) and the three
//     tests around it (clone / known_bytes / from_bytes, columnar.rs:260-339):
//       word 0            = 8 * (k + 1): where the index of k + 1 byte offsets ends
//       word 1 + i        = end of slice i in bytes (slice i STARTS at the previous end rounded up
//                           to a multiple of 8)
//       then the k slices, each zero padded to whole words
//   * the containers whose slices we need:
//       ((u64, u64), u64, i64)              -> 4 slices: keys, vals, times, diffs (tuples chain
//                                              their fields' slices in order)
//       (u64, i64)                          -> 2 slices
//       ((Row, Row), Timestamp, Diff)       -> 6 slices: key bounds (u64 END offsets, one per row),
//                                              key bytes, val bounds, val bytes, times, diffs
//                                              (`Rows`, src/repr/src/row.rs:447-452,546-560,606-611;
//                                              `Timestamps`, src/repr/src/timestamp.rs:163-181;
//                                              `Overflows`, src/ore/src/overflowing.rs:127-200)
//   * the ship heuristic (`at_serialized_capacity`, columnar.rs:150-175) and `ColumnBuilder`
//     (src/timely-util/src/columnar/builder.rs:28-111): after every push, if the serialized size is
//     within 10 % of the next multiple of 2 MiB (2^18 words) the current container is minted as
//     `Column::Align` and cleared; `finish` hands out what remains.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../include/mzgpu.h"

namespace mzo {

struct ColSlice {
  const uint8_t* p;
  size_t len;  // bytes
};

// indexed::length_in_words: 1 + sum over slices of (1 + ceil(len / 8))
inline size_t col_length_in_words(const std::vector<ColSlice>& s) {
  size_t w = 1;
  for (const ColSlice& x : s) w += 1 + (x.len + 7) / 8;
  return w;
}

// indexed::encode
inline void col_encode(std::vector<uint64_t>& store, const std::vector<ColSlice>& s) {
  const uint64_t offsets_end = 8 * (uint64_t)(s.size() + 1);
  store.push_back(offsets_end);
  uint64_t pos = offsets_end;
  for (const ColSlice& x : s) {
    store.push_back(pos + x.len);       // the unpadded end ...
    pos += (x.len + 7) & ~(uint64_t)7;  // ... the next slice starts word aligned
  }
  for (const ColSlice& x : s) {
    const size_t words = (x.len + 7) / 8, at = store.size();
    store.resize(at + words, 0);  // zero padding
    if (x.len) std::memcpy(store.data() + at, x.p, x.len);
  }
}

// indexed::decode; false if the index is not one this format can have produced
inline bool col_decode(const uint64_t* store, size_t n_words, std::vector<ColSlice>* out) {
  out->clear();
  if (n_words == 0 || store[0] % 8 != 0 || store[0] < 8 || store[0] / 8 > n_words) return false;
  const size_t k = store[0] / 8 - 1;
  const uint64_t last = store[k];
  if (last > 8 * (uint64_t)n_words) return false;
  const uint8_t* bytes = (const uint8_t*)store;
  for (size_t i = 0; i < k; i++) {
    uint64_t upper = store[i + 1] < last ? store[i + 1] : last;
    uint64_t lower = (store[i] + 7) & ~(uint64_t)7;
    if (lower > upper) lower = upper;
    out->push_back(ColSlice{bytes + lower, (size_t)(upper - lower)});
  }
  return true;
}

// at_serialized_capacity (columnar.rs:164-175)
inline bool col_at_capacity(size_t words) {
  const size_t ship = (size_t)1 << 18;
  const size_t round = (words + (ship - 1)) & ~(ship - 1);
  return round - words < round / 10;
}

// A typed container of ((key, val), time, diff) updates in one of the three layouts.
struct ColumnTyped {
  int layout = MZGPU_COLUMN_U64X4;
  // fixed-width layouts
  std::vector<uint64_t> key, val, time;
  std::vector<int64_t> diff;
  // Row layout: `Rows` for keys and vals
  std::vector<uint64_t> kb, vb;  // bounds: END offset of every row
  std::vector<uint8_t> kv, vv;   // the rows' bytes back to back
  size_t len() const { return diff.size(); }
  void clear() {
    key.clear(), val.clear(), time.clear(), diff.clear(), kb.clear(), vb.clear(), kv.clear(), vv.clear();
  }
  std::vector<ColSlice> as_bytes() const {
    auto u = [](const std::vector<uint64_t>& v) { return ColSlice{(const uint8_t*)v.data(), v.size() * 8}; };
    ColSlice d{(const uint8_t*)diff.data(), diff.size() * 8};
    if (layout == MZGPU_COLUMN_U64X2) return {u(key), d};
    if (layout == MZGPU_COLUMN_U64X4) return {u(key), u(val), u(time), d};
    return {u(kb), ColSlice{kv.data(), kv.size()}, u(vb), ColSlice{vv.data(), vv.size()}, u(time), d};
  }
};

// Row bytes of a packed Row word (mzgpu_rowkey_pack: len << 56 | bytes big-endian, zero padded)
inline void unpack_row(uint64_t w, std::vector<uint8_t>& bytes, std::vector<uint64_t>& bounds) {
  const unsigned len = (unsigned)(w >> 56);
  for (unsigned i = 0; i < len; i++) bytes.push_back((uint8_t)(w >> (8 * (6 - i))));
  bounds.push_back(bytes.size());
}
inline bool pack_row(const uint8_t* p, size_t len, uint64_t* w) {
  if (len > 7) return false;
  uint64_t x = (uint64_t)len << 56;
  for (size_t i = 0; i < len; i++) x |= (uint64_t)p[i] << (8 * (6 - i));
  *w = x;
  return true;
}

inline void column_push(ColumnTyped& c, const mzgpu_r32& r) {
  if (c.layout == MZGPU_COLUMN_ROWROW) {
    unpack_row(r.key, c.kv, c.kb);
    unpack_row(r.val, c.vv, c.vb);
  } else {
    c.key.push_back(r.key);
    if (c.layout == MZGPU_COLUMN_U64X4) c.val.push_back(r.val);
  }
  if (c.layout != MZGPU_COLUMN_U64X2) c.time.push_back(r.time);
  c.diff.push_back(r.diff);
}

// Column::borrow().into_index_iter() over serialized words: the rows back, as R32 (layout U64X2
// leaves val = time = 0).  Returns 0 ok, -1 malformed, -2 a Row longer than 7 bytes.
inline int column_rows(int layout, const uint64_t* words, size_t n_words, std::vector<mzgpu_r32>* out) {
  std::vector<ColSlice> s;
  if (!col_decode(words, n_words, &s)) return -1;
  const size_t k = layout == MZGPU_COLUMN_U64X2 ? 2 : (layout == MZGPU_COLUMN_U64X4 ? 4 : 6);
  if (s.size() != k) return -1;
  const ColSlice &ds = s[k - 1];
  if (ds.len % 8) return -1;
  const size_t n = ds.len / 8;
  auto u64s = [&](const ColSlice& x) -> const uint64_t* { return (const uint64_t*)x.p; };
  if (layout != MZGPU_COLUMN_ROWROW) {
    for (size_t i = 0; i + 1 < k; i++)
      if (s[i].len != 8 * n) return -1;
    for (size_t i = 0; i < n; i++) {
      mzgpu_r32 r{};
      r.key = u64s(s[0])[i];
      if (layout == MZGPU_COLUMN_U64X4) r.val = u64s(s[1])[i], r.time = u64s(s[2])[i];
      r.diff = ((const int64_t*)ds.p)[i];
      out->push_back(r);
    }
    return 0;
  }
  if (s[0].len != 8 * n || s[2].len != 8 * n || s[4].len != 8 * n) return -1;
  uint64_t klo = 0, vlo = 0;
  for (size_t i = 0; i < n; i++) {
    const uint64_t khi = u64s(s[0])[i], vhi = u64s(s[2])[i];
    if (khi < klo || khi > s[1].len || vhi < vlo || vhi > s[3].len) return -1;
    mzgpu_r32 r{};
    if (!pack_row(s[1].p + klo, khi - klo, &r.key) || !pack_row(s[3].p + vlo, vhi - vlo, &r.val)) return -2;
    r.time = u64s(s[4])[i];
    r.diff = ((const int64_t*)ds.p)[i];
    out->push_back(r);
    klo = khi, vlo = vhi;
  }
  return 0;
}

// ColumnBuilder (builder.rs:28-111): containers minted by push_into, the rest by finish.
struct ColumnBuilder {
  ColumnTyped current;
  std::vector<std::vector<uint64_t>> pending;  // Column::Align allocations, in order
  explicit ColumnBuilder(int layout) { current.layout = layout; }
  void push(const mzgpu_r32& r) {
    column_push(current, r);
    const size_t words = col_length_in_words(current.as_bytes());
    if (col_at_capacity(words)) {
      std::vector<uint64_t> alloc;
      alloc.reserve(words);
      col_encode(alloc, current.as_bytes());
      pending.push_back(std::move(alloc));
      current.clear();
    }
  }
  // finish: the remainder (a Column::Typed in the reference; serialized here the way
  // `into_bytes` would put it on the wire)
  void finish() {
    if (current.len()) {
      std::vector<uint64_t> alloc;
      col_encode(alloc, current.as_bytes());
      pending.push_back(std::move(alloc));
      current.clear();
    }
  }
};

}  // namespace mzo
