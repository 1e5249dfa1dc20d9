// oracle/common.hpp — shared helpers of the CPU restatement (TEST INFRASTRUCTURE, see oracle.h).
// chunk.Column accessors follow pkg/util/chunk/column.go; every function cites the Go source.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>
#include "oracle.h"

namespace orc {

void set_error(const std::string& msg);

// ---- MySQL type codes not listed in tidbgpu.h (pkg/parser/mysql/type.go:17-48) -----------------
enum : int {
  TypeUnspecified = 0, TypeTiny = 1, TypeShort = 2, TypeLong = 3, TypeFloat = 4, TypeDouble = 5,
  TypeNull = 6, TypeTimestamp = 7, TypeLonglong = 8, TypeInt24 = 9, TypeDate = 10, TypeDuration = 11,
  TypeDatetime = 12, TypeYear = 13, TypeNewDate = 14, TypeVarchar = 15, TypeBit = 16,
  TypeJSON = 0xf5, TypeNewDecimal = 0xf6, TypeEnum = 0xf7, TypeSet = 0xf8, TypeTinyBlob = 0xf9,
  TypeMediumBlob = 0xfa, TypeLongBlob = 0xfb, TypeBlob = 0xfc, TypeVarString = 0xfd, TypeString = 0xfe,
  TypeGeometry = 0xff, TypeTiDBVectorFloat32 = 0xe1
};
// pkg/parser/mysql/type.go:54-77
enum : uint32_t { NotNullFlag = 1u << 0, UnsignedFlag = 1u << 5, EnumSetAsIntFlag = 1u << 21 };

struct FieldType {
  int tp = TypeLonglong;
  uint32_t flag = 0;
  bool binary_coll = false;  // collate.CanUseRawMemAsKey(collator) for string types
};

constexpr int VarElemLen = -1;

// pkg/util/chunk/codec.go:165-179 getFixedLen
inline int fixed_len(int tp) {
  switch (tp) {
    case TypeFloat: return 4;
    case TypeTiny: case TypeShort: case TypeInt24: case TypeLong: case TypeLonglong:
    case TypeDouble: case TypeYear: case TypeDuration: return 8;
    case TypeDate: case TypeDatetime: case TypeTimestamp: return 8;   // sizeTime (CoreTime uint64)
    case TypeNewDecimal: return 40;                                    // types.MyDecimalStructSize
    default: return VarElemLen;
  }
}

// ---- chunk.Column accessors (pkg/util/chunk/column.go) -------------------------------------------
// Column.IsNull column.go:225 — bit 1 means NOT NULL, LSB-first
inline bool col_is_null(const tg_column& c, int64_t row) {
  if (!c.null_bitmap) return false;
  return ((c.null_bitmap[row >> 3] >> (row & 7)) & 1) == 0;
}
inline const uint8_t* col_raw(const tg_column& c, int64_t row) {  // Column.GetRaw (fixed width)
  return c.data + row * (int64_t)c.elem_len;
}
inline int64_t col_i64(const tg_column& c, int64_t row) {
  int64_t v; std::memcpy(&v, c.data + row * 8, 8); return v;
}
inline double col_f64(const tg_column& c, int64_t row) {
  double v; std::memcpy(&v, c.data + row * 8, 8); return v;
}
inline float col_f32(const tg_column& c, int64_t row) {
  float v; std::memcpy(&v, c.data + row * 4, 4); return v;
}
// Chunk.NumRows chunk.go:384 (logical rows) and physical rows of column 0
inline int64_t chunk_logical_rows(const tg_chunk& c) {
  if (c.sel) return c.nsel;
  return c.ncols > 0 ? c.cols[0].length : 0;
}
inline int64_t chunk_physical_rows(const tg_chunk& c) { return c.ncols > 0 ? c.cols[0].length : 0; }

// ---- owned output column: chunk.Column with append (column.go:288 AppendCellNTimes, :339
// AppendNNulls, appendNullBitmap :240) -----------------------------------------------------------
struct OColumn {
  int elem_len = 8;
  int64_t length = 0;
  std::vector<uint8_t> null_bitmap;   // bit 1 = NOT NULL
  std::vector<uint8_t> data;
  explicit OColumn(int el = 8) : elem_len(el) {}
  void append_bit(bool not_null) {
    if ((length & 7) == 0) null_bitmap.push_back(0);
    if (not_null) null_bitmap[length >> 3] |= (uint8_t)(1u << (length & 7));
  }
  void append_raw(const uint8_t* p) {
    append_bit(true);
    data.insert(data.end(), p, p + elem_len);
    length++;
  }
  void append_null() {
    append_bit(false);
    data.insert(data.end(), (size_t)elem_len, 0);
    length++;
  }
  void append_i64(int64_t v) { append_raw(reinterpret_cast<const uint8_t*>(&v)); }
  void append_f64(double v) { append_raw(reinterpret_cast<const uint8_t*>(&v)); }
  bool is_null(int64_t row) const { return ((null_bitmap[row >> 3] >> (row & 7)) & 1) == 0; }
};

int fetch_result(const std::vector<std::vector<OColumn>*>& parts, int ncols, tg_mut_chunk* out);

// ---- filters: expression.VectorizedFilter over a CNF of tg_filter_item -------------------------
// (pkg/expression/chunk_executor.go:413, VecEvalBool expression.go:409-494: a row is selected iff
// every CNF item evaluates to non-NULL true)
bool filter_row(const tg_chunk& chk, int64_t phys_row, const tg_filter_item* items, int n_items);

int compare_int(int64_t a, bool ua, int64_t b, bool ub);   // types.CompareInt types/compare.go:86
int compare_real(double a, double b);                      // Go cmp.Compare on float64

}  // namespace orc
