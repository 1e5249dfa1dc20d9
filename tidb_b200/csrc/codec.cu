// codec.cu — the chunk wire format (host code only).
//
// What it replaces: chunk.Codec (pkg/util/chunk/codec.go) — Encode :41 / encodeColumn :49, DecodeToChunk :93 /
// decodeColumn :101, setAllNotNull :145, getFixedLen :165.  Coprocessor / TiFlash responses arrive in this format; the Go
// decoder makes every Column alias the gRPC buffer.  tg_chunk_decode does the same with tg_column views (zero copy), so a
// response can go from the wire straight into tg_join_probe_push / tg_agg_push; tg_chunk_decode_into additionally lands the
// fixed-width columns in caller-owned (pinned, tg_host_alloc) buffers, the form the H2D copy engine wants.
//
// Per column:  u32 length | u32 nullCount | [nullBitmap, (length+7)/8 bytes, only if nullCount > 0]
//              | [offsets, (length+1) x i64, only for var-len types] | data (fixedLen*length bytes, or offsets[length])
// all little endian; bitmap bit 1 = NOT NULL, LSB first.
#include "common.cuh"

namespace tg {

static int64_t null_count(const tg_column& c) {
  // Column.nullCount (column.go:243): zero bits among the first `length` bits
  if (!c.null_bitmap) return 0;
  int64_t n = c.length, zeros = 0;
  for (int64_t i = 0; i < n / 8; i++) zeros += 8 - __builtin_popcount(c.null_bitmap[i]);
  for (int64_t i = n & ~7ll; i < n; i++) zeros += !((c.null_bitmap[i >> 3] >> (i & 7)) & 1);
  return zeros;
}

static int column_wire_bytes(const tg_column& c, size_t* out) {
  if (c.length < 0 || (uint64_t)c.length > 0xFFFFFFFFull) return fail(TG_ERR_INVALID, "column length does not fit the u32 wire field");
  size_t b = 8;
  if (null_count(c) > 0) b += (size_t)((c.length + 7) / 8);
  if (c.elem_len < 0) {
    if (!c.offsets && c.length > 0) return fail(TG_ERR_INVALID, "var-len column without offsets");
    b += (size_t)(c.length + 1) * 8 + (size_t)(c.length > 0 ? c.offsets[c.length] : 0);
  } else b += (size_t)c.length * (size_t)c.elem_len;
  *out = b;
  return TG_OK;
}

}  // namespace tg

using namespace tg;

extern "C" {

int tg_chunk_wire_size(const tg_chunk* chk, size_t* bytes) {
  if (!chk || !bytes) return fail(TG_ERR_INVALID, "chk / bytes is NULL");
  if (chk->sel) return fail(TG_ERR_UNSUPPORTED, "Codec.Encode ignores sel; compact the chunk first");
  size_t total = 0;
  for (int c = 0; c < chk->ncols; c++) { size_t b = 0; TG_TRY(column_wire_bytes(chk->cols[c], &b)); total += b; }
  *bytes = total;
  return TG_OK;
}

int tg_chunk_encode(const tg_chunk* chk, uint8_t* buf, size_t cap, size_t* written) {
  size_t need = 0;
  TG_TRY(tg_chunk_wire_size(chk, &need));
  if (!buf || cap < need) return fail(TG_ERR_CAPACITY, "encode buffer too small (see tg_chunk_wire_size)");
  uint8_t* p = buf;
  for (int ci = 0; ci < chk->ncols; ci++) {
    const tg_column& c = chk->cols[ci];
    const uint32_t len = (uint32_t)c.length, nulls = (uint32_t)null_count(c);
    std::memcpy(p, &len, 4); std::memcpy(p + 4, &nulls, 4); p += 8;            // codec.go:52-58
    if (nulls > 0) {                                                          // :61-64
      const size_t nb = (size_t)((c.length + 7) / 8);
      std::memcpy(p, c.null_bitmap, nb);
      p += nb;
    }
    if (c.elem_len < 0) {                                                     // :67-71
      static const int64_t zero = 0;
      const size_t ob = (size_t)(c.length + 1) * 8;
      std::memcpy(p, c.length > 0 ? c.offsets : &zero, ob);
      p += ob;
      const size_t db = (size_t)(c.length > 0 ? c.offsets[c.length] : 0);
      if (db) std::memcpy(p, c.data, db);
      p += db;
    } else {
      const size_t db = (size_t)c.length * (size_t)c.elem_len;                 // :74
      if (db) std::memcpy(p, c.data, db);
      p += db;
    }
  }
  if (written) *written = (size_t)(p - buf);
  return TG_OK;
}

// zero-copy decode: cols_out[i] alias `buf` (like decodeColumn :101-140); null_bitmap = NULL means "no NULLs" — the
// reference materialises an all-ones bitmap there (setAllNotNull :145), tg_column's convention makes that unnecessary
int tg_chunk_decode(const uint8_t* buf, size_t len, int32_t ncols, const int32_t* mysql_types, tg_column* cols_out,
                    size_t* consumed) {
  if (!buf || !mysql_types || !cols_out || ncols < 0) return fail(TG_ERR_INVALID, "buf / mysql_types / cols_out is NULL");
  size_t pos = 0;
  for (int ci = 0; ci < ncols; ci++) {
    if (len - pos < 8) return fail(TG_ERR_INVALID, "truncated chunk: column header");
    uint32_t n = 0, nulls = 0;
    std::memcpy(&n, buf + pos, 4); std::memcpy(&nulls, buf + pos + 4, 4); pos += 8;
    tg_column& c = cols_out[ci];
    std::memset(&c, 0, sizeof(c));
    c.length = (int64_t)n;
    if (nulls > 0) {
      const size_t nb = ((size_t)n + 7) / 8;
      if (len - pos < nb) return fail(TG_ERR_INVALID, "truncated chunk: null bitmap");
      c.null_bitmap = buf + pos; pos += nb;
    }
    const int fl = fixed_len(mysql_types[ci]);
    c.elem_len = fl;
    size_t db;
    if (fl < 0) {
      const size_t ob = ((size_t)n + 1) * 8;
      if (len - pos < ob) return fail(TG_ERR_INVALID, "truncated chunk: offsets");
      // like bytesToI64Slice (codec.go:154) the view may be unaligned inside the message; read it with memcpy here
      c.offsets = reinterpret_cast<const int64_t*>(buf + pos);
      int64_t last = 0;
      std::memcpy(&last, buf + pos + (size_t)n * 8, 8);
      pos += ob;
      if (last < 0) return fail(TG_ERR_INVALID, "corrupt chunk: negative data length");
      db = (size_t)last;
    } else db = (size_t)n * (size_t)fl;
    if (len - pos < db) return fail(TG_ERR_INVALID, "truncated chunk: data");
    c.data = buf + pos; pos += db;
  }
  if (consumed) *consumed = pos;
  return TG_OK;
}

// decode + copy of the FIXED-WIDTH columns into caller-owned buffers (out->cols[i].data must hold capacity_rows elements,
// null_bitmap ceil(capacity_rows/8) bytes or NULL if the caller knows the column is NOT NULL); var-len columns are skipped
// (out->cols[i].data == NULL) or rejected.  The bitmap is always written (all ones when the wire carried none).
int tg_chunk_decode_into(const uint8_t* buf, size_t len, int32_t ncols, const int32_t* mysql_types, tg_mut_chunk* out,
                         int64_t* rows, size_t* consumed) {
  if (!out || out->ncols != ncols) return fail(TG_ERR_INVALID, "out chunk column count does not match");
  std::vector<tg_column> v((size_t)ncols);
  TG_TRY(tg_chunk_decode(buf, len, ncols, mysql_types, v.data(), consumed));
  int64_t n = ncols ? v[0].length : 0;
  for (int ci = 0; ci < ncols; ci++) {
    if (v[ci].length != n) return fail(TG_ERR_INVALID, "columns of one chunk have different lengths");
    tg_mut_column& o = out->cols[ci];
    if (!o.data) continue;
    if (v[ci].elem_len < 0) return fail(TG_ERR_UNSUPPORTED, "tg_chunk_decode_into copies fixed-width columns only");
    if (o.elem_len != v[ci].elem_len) return fail(TG_ERR_INVALID, "out column elem_len does not match the type");
    if (n > out->capacity_rows) return fail(TG_ERR_CAPACITY, "out chunk capacity_rows too small");
    if (n) std::memcpy(o.data, v[ci].data, (size_t)n * (size_t)o.elem_len);
    if (o.null_bitmap) {
      const size_t nb = (size_t)((n + 7) / 8);
      if (v[ci].null_bitmap) std::memcpy(o.null_bitmap, v[ci].null_bitmap, nb);
      else std::memset(o.null_bitmap, 0xFF, nb);
    } else if (v[ci].null_bitmap) return fail(TG_ERR_INVALID, "column carries NULLs but the out column has no null_bitmap");
  }
  if (rows) *rows = n;
  return TG_OK;
}

}  // extern "C"
