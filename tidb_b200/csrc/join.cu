// join.cu — tg_join_*: the GPU hash join behind HashJoinV2Exec's Open/Next/Close contract
// (pkg/executor/join/hash_join_v2.go:608, :690, :1161, :647).  Host-side orchestration only; the
// kernels live in join_kernels.cuh.
#include <memory>
#include <deque>
#include <algorithm>
#include <condition_variable>
#include "join_kernels.cuh"
#include "partition_kernels.cuh"

namespace tg {

static const int64_t kGeneralBatchRows = 16ll << 20;   // sub-batch of the general probe path (bounds temp memory)
static const int64_t kStageBatchRows = 4ll << 20;      // host staging batch for small pushed chunks
static const int64_t kDirectPushRows = 128ll << 10;    // chunks at least this big are copied straight from the caller
static const int64_t kNextWindowRows = 1ll << 20;      // D2H window that serves small tg_join_next calls

struct Side {
  int ncols = 0;
  std::vector<int> types;
  std::vector<uint32_t> flags;
  std::vector<int> elem;
  std::vector<char> needed;       // staged to the device (key ∪ used ∪ filter columns)
  int key_col = -1;               // the single equal-condition key; -1 when the join has several (key_cols)
  std::vector<int> key_cols;      // several equal conditions: the key columns in condition order
  std::vector<int> key_reject;    //   per key column: mixed-signedness pair and this is the signed side (negative = no key)
  std::vector<int> used;          // columns of this side that appear in the output
  DevFilter filter{};
};

// device-resident columns of one side for one batch
struct ColStore {
  std::vector<std::unique_ptr<DevBuf>> data, nulls;
  std::vector<char> has_nulls;
  int64_t rows = 0, cap_rows = 0;
  void init(int ncols) {
    data.clear(); nulls.clear();
    for (int i = 0; i < ncols; i++) { data.emplace_back(new DevBuf()); nulls.emplace_back(new DevBuf()); }
    has_nulls.assign(ncols, 0);
    rows = cap_rows = 0;
  }
  DevCols view(const Side& s) const {
    DevCols v{};
    for (int c = 0; c < s.ncols && c < TG_MAX_COLS; c++) {
      v.data[c] = s.needed[c] ? data[c]->p : nullptr;
      v.nulls[c] = (s.needed[c] && has_nulls[c]) ? nulls[c]->as<uint8_t>() : nullptr;
      v.elem_len[c] = s.elem[c];
    }
    return v;
  }
};

// host staging of pushed chunks (pinned), one buffer per needed column
struct HostStage {
  std::vector<std::unique_ptr<PinBuf>> data, nulls;
  std::vector<char> has_nulls;
  int64_t rows = 0;
  void init(int ncols) {
    data.clear(); nulls.clear();
    for (int i = 0; i < ncols; i++) { data.emplace_back(new PinBuf()); nulls.emplace_back(new PinBuf()); }
    has_nulls.assign(ncols, 0);
    rows = 0;
  }
  void reset() { rows = 0; std::fill(has_nulls.begin(), has_nulls.end(), 0); for (auto& d : data) d->used = 0; for (auto& d : nulls) d->used = 0; }
};

struct ResultBatch {
  std::vector<std::unique_ptr<DevBuf>> cols, bitmaps;
  int64_t rows = 0;
  int64_t consumed = 0;
};

}  // namespace tg

using namespace tg;

struct JoinImpl;

// The C-ABI handle is a small SHELL that outlives tg_join_close: close frees the implementation (GPU buffers, streams,
// staging) but parks the shell in a bounded graveyard (common.cuh), so a caller that raced with close — parked in
// tg_join_next_wait, blocked on `mu`, or about to enter with a pointer it read just before the close — finds a live
// `closed` flag and gets TG_ERR_CANCELLED instead of touching freed memory (exec.Executor: "Close may be called ...
// with Next() at the same time", executor.go:65).  A second close is a no-op.
struct tg_join {
  std::mutex mu;                     // build / probe-input side (one pushing thread)
  std::mutex res_mu;                 // result queue (one pulling thread may run concurrently with the pusher, like the
  std::condition_variable res_cv;    //   reference's probe fetcher vs joinResultCh consumer, hash_join_v2.go:840 / :1176)
  std::atomic<bool> closed{false};
  JoinImpl* impl = nullptr;          // guarded by mu + res_mu; nullptr once closed
};

struct JoinImpl {
  std::mutex& res_mu;                // the shell's (see tg_join)
  std::condition_variable& res_cv;
  explicit JoinImpl(tg_join& shell) : res_mu(shell.res_mu), res_cv(shell.res_cv) {}
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t d2h_stream = nullptr; // tg_join_next copies results on its own stream: D2H overlaps the next H2D + probe
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int nsm = 148;
  double load_factor = 0.5;
  bool default_load_factor = true;

  int join_type = 0;
  bool build_is_right = true;
  Side build, probe;
  int n_lused = 0, n_rused = 0;
  std::vector<int> lused, rused;
  int probe_kind = PK_INNER;
  bool need_scan = false;          // JoinProbe.NeedScanRowTable
  int scan_mode = 0;
  bool has_flag_col = false;
  int n_out = 0;
  std::vector<int> out_elem;
  KeySpec build_key{}, probe_key{};   // data pointers filled per launch
  bool multi_key = false;             // several equal conditions: synthetic 64-bit candidate key + residual equalities (k_composite_key)
  DevBuf bkey_syn, bkey_syn_nn, pkey_syn, pkey_syn_nn;
  std::vector<tg_other_item> other;   // OtherCondition (sides already mapped: 0 = probe child, 1 = build child)
  std::vector<int> other_build_cols;  // build columns it reads (kept in the row store although they may not be output)
  DevOther dev_other{};               // compiled after the build (row-store word of every build operand)

  // build state
  HostStage bstage;
  ColStore bcols;
  bool built = false;
  DevBuf table, rows_store, row_slot, row_rank, slot_used, scalars;
  TableView tv{};
  RowSpec rowspec{};
  std::vector<int> build_word_of_col;   // build column → row-store word (mode G)
  int u1_payload_col = -1;

  // probe state
  HostStage pstage;
  ColStore pcols_dev;
  DevBuf tmp_cnt, tmp_slot, tmp_off, tmp_sums, out_cursor;
  DevBuf part_scratch;                               // counts | cursors | offsets of the L2 partition pass
  std::unique_ptr<DevBuf> part_cols[1 + TG_FAST_MAX_PCOLS];   // partitioned copies of the probe key and payload columns
  std::vector<std::unique_ptr<DevBuf>> tmp_valid;
  std::deque<std::unique_ptr<ResultBatch>> results;
  std::vector<std::unique_ptr<ResultBatch>> free_batches;   // recycled (cudaFree would synchronise the whole device)
  std::unique_ptr<ResultBatch> dev_result;      // tg_join_probe_dev output (reused across calls)
  std::atomic<bool> probe_finished{false};
  std::atomic<int64_t> d2h_bytes{0};
  // small-Next window
  PinBuf win;
  int64_t win_lo = 0, win_hi = 0;
  ResultBatch* win_batch = nullptr;

  tg_join_stats stats{};
};

namespace tg {

static int env_int(const char* name, int dflt) { const char* v = getenv(name); return (v && *v) ? atoi(v) : dflt; }

static int grid_for(const JoinImpl* j, int64_t n, int block, int per_sm) {
  int64_t need = (n + block - 1) / block;
  int64_t cap = (int64_t)j->nsm * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- descriptor → handle -------------------------------------------------------------------------------
static int fill_side(Side& s, int n, const int32_t* types, const uint32_t* flags) {
  if (n <= 0 || n > TG_MAX_COLS) return fail(TG_ERR_UNSUPPORTED, "child schema must have 1..16 columns");
  s.ncols = n;
  s.types.assign(types, types + n);
  s.flags.resize(n);
  for (int i = 0; i < n; i++) s.flags[i] = flags ? flags[i] : 0;
  s.elem.resize(n);
  for (int i = 0; i < n; i++) s.elem[i] = fixed_len(types[i]);
  s.needed.assign(n, 0);
  return TG_OK;
}

static int key_kind_of(int tp) {
  if (is_int_family(tp)) return KEY_I64;
  if (tp == TG_TYPE_DOUBLE) return KEY_F64;
  if (tp == TG_TYPE_FLOAT) return KEY_F32;
  if (tp == TG_TYPE_DATE || tp == TG_TYPE_DATETIME || tp == TG_TYPE_TIMESTAMP) return KEY_TIME;   // getKeyProp join_table_meta.go:154
  return -1;
}
static bool key_unsigned(int tp, uint32_t flag) {
  // getKeyProp join_table_meta.go:130: YEAR always unsigned, DURATION always signed
  if (tp == TG_TYPE_YEAR) return true;
  if (tp == TG_TYPE_DURATION) return false;
  return (flag & TG_FLAG_UNSIGNED) != 0;
}

static int check_filter(const Side& s, const tg_filter_item* items, int n, DevFilter& out) {
  if (n < 0 || n > TG_MAX_FILTER) return fail(TG_ERR_UNSUPPORTED, "at most 8 CNF filter items are offloaded");
  out.n = n;
  if (n > 0 && !items) return fail(TG_ERR_INVALID, "filter items are NULL");
  for (int i = 0; i < n; i++) {
    const tg_filter_item& it = items[i];
    if (it.lhs_col < 0 || it.lhs_col >= s.ncols || it.rhs_col >= s.ncols) return fail(TG_ERR_INVALID, "filter column out of range");
    // the compare family must fit the column types: a real compare on integer columns (or the reverse) would reinterpret bits
    const bool lreal = s.types[it.lhs_col] == TG_TYPE_DOUBLE;
    if ((it.is_real != 0) != lreal || (it.rhs_col >= 0 && (s.types[it.rhs_col] == TG_TYPE_DOUBLE) != lreal))
      return fail(TG_ERR_UNSUPPORTED, "filter compares columns of different families (the planner casts before the filter)");
    if (s.elem[it.lhs_col] != 8 || (it.rhs_col >= 0 && s.elem[it.rhs_col] != 8))
      return fail(TG_ERR_UNSUPPORTED, "filters are offloaded on 8-byte columns only");
    if (it.op < TG_CMP_LT || it.op > TG_CMP_NE) return fail(TG_ERR_INVALID, "bad filter op");
    out.items[i] = it;
  }
  return TG_OK;
}

static int setup(JoinImpl* j, const tg_join_desc* d) {
  if (!d) return fail(TG_ERR_INVALID, "desc is NULL");
  j->join_type = d->join_type;
  j->build_is_right = d->build_is_right != 0;
  Side left, right;
  TG_TRY(fill_side(left, d->n_left_cols, d->left_types, d->left_flags));
  TG_TRY(fill_side(right, d->n_right_cols, d->right_types, d->right_flags));
  if (d->nkeys < 1 || d->nkeys > TG_MAX_JOIN_KEYS) return fail(TG_ERR_UNSUPPORTED, "GPU hash join handles 1..4 equal-condition keys");
  if (!d->left_key_idx || !d->right_key_idx) return fail(TG_ERR_INVALID, "key index arrays are NULL");
  j->multi_key = d->nkeys > 1;
  std::vector<tg_other_item> residual;   // several keys: `left_key_i = right_key_i`, re-checked on every candidate pair
  if (!j->multi_key) {
    int lk = d->left_key_idx[0], rk = d->right_key_idx[0];
    if (lk < 0 || lk >= left.ncols || rk < 0 || rk >= right.ncols) return fail(TG_ERR_INVALID, "key column out of range");
    left.key_col = lk; right.key_col = rk;
    int lkind = key_kind_of(left.types[lk]), rkind = key_kind_of(right.types[rk]);
    if (lkind < 0 || rkind < 0) return fail(TG_ERR_UNSUPPORTED, "join key type is not offloaded (int family / float / double / date-time only)");
    if ((lkind == KEY_I64) != (rkind == KEY_I64)) return fail(TG_ERR_UNSUPPORTED, "integer vs real join keys are cast by the planner before the join");
    if ((lkind == KEY_TIME) != (rkind == KEY_TIME)) return fail(TG_ERR_UNSUPPORTED, "date-time keys only join date-time keys (codec.go:707)");
  } else {
    // FixedSerializedKey mode (join_table_meta.go:174-178) restricted to 8-byte integer-family columns
    for (int q = 0; q < d->nkeys; q++) {
      int lk = d->left_key_idx[q], rk = d->right_key_idx[q];
      if (lk < 0 || lk >= left.ncols || rk < 0 || rk >= right.ncols) return fail(TG_ERR_INVALID, "key column out of range");
      if (!is_int_family(left.types[lk]) || !is_int_family(right.types[rk]) || left.elem[lk] != 8 || right.elem[rk] != 8)
        return fail(TG_ERR_UNSUPPORTED, "joins on several key columns are offloaded for 8-byte integer-family keys only");
      const bool lu = key_unsigned(left.types[lk], left.flags[lk]), ru = key_unsigned(right.types[rk], right.flags[rk]);
      left.key_cols.push_back(lk); right.key_cols.push_back(rk);
      left.key_reject.push_back(lu != ru && !lu); right.key_reject.push_back(lu != ru && !ru);
      tg_other_item it{};
      it.op = TG_CMP_EQ; it.is_real = 0; it.lhs_side = 0; it.lhs_col = lk; it.rhs_side = 1; it.rhs_col = rk;
      it.lhs_unsigned = lu; it.rhs_unsigned = ru;
      residual.push_back(it);
    }
  }
  auto used_list = [&](int n, const int32_t* v, int ncols, std::vector<int>& out) -> int {
    out.clear();
    if (n < 0) { for (int i = 0; i < ncols; i++) out.push_back(i); return TG_OK; }
    for (int i = 0; i < n; i++) { if (v[i] < 0 || v[i] >= ncols) return fail(TG_ERR_INVALID, "used column out of range"); out.push_back(v[i]); }
    return TG_OK;
  };
  TG_TRY(used_list(d->n_lused, d->lused, left.ncols, j->lused));
  TG_TRY(used_list(d->n_rused, d->rused, right.ncols, j->rused));
  left.used = j->lused; right.used = j->rused;
  j->has_flag_col = false;
  // NewJoinProbe base_join_probe.go:850-932
  bool brt = j->build_is_right;
  switch (j->join_type) {
    case TG_JOIN_INNER: j->probe_kind = PK_INNER; j->need_scan = false; break;
    case TG_JOIN_LEFT_OUTER:
      j->need_scan = !brt; j->probe_kind = j->need_scan ? PK_INNER : PK_PROBE_OUTER; j->scan_mode = 0; break;
    case TG_JOIN_RIGHT_OUTER:
      j->need_scan = brt; j->probe_kind = j->need_scan ? PK_INNER : PK_PROBE_OUTER; j->scan_mode = 0; break;
    case TG_JOIN_SEMI: case TG_JOIN_ANTI_SEMI:
      if (!j->rused.empty()) return fail(TG_ERR_INVALID, "len(rUsed) != 0 for semi join");
      j->need_scan = !brt;
      if (brt) j->probe_kind = j->join_type == TG_JOIN_SEMI ? PK_SEMI : PK_ANTI;
      else { j->probe_kind = PK_MARK_ONLY; j->scan_mode = j->join_type == TG_JOIN_SEMI ? 1 : 0; }
      break;
    case TG_JOIN_LEFT_OUTER_SEMI: case TG_JOIN_ANTI_LEFT_OUTER_SEMI:
      if (!j->rused.empty()) return fail(TG_ERR_INVALID, "len(rUsed) != 0 for left outer semi join");
      if (!brt) return fail(TG_ERR_UNSUPPORTED, "left outer semi join needs the right side as build side");
      j->probe_kind = j->join_type == TG_JOIN_LEFT_OUTER_SEMI ? PK_LEFT_OUTER_SEMI : PK_ANTI_LEFT_OUTER_SEMI;
      j->need_scan = false; j->has_flag_col = true;
      break;
    default: return fail(TG_ERR_INVALID, "unknown join type");
  }
  j->build = brt ? right : left;
  j->probe = brt ? left : right;
  TG_TRY(check_filter(j->build, d->build_filter, d->n_build_filter, j->build.filter));
  TG_TRY(check_filter(j->probe, d->probe_filter, d->n_probe_filter, j->probe.filter));
  // OtherCondition (inner_join_probe.go:72-79): sides re-mapped to probe (0) / build (1)
  j->other.clear(); j->other_build_cols.clear();
  if (d->n_other_cond < 0 || d->n_other_cond + (int)residual.size() > TG_MAX_OTHER)
    return fail(TG_ERR_UNSUPPORTED, "at most 8 OtherCondition items (key equalities of a multi-column key included) are offloaded");
  if (d->n_other_cond > 0 && !d->other_cond) return fail(TG_ERR_INVALID, "other_cond is NULL");
  std::vector<tg_other_item> items = residual;
  for (int i = 0; i < d->n_other_cond; i++) items.push_back(d->other_cond[i]);
  if (!items.empty()) {
    if (j->need_scan || j->probe_kind == PK_MARK_ONLY) return fail(TG_ERR_UNSUPPORTED, "OtherCondition / several key columns with a build-side scan (outer side / left side is the build side) are not offloaded");
    if (j->has_flag_col) return fail(TG_ERR_UNSUPPORTED, "OtherCondition / several key columns on left outer semi joins (NULL-aware match flag) are not offloaded");
    for (size_t i = 0; i < items.size(); i++) {
      tg_other_item it = items[i];
      if (it.op < TG_CMP_LT || it.op > TG_CMP_NE) return fail(TG_ERR_INVALID, "bad OtherCondition op");
      auto remap = [&](int32_t& side, int32_t col, bool may_be_const) -> int {
        if (side < 0) return may_be_const ? TG_OK : fail(TG_ERR_INVALID, "OtherCondition: the left operand must be a column");
        if (side > 1) return fail(TG_ERR_INVALID, "OtherCondition side must be 0 (left) or 1 (right)");
        const Side& sd = side == 0 ? left : right;
        if (col < 0 || col >= sd.ncols) return fail(TG_ERR_INVALID, "OtherCondition column out of range");
        if (sd.elem[col] != 8) return fail(TG_ERR_UNSUPPORTED, "OtherCondition is offloaded on 8-byte columns only");
        if ((sd.types[col] == TG_TYPE_DOUBLE) != (it.is_real != 0)) return fail(TG_ERR_UNSUPPORTED, "OtherCondition compares columns of different families (the planner casts first)");
        const bool is_build = (side == 1) == brt;
        Side& mine = is_build ? j->build : j->probe;
        mine.needed[col] = 1;
        if (is_build && std::find(j->other_build_cols.begin(), j->other_build_cols.end(), col) == j->other_build_cols.end()) j->other_build_cols.push_back(col);
        side = is_build ? 1 : 0;
        return TG_OK;
      };
      TG_TRY(remap(it.lhs_side, it.lhs_col, false));
      TG_TRY(remap(it.rhs_side, it.rhs_col, true));
      j->other.push_back(it);
    }
  }
  for (Side* s : {&j->build, &j->probe}) {
    if (s->key_col >= 0) s->needed[s->key_col] = 1;
    for (int c : s->key_cols) s->needed[c] = 1;
    for (int c : s->used) {
      if (s->elem[c] != 8 && s->elem[c] != 4) return fail(TG_ERR_UNSUPPORTED, "only 4/8-byte fixed-width columns are offloaded (no DECIMAL / var-len yet)");
      s->needed[c] = 1;
    }
    for (int i = 0; i < s->filter.n; i++) {
      s->needed[s->filter.items[i].lhs_col] = 1;
      if (s->filter.items[i].rhs_col >= 0) s->needed[s->filter.items[i].rhs_col] = 1;
    }
  }
  // key specs; mixed signedness (NeedSignFlag, join_table_meta.go:296-303): the signed side's negative
  // values can never match
  int bk = j->build.key_col, pk = j->probe.key_col;
  j->build_key = KeySpec{nullptr, nullptr, j->multi_key ? KEY_I64 : key_kind_of(j->build.types[bk]), 0};
  j->probe_key = KeySpec{nullptr, nullptr, j->multi_key ? KEY_I64 : key_kind_of(j->probe.types[pk]), 0};
  if (!j->multi_key && j->build_key.kind == KEY_I64) {
    bool bu = key_unsigned(j->build.types[bk], j->build.flags[bk]), pu = key_unsigned(j->probe.types[pk], j->probe.flags[pk]);
    if (bu != pu) { j->build_key.reject_negative = !bu; j->probe_key.reject_negative = !pu; }
  }
  // output schema: LUsed of left ‖ RUsed of right [‖ matched flag]
  j->n_lused = (int)j->lused.size(); j->n_rused = (int)j->rused.size();
  j->out_elem.clear();
  for (int c : j->lused) j->out_elem.push_back(left.elem[c]);
  for (int c : j->rused) j->out_elem.push_back(right.elem[c]);
  if (j->has_flag_col) j->out_elem.push_back(8);
  j->n_out = (int)j->out_elem.size();
  if (j->n_out > TG_MAX_OUT) return fail(TG_ERR_UNSUPPORTED, "too many output columns");
  j->device = d->device;
  j->default_load_factor = !(d->load_factor > 0.05 && d->load_factor <= 0.95);
  j->load_factor = (d->load_factor > 0.05 && d->load_factor <= 0.95) ? d->load_factor : 0.35;   // measured best with the lean segment probe (profiles/r2_sweep_probe_lf_parts.jsonl: 1.93 vs 1.96 ms at 0.4, 2.12 at 0.5)
  return TG_OK;
}

// ---- host chunk → staging ---------------------------------------------------------------------------
static int64_t chunk_logical_rows(const tg_chunk* c) { return c->sel ? c->nsel : (c->ncols > 0 ? c->cols[0].length : 0); }

static int validate_chunk(const Side& s, const tg_chunk* chk) {
  if (!chk || chk->ncols != s.ncols) return fail(TG_ERR_INVALID, "chunk column count does not match the child schema");
  int64_t phys = chk->ncols ? chk->cols[0].length : 0;
  for (int c = 0; c < s.ncols; c++) {
    if (!s.needed[c]) continue;
    if (chk->cols[c].elem_len != s.elem[c]) return fail(TG_ERR_INVALID, "chunk column elem_len does not match the schema type");
    if (chk->cols[c].length != phys) return fail(TG_ERR_INVALID, "chunk columns have different lengths");
    if (phys && !chk->cols[c].data) return fail(TG_ERR_INVALID, "chunk column data is NULL");
  }
  return TG_OK;
}

// append the logical rows of a host chunk to the staging buffers (gathers through sel)
static int stage_append(HostStage& st, const Side& s, const tg_chunk* chk) {
  int64_t n = chunk_logical_rows(chk);
  if (n == 0) return TG_OK;
  for (int c = 0; c < s.ncols; c++) {
    if (!s.needed[c]) continue;
    const tg_column& col = chk->cols[c];
    int el = s.elem[c];
    PinBuf& d = *st.data[c];
    TG_TRY(d.reserve((size_t)(st.rows + n) * el));
    uint8_t* dst = d.p + (size_t)st.rows * el;
    if (!chk->sel) std::memcpy(dst, col.data, (size_t)n * el);
    else if (el == 8) { auto* o = reinterpret_cast<uint64_t*>(dst); auto* in = reinterpret_cast<const uint64_t*>(col.data); for (int64_t i = 0; i < n; i++) o[i] = in[chk->sel[i]]; }
    else { auto* o = reinterpret_cast<uint32_t*>(dst); auto* in = reinterpret_cast<const uint32_t*>(col.data); for (int64_t i = 0; i < n; i++) o[i] = in[chk->sel[i]]; }
    d.used = (size_t)(st.rows + n) * el;
    // null bitmap: materialised lazily, the first time a chunk brings one
    PinBuf& nb = *st.nulls[c];
    bool bring = col.null_bitmap != nullptr;
    if (bring || st.has_nulls[c]) {
      size_t need = (size_t)((st.rows + n + 7) / 8) + 1;
      TG_TRY(nb.reserve(need));
      if (!st.has_nulls[c]) { std::memset(nb.p, 0xff, (size_t)((st.rows + 7) / 8) + 1); st.has_nulls[c] = 1; }
      if (bring && !chk->sel) append_bits(nb.p, st.rows, col.null_bitmap, n);
      else {
        for (int64_t i = 0; i < n; i++) {
          bool nn = bring ? bit_not_null(col.null_bitmap, chk->sel ? chk->sel[i] : i) : true;
          int64_t r = st.rows + i;
          if (nn) nb.p[r >> 3] |= (uint8_t)(1u << (r & 7)); else nb.p[r >> 3] &= (uint8_t)~(1u << (r & 7));
        }
      }
      nb.used = need;
    }
  }
  st.rows += n;
  return TG_OK;
}

// staging → device column store (replaces its contents)
static int stage_to_device(JoinImpl* j, HostStage& st, const Side& s, ColStore& cs) {
  cs.rows = st.rows;
  for (int c = 0; c < s.ncols; c++) {
    if (!s.needed[c]) continue;
    size_t bytes = (size_t)st.rows * s.elem[c];
    TG_TRY(cs.data[c]->ensure(j->device, bytes + 16));
    if (bytes) { TG_CUDA(cudaMemcpyAsync(cs.data[c]->p, st.data[c]->p, bytes, cudaMemcpyHostToDevice, j->stream)); j->stats.h2d_bytes += bytes; }
    cs.has_nulls[c] = st.has_nulls[c];
    if (st.has_nulls[c]) {
      size_t nb = (size_t)((st.rows + 7) / 8);
      TG_TRY(cs.nulls[c]->ensure(j->device, nb + 16));
      if (nb) { TG_CUDA(cudaMemcpyAsync(cs.nulls[c]->p, st.nulls[c]->p, nb, cudaMemcpyHostToDevice, j->stream)); j->stats.h2d_bytes += nb; }
    }
  }
  return TG_OK;
}

// a whole (large) host chunk → device column store, straight from the caller's buffers
static int chunk_to_device(JoinImpl* j, const tg_chunk* chk, const Side& s, ColStore& cs) {
  int64_t n = chk->cols[0].length;
  cs.rows = n;
  for (int c = 0; c < s.ncols; c++) {
    if (!s.needed[c]) continue;
    size_t bytes = (size_t)n * s.elem[c];
    TG_TRY(cs.data[c]->ensure(j->device, bytes + 16));
    TG_CUDA(cudaMemcpyAsync(cs.data[c]->p, chk->cols[c].data, bytes, cudaMemcpyHostToDevice, j->stream));
    j->stats.h2d_bytes += bytes;
    cs.has_nulls[c] = chk->cols[c].null_bitmap != nullptr;
    if (cs.has_nulls[c]) {
      size_t nb = (size_t)((n + 7) / 8);
      TG_TRY(cs.nulls[c]->ensure(j->device, nb + 16));
      TG_CUDA(cudaMemcpyAsync(cs.nulls[c]->p, chk->cols[c].null_bitmap, nb, cudaMemcpyHostToDevice, j->stream));
      j->stats.h2d_bytes += nb;
    }
  }
  return TG_OK;
}

// device chunk → device column store (device-to-device, appended)
static int devchunk_append(JoinImpl* j, const tg_chunk* chk, const Side& s, ColStore& cs) {
  if (chk->sel) return fail(TG_ERR_UNSUPPORTED, "device-resident chunks must not carry a sel vector");
  int64_t n = chk->ncols ? chk->cols[0].length : 0;
  if (n == 0) return TG_OK;
  for (int c = 0; c < s.ncols; c++) {
    if (!s.needed[c]) continue;
    int el = s.elem[c];
    TG_TRY(cs.data[c]->ensure_preserve(j->device, (size_t)(cs.rows + n) * el + 16, (size_t)cs.rows * el, j->stream));
    TG_CUDA(cudaMemcpyAsync(cs.data[c]->as<uint8_t>() + (size_t)cs.rows * el, chk->cols[c].data, (size_t)n * el, cudaMemcpyDeviceToDevice, j->stream));
    if (chk->cols[c].null_bitmap || cs.has_nulls[c]) {
      if (cs.rows % 8 != 0) return fail(TG_ERR_UNSUPPORTED, "device-resident chunks with NULL bitmaps must start at a multiple of 8 rows");
      size_t need = (size_t)((cs.rows + n + 7) / 8) + 16;
      size_t had = (size_t)((cs.rows + 7) / 8);
      TG_TRY(cs.nulls[c]->ensure_preserve(j->device, need, cs.has_nulls[c] ? had : 0, j->stream));
      if (!cs.has_nulls[c]) { TG_CUDA(cudaMemsetAsync(cs.nulls[c]->p, 0xff, had, j->stream)); cs.has_nulls[c] = 1; }
      if (chk->cols[c].null_bitmap) TG_CUDA(cudaMemcpyAsync(cs.nulls[c]->as<uint8_t>() + had, chk->cols[c].null_bitmap, (size_t)((n + 7) / 8), cudaMemcpyDeviceToDevice, j->stream));
      else TG_CUDA(cudaMemsetAsync(cs.nulls[c]->as<uint8_t>() + had, 0xff, (size_t)((n + 7) / 8), j->stream));
    }
  }
  cs.rows += n;
  return TG_OK;
}

// borrow a device chunk as a column view (no copy)
static int devchunk_view(const tg_chunk* chk, const Side& s, DevCols& v, int64_t* rows) {
  if (!chk || chk->ncols != s.ncols) return fail(TG_ERR_INVALID, "chunk column count does not match the child schema");
  if (chk->sel) return fail(TG_ERR_UNSUPPORTED, "device-resident chunks must not carry a sel vector");
  std::memset(&v, 0, sizeof(v));
  *rows = chk->ncols ? chk->cols[0].length : 0;
  for (int c = 0; c < s.ncols; c++) {
    v.elem_len[c] = s.elem[c];
    if (!s.needed[c]) continue;
    if (chk->cols[c].elem_len != s.elem[c]) return fail(TG_ERR_INVALID, "chunk column elem_len does not match the schema type");
    if (chk->cols[c].length != *rows) return fail(TG_ERR_INVALID, "chunk columns have different lengths");
    v.data[c] = chk->cols[c].data;
    v.nulls[c] = chk->cols[c].null_bitmap;
  }
  return TG_OK;
}

// several equal conditions: the synthetic candidate-key column of one side's `n` device-resident rows (k_composite_key)
static int composite_key(JoinImpl* j, const Side& s, const DevCols& v, int64_t n, DevBuf& key, DevBuf& not_null) {
  TG_TRY(key.ensure(j->device, (size_t)(n + 1) * 8 + 16));
  TG_TRY(not_null.ensure(j->device, (size_t)((n + 31) / 32) * 4 + 16));
  if (n <= 0) return TG_OK;
  MultiKeySrc src{};
  src.nk = (int)s.key_cols.size();
  for (int q = 0; q < src.nk; q++) {
    const int c = s.key_cols[q];
    src.data[q] = reinterpret_cast<const int64_t*>(v.data[c]);
    src.nulls[q] = v.nulls[c];
    src.reject[q] = s.key_reject[q];
    if (!src.data[q]) return fail(TG_ERR_INVALID, "key column data is NULL");
  }
  k_composite_key<<<grid_for(j, n, 256, 8), 256, 0, j->stream>>>(src, n, key.as<int64_t>(), not_null.as<uint32_t>());
  j->stats.kernel_launches++;
  return TG_OK;
}

// ---- build --------------------------------------------------------------------------------------------
static int build_table(JoinImpl* j) {
  const Side& b = j->build;
  int64_t n = j->bcols.rows;
  j->stats.build_rows = n;
  DevCols bview = j->bcols.view(b);
  KeySpec ks = j->build_key;
  if (j->multi_key) {
    TG_TRY(composite_key(j, b, bview, n, j->bkey_syn, j->bkey_syn_nn));
    ks.data = j->bkey_syn.p; ks.nulls = j->bkey_syn_nn.as<uint8_t>();
  } else { ks.data = bview.data[b.key_col]; ks.nulls = bview.nulls[b.key_col]; }
  unsigned long long nslots = (unsigned long long)((double)(n > 0 ? n : 1) / j->load_factor) + 32;
  // the L2 partition pass handles at most TG_MAX_PARTS slices and wants them <= ~33 MB (45 MB slices: 2.10 vs 1.93 ms,
  // profiles/r2_sweep_probe_lf_parts.jsonl): with the DEFAULT load factor a table that would need more slices is made denser,
  // down to load factor 0.5, instead of growing its slices
  if (j->default_load_factor) {
    const unsigned long long fit = ((unsigned long long)TG_MAX_PARTS * (33ull << 20)) / sizeof(Slot);
    const unsigned long long dense = (unsigned long long)((double)(n > 0 ? n : 1) / 0.5) + 32;
    if (nslots > fit) nslots = std::max(fit, dense);
  }
  nslots &= ~1ull;   // even: slots are addressed as 32-byte pairs
  const int pair_home = env_int("TG_PAIR_HOME", 1);
  if (nslots + 1 >= 0xFFFFFFFFull) return fail(TG_ERR_UNSUPPORTED, "build side too large for 32-bit slot ids");
  TG_TRY(j->table.ensure(j->device, (size_t)(nslots + 1) * sizeof(Slot)));
  TG_TRY(j->row_slot.ensure(j->device, (size_t)(n + 1) * 4));
  TG_TRY(j->row_rank.ensure(j->device, (size_t)(n + 1) * 4));
  TG_TRY(j->scalars.ensure(j->device, 64));
  Slot* slots = j->table.as<Slot>();
  unsigned long long* sc = j->scalars.as<unsigned long long>();   // [0] distinct [1] maxcnt [2] cursor
  TG_CUDA(cudaEventRecord(j->ev0, j->stream));
  TG_CUDA(cudaMemsetAsync(sc, 0, 64, j->stream));
  k_table_init<<<grid_for(j, (int64_t)nslots + 1, 256, 8), 256, 0, j->stream>>>(slots, nslots + 1, nslots);
  j->stats.kernel_launches++;
  if (n > 0) {
    k_build_insert<<<grid_for(j, n, 256, 8), 256, 0, j->stream>>>(ks, bview, b.filter, n, slots, nslots, pair_home,
                                                                  j->row_slot.as<uint32_t>(), j->row_rank.as<uint32_t>());
    j->stats.kernel_launches++;
  }
  k_table_stats<<<grid_for(j, (int64_t)nslots + 1, 256, 8), 256, 0, j->stream>>>(slots, nslots + 1, sc, sc + 1);
  j->stats.kernel_launches++;
  unsigned long long host_sc[3] = {0, 0, 0};
  TG_CUDA(cudaMemcpyAsync(host_sc, sc, 16, cudaMemcpyDeviceToHost, j->stream));
  TG_CUDA(cudaStreamSynchronize(j->stream));
  j->stats.distinct_keys = (int64_t)host_sc[0];
  j->stats.max_dup = (int64_t)host_sc[1];
  j->stats.table_slots = (int64_t)nslots;
  if (host_sc[1] > kCntMask) return fail(TG_ERR_UNSUPPORTED, "a single join key repeats more than 2^28 times on the build side");

  // choose the table mode.  U1: unique keys and the build side contributes at most its key (int64) and one
  // 8-byte NOT NULL payload column, and no build-side scan is needed afterwards.
  std::vector<int> payload;   // used build columns other than the key
  bool key_out = false;
  for (int c : b.used) { if (c == b.key_col) key_out = true; else if (std::find(payload.begin(), payload.end(), c) == payload.end()) payload.push_back(c); }
  bool u1 = host_sc[1] <= 1 && !j->need_scan && payload.size() <= 1 && (!key_out || j->build_key.kind == KEY_I64) && j->other.empty();
  if (u1 && payload.size() == 1) {
    int pc = payload[0];
    if (b.elem[pc] != 8 || j->bcols.has_nulls[pc]) u1 = false;
  }
  j->tv = TableView{slots, nslots, nullptr, 0, -1, TABLE_NONE, pair_home};
  j->build_word_of_col.assign(b.ncols, -1);
  if (u1) {
    j->u1_payload_col = payload.empty() ? -1 : payload[0];
    if (n > 0) {
      const unsigned long long* pl = j->u1_payload_col >= 0 ? reinterpret_cast<const unsigned long long*>(bview.data[j->u1_payload_col]) : nullptr;
      k_build_scatter_u1<<<grid_for(j, n, 256, 8), 256, 0, j->stream>>>(j->row_slot.as<uint32_t>(), pl, n, slots);
      j->stats.kernel_launches++;
    }
    j->tv.mode = TABLE_U1;
  } else {
    RowSpec rs{};
    int w = 0;
    bool any_nullable = false;
    std::vector<int> cols;
    for (int c : b.used) if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c);
    for (int c : j->other_build_cols) if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c);   // read by OtherCondition only
    if ((int)cols.size() + 1 > TG_MAX_COLS) return fail(TG_ERR_UNSUPPORTED, "too many build columns");
    for (int c : cols) {
      rs.col[w] = c; rs.elem_len[w] = b.elem[c];
      rs.null_bit[w] = j->bcols.has_nulls[c] ? w : -1;
      any_nullable |= j->bcols.has_nulls[c] != 0;
      j->build_word_of_col[c] = w;
      w++;
    }
    rs.null_word = any_nullable ? w : -1;
    rs.nwords = w + (any_nullable ? 1 : 0);
    if (rs.nwords == 0) rs.nwords = 1;   // semi joins: no payload at all, keep a dummy word so offsets stay valid
    j->rowspec = rs;
    k_table_assign<<<grid_for(j, (int64_t)nslots + 1, 256, 8), 256, 0, j->stream>>>(slots, nslots + 1, sc + 2);
    j->stats.kernel_launches++;
    TG_TRY(j->rows_store.ensure(j->device, (size_t)(n + 1) * rs.nwords * 8));
    if (n > 0 && w > 0) {
      k_build_scatter_rows<<<grid_for(j, n, 256, 8), 256, 0, j->stream>>>(j->row_slot.as<uint32_t>(), j->row_rank.as<uint32_t>(), n, slots,
                                                                          bview, rs, j->rows_store.as<unsigned long long>());
      j->stats.kernel_launches++;
    }
    j->tv.mode = TABLE_G;
    j->tv.rows = j->rows_store.as<unsigned long long>();
    j->tv.row_words = rs.nwords;
    j->tv.null_word = rs.null_word;
    // compile OtherCondition against the row store
    j->dev_other = DevOther{};
    j->dev_other.n = (int)j->other.size();
    for (size_t q = 0; q < j->other.size(); q++) {
      const tg_other_item& it = j->other[q];
      OtherItemDev& o = j->dev_other.it[q];
      o.op = it.op; o.is_real = it.is_real; o.l_unsigned = it.lhs_unsigned; o.r_unsigned = it.rhs_unsigned;
      o.const_i64 = it.const_i64; o.const_f64 = it.const_f64;
      auto operand = [&](int side, int col, int32_t& src, int32_t& idx, int32_t& nbit) {
        nbit = -1;
        if (side < 0) { src = OSRC_CONST; idx = 0; }
        else if (side == 0) { src = OSRC_PROBE; idx = col; }
        else { src = OSRC_BUILD; idx = j->build_word_of_col[col]; nbit = rs.null_bit[idx]; }
      };
      operand(it.lhs_side, it.lhs_col, o.l_src, o.l_idx, o.l_null_bit);
      operand(it.rhs_side, it.rhs_col, o.r_src, o.r_idx, o.r_null_bit);
    }
  }
  if (j->need_scan) {
    TG_TRY(j->slot_used.ensure(j->device, (size_t)nslots + 1));
    TG_CUDA(cudaMemsetAsync(j->slot_used.p, 0, (size_t)nslots + 1, j->stream));
  }
  TG_CUDA(cudaEventRecord(j->ev1, j->stream));
  TG_CUDA(cudaStreamSynchronize(j->stream));
  TG_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, j->ev0, j->ev1);
  j->stats.build_ms = ms;
  j->stats.table_mode = j->tv.mode;
  // valid keys = rows that landed in the table
  j->stats.build_valid_keys = -1;
  j->built = true;
  return TG_OK;
}

// ---- output plumbing -------------------------------------------------------------------------------------
static int ensure_result(JoinImpl* j, ResultBatch& rb, int64_t cap_rows, bool preserve, int64_t used_rows) {
  if ((int)rb.cols.size() != j->n_out) {
    rb.cols.clear(); rb.bitmaps.clear();
    for (int i = 0; i < j->n_out; i++) { rb.cols.emplace_back(new DevBuf()); rb.bitmaps.emplace_back(new DevBuf()); }
  }
  for (int c = 0; c < j->n_out; c++) {
    size_t bytes = (size_t)(cap_rows + 8) * j->out_elem[c];
    if (preserve) TG_TRY(rb.cols[c]->ensure_preserve(j->device, bytes, (size_t)used_rows * j->out_elem[c], j->stream));
    else TG_TRY(rb.cols[c]->ensure(j->device, bytes));
  }
  return TG_OK;
}

// which output columns can carry NULLs for this probe batch
static void out_nullable(const JoinImpl* j, const DevCols& pview, std::vector<char>& nullable) {
  nullable.assign(j->n_out, 0);
  bool probe_is_left = j->build_is_right;
  int n_l = j->n_lused;
  for (int o = 0; o < j->n_out; o++) {
    if (j->has_flag_col && o == j->n_out - 1) { nullable[o] = 0; continue; }
    bool from_left = o < n_l;
    bool from_probe = from_left == probe_is_left;
    int col = from_left ? j->lused[o] : j->rused[o - n_l];
    if (from_probe) nullable[o] = pview.nulls[col] != nullptr || j->need_scan;   // scan phase NULL-pads the probe side
    else nullable[o] = j->bcols.has_nulls[col] || j->probe_kind == PK_PROBE_OUTER;
  }
}

static void fill_outspec_probe(const JoinImpl* j, OutCols& oc) {
  bool probe_is_left = j->build_is_right;
  int n_l = j->n_lused;
  oc.n = j->n_out;
  for (int o = 0; o < j->n_out; o++) {
    OutSpec& sp = oc.spec[o];
    sp.elem_len = j->out_elem[o];
    sp.null_bit = -1;
    if (j->has_flag_col && o == j->n_out - 1) { sp.src = SRC_FLAG; sp.idx = 0; continue; }
    bool from_left = o < n_l;
    int col = from_left ? j->lused[o] : j->rused[o - n_l];
    if (from_left == probe_is_left) { sp.src = SRC_PROBE_COL; sp.idx = col; continue; }
    if (j->tv.mode == TABLE_U1) { sp.src = col == j->build.key_col ? SRC_BUILD_KEY : SRC_BUILD_META; sp.idx = 0; }
    else { sp.src = SRC_BUILD_WORD; sp.idx = j->build_word_of_col[col]; sp.null_bit = j->rowspec.null_bit[sp.idx]; }
  }
}

static int scan_counts(JoinImpl* j, int64_t n, unsigned long long* total_out) {
  int64_t nblocks = (n + TG_SCAN_BLOCK * TG_SCAN_ITEMS - 1) / (TG_SCAN_BLOCK * TG_SCAN_ITEMS);
  TG_TRY(j->tmp_sums.ensure(j->device, (size_t)(nblocks + 2) * 8));
  TG_TRY(j->tmp_off.ensure(j->device, (size_t)(n + 2) * 8));
  k_scan_block_sums<<<(unsigned)nblocks, TG_SCAN_BLOCK, 0, j->stream>>>(j->tmp_cnt.as<uint32_t>(), n, j->tmp_sums.as<unsigned long long>());
  k_scan_sums<<<1, 1024, 0, j->stream>>>(j->tmp_sums.as<unsigned long long>(), nblocks);
  k_scan_write<<<(unsigned)nblocks, TG_SCAN_BLOCK, 0, j->stream>>>(j->tmp_cnt.as<uint32_t>(), n, j->tmp_sums.as<unsigned long long>(),
                                                                  j->tmp_off.as<unsigned long long>());
  j->stats.kernel_launches += 3;
  TG_CUDA(cudaMemcpyAsync(total_out, j->tmp_sums.as<unsigned long long>() + nblocks, 8, cudaMemcpyDeviceToHost, j->stream));
  TG_CUDA(cudaStreamSynchronize(j->stream));
  return TG_OK;
}

// valid-byte streams → bitmaps for the nullable output columns
static int finish_bitmaps(JoinImpl* j, ResultBatch& rb, const std::vector<char>& nullable) {
  for (int c = 0; c < j->n_out; c++) {
    if (!nullable[c]) { rb.bitmaps[c]->release(); continue; }
    TG_TRY(rb.bitmaps[c]->ensure(j->device, (size_t)((rb.rows + 7) / 8) + 16));
    if (rb.rows) {
      k_pack_bitmap<<<grid_for(j, (rb.rows + 7) / 8, 256, 8), 256, 0, j->stream>>>(j->tmp_valid[c]->as<uint8_t>(), rb.rows, rb.bitmaps[c]->as<uint8_t>());
      j->stats.kernel_launches++;
    }
  }
  return TG_OK;
}

static bool fast_path_ok(const JoinImpl* j, const DevCols& pview) {
  if (j->tv.mode != TABLE_U1 || j->probe_kind != PK_INNER || j->need_scan) return false;
  if (j->probe.filter.n || j->probe_key.kind != KEY_I64 || j->probe_key.reject_negative) return false;
  if (pview.nulls[j->probe.key_col]) return false;
  for (int c : j->probe.used) if (j->probe.elem[c] != 8 || pview.nulls[c]) return false;
  for (int e : j->out_elem) if (e != 8) return false;
  return true;
}

// single-pass unique-key probe (k_probe_inner_uq): inner join, every build key unique, no OtherCondition, 8-byte output
// columns that cannot be NULL (probe used columns without a bitmap in this batch, build columns without NULLs)
static bool uq_path_ok(const JoinImpl* j, const DevCols& pview) {
  static int en = -1;
  if (en < 0) en = env_int("TG_PROBE_UQ", 1);
  if (!en || j->probe_kind != PK_INNER || j->need_scan || !j->other.empty() || j->stats.max_dup > 1) return false;
  if (j->tv.mode != TABLE_U1 && j->tv.mode != TABLE_G) return false;
  for (int c : j->probe.used) if (j->probe.elem[c] != 8 || pview.nulls[c]) return false;
  for (int c : j->build.used) if (j->build.elem[c] != 8 || j->bcols.has_nulls[c]) return false;
  for (int e : j->out_elem) if (e != 8) return false;
  return true;
}

// ---- fast-path launch tuning (env overrides are for A/B sweeps on the GPU box; defaults are the measured best) ----
struct ProbeTuning { int variant; int R; int evict_last; int ctas_per_sm; int partition; int subseg; int parts; int part_min_mb; int part_min_rows; int seg_vec; int seg_lean; int carveout; int tma; int stages; int tma_ctas; int cta_agg; };
static ProbeTuning probe_tuning() {
  ProbeTuning t;
  t.variant = env_int("TG_PROBE_VARIANT", 1);      // 0: CTA-tile kernel (shared-memory offsets), 1: warp-autonomous kernel
  t.R = env_int("TG_PROBE_R", 4);
  t.evict_last = env_int("TG_PROBE_EVICT_LAST", 0);
  t.ctas_per_sm = env_int("TG_PROBE_CTAS_PER_SM", 0);   // 0 = exactly the resident CTA count (occupancy query)
  t.partition = env_int("TG_PROBE_PARTITION", 1);   // regroup big probes into L2-sized partitions first (0 = never, 2 = counted/dense variant)
  t.subseg = env_int("TG_PROBE_SUBSEG", 0);         // 1 = CTA-private sub-segments in the L2 partition pass (no global cursor atomics): MEASURED SLOWER (scatter 0.62 vs 0.575 ms: 7104 write streams; probe 2.3 vs 1.36 ms: the interleaved empty tails let warps drift across partitions) - kept for the record, off
  t.parts = env_int("TG_PROBE_PARTS", 0);           // 0 = auto: table slices of <= 32 MB
  t.part_min_mb = env_int("TG_PROBE_PART_MIN_MB", 64);
  t.part_min_rows = env_int("TG_PROBE_PART_MIN_ROWS", 1 << 22);
  t.seg_vec = env_int("TG_PROBE_SEG_VEC", 1);            // 128-bit loads/stores in the segment probe
  t.seg_lean = env_int("TG_PROBE_SEG_LEAN", 1);          // 1 = lean full-tile path (default: 1.954 vs 2.089 ms per step, profiles/r2_sweep_probe.jsonl), 0 = round-1 kernel, 2 = + register prefetch (2.01 ms)
  t.carveout = env_int("TG_PROBE_CARVEOUT", -1);         // EXPERIMENTAL: preferred shared-memory carve-out (%) of the segment probe kernels, -1 = driver default
  t.tma = 0;                                        // (the TMA-fed probe kernels were removed in round 2)
  t.stages = env_int("TG_PROBE_STAGES", 4);
  t.tma_ctas = env_int("TG_PROBE_TMA_CTAS", 3);
  t.cta_agg = env_int("TG_PROBE_CTA_AGG", 1);        // one output-cursor atomic per CTA tile (TMA kernel)
  return t;
}

// classify the output columns of the fused fast path by the register that feeds them; false = shape not covered by
// the templated kernels (the CTA-tile kernel handles it)
static bool build_fast_out(const JoinImpl* j, const OutCols& oc, const DevCols& pview, FastOut& fo) {
  std::memset(&fo, 0, sizeof(fo));
  int pcol_of[TG_FAST_MAX_PCOLS];
  for (int c = 0; c < oc.n; c++) {
    const OutSpec& sp = oc.spec[c];
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(oc.data[c]);
    if (sp.src == SRC_BUILD_KEY || (sp.src == SRC_PROBE_COL && sp.idx == j->probe.key_col)) {
      if (fo.n_key_dst >= TG_FAST_MAX_KEYDST) return false;
      fo.key_dst[fo.n_key_dst++] = dst;
    } else if (sp.src == SRC_BUILD_META) {
      if (fo.n_meta_dst >= TG_FAST_MAX_METADST) return false;
      fo.meta_dst[fo.n_meta_dst++] = dst;
    } else if (sp.src == SRC_PROBE_COL) {
      for (int q = 0; q < fo.n_pcols; q++) if (pcol_of[q] == sp.idx) return false;   // same column twice
      if (fo.n_pcols >= TG_FAST_MAX_PCOLS) return false;
      int k = fo.n_pcols++;
      pcol_of[k] = sp.idx;
      fo.psrc[k] = reinterpret_cast<const unsigned long long*>(pview.data[sp.idx]);
      fo.pdst[k] = dst;
    } else return false;
  }
  return true;
}

// (NPC, NKD, NMD) dispatch
template <template <int, int, int> class F, typename... A>
static int dispatch_shape(const FastOut& fo, A&&... a) {
#define TG_SHAPE(P, K, M) if (fo.n_pcols == P && fo.n_key_dst == K && fo.n_meta_dst == M) return F<P, K, M>::run(a...);
#define TG_SHAPE_KM(P) TG_SHAPE(P, 0, 0) TG_SHAPE(P, 0, 1) TG_SHAPE(P, 1, 0) TG_SHAPE(P, 1, 1) TG_SHAPE(P, 2, 0) TG_SHAPE(P, 2, 1)
  TG_SHAPE_KM(0) TG_SHAPE_KM(1) TG_SHAPE_KM(2) TG_SHAPE_KM(3)
#undef TG_SHAPE_KM
#undef TG_SHAPE
  return fail(TG_ERR_CUDA, "internal: fused probe shape not instantiated");
}

template <int NPC, int NKD, int NMD>
struct LaunchWarp {
  static int run(JoinImpl* j, const int64_t* pkey, int64_t n, const FastOut& fo, unsigned long long* cur, const ProbeTuning& t, const SegSpec& seg) {
    constexpr int R = 4;
    static int resident = 0;   // CTAs of this instantiation one SM holds (register-bound, 3 on sm_100a)
    if (!resident) {
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_probe_inner_u1_w<R, NPC, NKD, NMD>, 256, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 3; }
      resident = nb;
      if (getenv("TG_DEBUG")) fprintf(stderr, "[tidbgpu] k_probe_inner_u1_w<%d,%d,%d>: %d resident CTAs per SM\n", NPC, NKD, NMD, nb);
    }
    int64_t tiles = (n + 32 * R - 1) / (32 * R);
    int64_t ctas = (tiles + 7) / 8;
    int per_sm = t.ctas_per_sm > 0 ? t.ctas_per_sm : resident;
    int grid = (int)std::min<int64_t>(ctas, (int64_t)j->nsm * per_sm);
    k_probe_inner_u1_w<R, NPC, NKD, NMD><<<grid, 256, 0, j->stream>>>(pkey, n, j->tv, fo, cur, seg);
    return TG_OK;
  }
};
template <int NPC, int NKD, int NMD>
struct LaunchSeg {
  static int run(JoinImpl* j, const int64_t* pkey, int64_t n, const FastOut& fo, unsigned long long* cur, const ProbeTuning& t, const SegSpec& seg) {
    static int resident = 0;
    if (!resident) {
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_probe_inner_u1_seg<NPC, NKD, NMD>, 256, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 3; }
      resident = nb;
    }
    int64_t ctas = (n / 128 + 7) / 8;
    int per_sm = t.ctas_per_sm > 0 ? t.ctas_per_sm : resident;
    int grid = (int)std::min<int64_t>(ctas, (int64_t)j->nsm * per_sm);
    if (t.seg_lean && j->tv.pair_home) {   // lean variants (see join_kernels.cuh); 3 CTAs per SM as well
      if (t.carveout >= 0) {
        cudaFuncSetAttribute(k_probe_inner_u1_seg_lean<NPC, NKD, NMD, false>, cudaFuncAttributePreferredSharedMemoryCarveout, t.carveout);
        cudaFuncSetAttribute(k_probe_inner_u1_seg_lean<NPC, NKD, NMD, true>, cudaFuncAttributePreferredSharedMemoryCarveout, t.carveout);
      }
      if (t.seg_lean >= 2) k_probe_inner_u1_seg_lean<NPC, NKD, NMD, true><<<grid, 256, 0, j->stream>>>(pkey, n, j->tv, fo, cur, seg);
      else k_probe_inner_u1_seg_lean<NPC, NKD, NMD, false><<<grid, 256, 0, j->stream>>>(pkey, n, j->tv, fo, cur, seg);
      return TG_OK;
    }
    if (t.carveout >= 0) cudaFuncSetAttribute(k_probe_inner_u1_seg<NPC, NKD, NMD>, cudaFuncAttributePreferredSharedMemoryCarveout, t.carveout);
    k_probe_inner_u1_seg<NPC, NKD, NMD><<<grid, 256, 0, j->stream>>>(pkey, n, j->tv, fo, cur, seg);
    return TG_OK;
  }
};
static int launch_probe_warp(JoinImpl* j, const int64_t* pkey, int64_t n, const FastOut& fo, unsigned long long* cur, const ProbeTuning& t,
                             const SegSpec& seg = SegSpec{nullptr, 0, 0, 0, nullptr}) {
  return dispatch_shape<LaunchWarp>(fo, j, pkey, n, fo, cur, t, seg);
}
static int launch_probe_seg(JoinImpl* j, const int64_t* pkey, int64_t n, const FastOut& fo, unsigned long long* cur, const ProbeTuning& t, const SegSpec& seg) {
  return dispatch_shape<LaunchSeg>(fo, j, pkey, n, fo, cur, t, seg);
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// probe `n` device-resident rows; results are appended to rb (rb.rows advanced)
static int probe_device(JoinImpl* j, const DevCols& pview, int64_t n, ResultBatch& rb, bool sync_count, const SegSpec* in_seg = nullptr) {
  const Side& p = j->probe;
  j->stats.probe_rows += n;
  KeySpec ks = j->probe_key;
  if (j->multi_key) {
    if (in_seg) return fail(TG_ERR_UNSUPPORTED, "segmented device chunks: joins on several key columns take the general path");
    TG_TRY(composite_key(j, p, pview, n, j->pkey_syn, j->pkey_syn_nn));
    ks.data = j->pkey_syn.p; ks.nulls = j->pkey_syn_nn.as<uint8_t>();
  } else { ks.data = pview.data[p.key_col]; ks.nulls = pview.nulls[p.key_col]; }
  TG_TRY(j->out_cursor.ensure(j->device, 64));
  if (in_seg && !fast_path_ok(j, pview)) return fail(TG_ERR_UNSUPPORTED, "segmented device chunks are only accepted by the fused fast path (unique build keys, <= 1 payload, no filters)");
  if (fast_path_ok(j, pview)) {
    TG_TRY(ensure_result(j, rb, rb.rows + n, rb.rows > 0, rb.rows));
    OutCols oc{};
    fill_outspec_probe(j, oc);
    for (int c = 0; c < j->n_out; c++) { oc.data[c] = rb.cols[c]->as<uint8_t>() + (size_t)rb.rows * 8; oc.valid[c] = nullptr; if (rb.bitmaps[c]->p) rb.bitmaps[c]->release(); }
    unsigned long long* cur = j->out_cursor.as<unsigned long long>();
    TG_CUDA(cudaMemsetAsync(cur, 0, 8, j->stream));
    if (n > 0) {
      const ProbeTuning& tune = probe_tuning();
      FastOut fo{};
      bool warp_ok = (tune.variant != 0 || in_seg) && build_fast_out(j, oc, pview, fo);
      if (in_seg && !warp_ok) return fail(TG_ERR_UNSUPPORTED, "segmented device chunks: output shape not covered by the warp kernels");
      if (warp_ok) {
        const int64_t* pkey = reinterpret_cast<const int64_t*>(ks.data);
        size_t table_bytes = (size_t)j->tv.nslots * sizeof(Slot);
        bool partitioned = false;   // the L2 partition pass + segment probe took the whole call
        bool src16 = aligned16(pkey);
        for (int c = 0; c < fo.n_pcols; c++) src16 = src16 && aligned16(fo.psrc[c]);
        const int64_t PTILE = 1024;   // rows per scatter tile (k_partition_scatter_bulk<.., 4>)
        if (tune.partition == 1 && src16 && scatter_bulk_enabled() && n >= (int64_t)tune.part_min_rows && table_bytes > ((size_t)tune.part_min_mb << 20)) {
          // L2 partition pass, count-free: regroup the probe rows by the TOP hash bits into P fixed-capacity segments.
          // slot = mulhi(hash, nslots) is monotone in the hash, so segment p only touches the contiguous table slice
          // [p/P, (p+1)/P) — ~32 MB that stay L2 resident while the probe kernel sweeps the segment.  Trades 32 B/row of
          // extra streaming traffic (0.57 ms per 100 M rows) for ~100 B/row of random HBM traffic (probe 2.5 → 1.45 ms);
          // profiles/r1_probe_lab.md.  A skewed probe side that overflows a segment raises `flag`; the partitioned probe
          // launch then exits at once and the gated direct launch behind it does the work — no host round trip.
          int P = tune.parts > 0 ? tune.parts : (int)((table_bytes + (32u << 20) - 1) / (32u << 20));
          if (P > TG_MAX_PARTS) P = TG_MAX_PARTS;
          const int64_t n_main = n / PTILE * PTILE;
          const int nc = 1 + fo.n_pcols;
          // Segment layout.  Default: one segment per partition filled through global cursors.  TG_PROBE_SUBSEG=1 (experiment,
          // measured slower, profiles/r2_subseg.md): every scatter CTA owns a private sub-segment of each partition —
          // G = grid, sub-segment (p, b) = rows [(p*G + b) * C, ...) — placed with a shared-memory cursor, no global atomics.
          const int G = tune.subseg ? scatter_bulk_grid_nc(j->device, n_main, nc) : 1;
          const int64_t nsegs = (int64_t)P * G;
          const int64_t C = tune.subseg ? ((int64_t)((double)n_main / nsegs * 1.06) + PTILE / P + 256 + 127) / 128 * 128   // + one tile's share: CTAs differ by a tile
                                        : ((int64_t)((double)n_main / P * 1.05) + 16384 + 127) / 128 * 128;
          if (P >= 2 && G >= 1 && nsegs * C / 128 < (1ll << 31)) {
            for (int c = 0; c < nc; c++) {
              if (!j->part_cols[c]) j->part_cols[c].reset(new DevBuf());
              TG_TRY(j->part_cols[c]->ensure(j->device, (size_t)nsegs * C * 8 + 64));
            }
            const int64_t ncur = std::max<int64_t>(nsegs, TG_MAX_PARTS);                 // k_segment_bases zeroes TG_MAX_PARTS cursors
            TG_TRY(j->part_scratch.ensure(j->device, (size_t)(ncur + 2 * TG_MAX_PARTS) * 8 + 64));
            unsigned long long* cursors = j->part_scratch.as<unsigned long long>();     // fill count per segment
            long long* bases = reinterpret_cast<long long*>(cursors + ncur);             // first row of each segment (one per partition)
            unsigned long long* flag = cursors + ncur + TG_MAX_PARTS;                    // overflow
            PartDst d{};
            d.nparts = P; d.ncols = nc;
            d.src[0] = pkey;
            for (int c = 0; c < fo.n_pcols; c++) d.src[1 + c] = fo.psrc[c];
            for (int c = 0; c < nc; c++) for (int q = 0; q < P; q++) d.dst[q][c] = j->part_cols[c]->p;
            if (tune.subseg) {
              TG_CUDA(cudaMemsetAsync(flag, 0, 8, j->stream));
              d.dst_base = nullptr; d.base_const = 0; d.capacity = C; d.overflow = flag; d.sub_cap = C; d.sub_grid = (uint32_t)G;
            } else {
              k_segment_bases<<<1, 32, 0, j->stream>>>(cursors, bases, flag, P, C);
              d.dst_base = bases; d.capacity = C; d.overflow = flag;
            }
            if (in_seg) { d.in_cnt = in_seg->cnt; d.in_cap = in_seg->cap; d.in_tiles_per_seg = (uint32_t)(in_seg->cap / PTILE); }
            TG_TRY(launch_partition_scatter<true>(j->device, j->stream, reinterpret_cast<const long long*>(pkey), nullptr, n_main, d, cursors,
                                                  &j->stats.kernel_launches));
            FastOut pf = fo;
            for (int c = 0; c < fo.n_pcols; c++) pf.psrc[c] = j->part_cols[1 + c]->as<unsigned long long>();
            // the 128-bit stores of the segment kernel need 16-byte aligned output columns: results appended behind an odd
            // number of rows fall back to the 8-byte kernel
            if (tune.seg_vec && (rb.rows & 1) == 0) TG_TRY(launch_probe_seg(j, j->part_cols[0]->as<int64_t>(), nsegs * C, pf, cur, tune, SegSpec{cursors, (uint32_t)(C / 128), 0, C, flag}));
            else TG_TRY(launch_probe_warp(j, j->part_cols[0]->as<int64_t>(), nsegs * C, pf, cur, tune, SegSpec{cursors, (uint32_t)(C / 128), 0, C, flag}));
            // gated fallback: probes the ORIGINAL input only after an overflow; the < 1024-row tail the scatter left behind
            // (dense input only) rides on the same launch — it is probed whatever the flag says
            if (in_seg) TG_TRY(launch_probe_warp(j, pkey, n_main, fo, cur, tune, SegSpec{in_seg->cnt, in_seg->tiles_per_seg, 1, in_seg->cap, flag, 0}));
            else TG_TRY(launch_probe_warp(j, pkey, n, fo, cur, tune, SegSpec{nullptr, 0, 1, 0, flag, n_main < n ? n_main : 0}));
            j->stats.kernel_launches += 3;
            partitioned = true;
          }
        }
        if (!partitioned && !in_seg && tune.partition == 2 && n >= (1ll << 20) && table_bytes > ((size_t)tune.part_min_mb << 20)) {
          // counted variant (kept for A/B runs): histogram pass → exact offsets → dense partitions
          int P = tune.parts > 0 ? tune.parts : (int)((table_bytes + (32u << 20) - 1) / (32u << 20));
          if (P > TG_MAX_PARTS) P = TG_MAX_PARTS;
          if (P >= 2) {
            int nc = 1 + fo.n_pcols;
            for (int c = 0; c < nc; c++) {
              if (!j->part_cols[c]) j->part_cols[c].reset(new DevBuf());
              TG_TRY(j->part_cols[c]->ensure(j->device, (size_t)n * 8 + 64));
            }
            TG_TRY(j->part_scratch.ensure(j->device, (size_t)TG_MAX_PARTS * 8 * 3 + 64));
            unsigned long long* counts = j->part_scratch.as<unsigned long long>();
            unsigned long long* cursors = counts + TG_MAX_PARTS;
            long long* offs = reinterpret_cast<long long*>(cursors + TG_MAX_PARTS);
            TG_CUDA(cudaMemsetAsync(counts, 0, (size_t)TG_MAX_PARTS * 8 * 3 + 8, j->stream));
            const long long* k64 = reinterpret_cast<const long long*>(pkey);
            TG_TRY(launch_partition_count<true>(j->device, j->stream, k64, nullptr, n, (uint32_t)P, counts, &j->stats.kernel_launches));
            k_partition_offsets<<<1, 32, 0, j->stream>>>(counts, (uint32_t)P, offs, cursors);
            j->stats.kernel_launches++;
            PartDst d{};
            d.nparts = P; d.ncols = nc;
            d.src[0] = pkey;
            for (int c = 0; c < fo.n_pcols; c++) d.src[1 + c] = fo.psrc[c];
            for (int c = 0; c < nc; c++) for (int q = 0; q < P; q++) d.dst[q][c] = j->part_cols[c]->p;
            d.dst_base = offs;
            TG_TRY(launch_partition_scatter<true>(j->device, j->stream, k64, nullptr, n, d, cursors, &j->stats.kernel_launches));
            pkey = j->part_cols[0]->as<int64_t>();
            for (int c = 0; c < fo.n_pcols; c++) fo.psrc[c] = j->part_cols[1 + c]->as<unsigned long long>();
          }
        }
        int64_t done = partitioned ? n : 0;
        if (done < n) {
          FastOut tail = fo;
          for (int c = 0; c < fo.n_pcols; c++) tail.psrc[c] = fo.psrc[c] + done;
          if (in_seg) TG_TRY(launch_probe_warp(j, pkey, n, fo, cur, tune, SegSpec{in_seg->cnt, in_seg->tiles_per_seg, 0, in_seg->cap, nullptr}));
          else TG_TRY(launch_probe_warp(j, pkey + done, n - done, tail, cur, tune));
          j->stats.kernel_launches++;
        }
      } else {
        constexpr int R = 4;
        int64_t tiles = (n + 256 * R - 1) / (256 * R);
        int grid = (int)std::min<int64_t>(tiles, (int64_t)j->nsm * (tune.ctas_per_sm > 0 ? tune.ctas_per_sm : 8));
        k_probe_inner_u1<R><<<grid, 256, 0, j->stream>>>(reinterpret_cast<const int64_t*>(ks.data), pview, n, j->tv, oc, cur);
        j->stats.kernel_launches++;
      }
    }
    if (sync_count) {
      unsigned long long got = 0;
      TG_CUDA(cudaMemcpyAsync(&got, cur, 8, cudaMemcpyDeviceToHost, j->stream));
      TG_CUDA(cudaStreamSynchronize(j->stream));
      rb.rows += (int64_t)got;
      j->stats.output_rows += (int64_t)got;
    }
    return TG_OK;
  }
  // single-pass path: inner join on unique build keys with filters / several payload columns (k_probe_inner_uq)
  if (uq_path_ok(j, pview) && !in_seg) {
    TG_TRY(ensure_result(j, rb, rb.rows + n, rb.rows > 0, rb.rows));
    OutCols oc{};
    fill_outspec_probe(j, oc);
    for (int c = 0; c < j->n_out; c++) { oc.data[c] = rb.cols[c]->as<uint8_t>() + (size_t)rb.rows * 8; oc.valid[c] = nullptr; if (rb.bitmaps[c]->p) rb.bitmaps[c]->release(); }
    unsigned long long* cur = j->out_cursor.as<unsigned long long>();
    TG_CUDA(cudaMemsetAsync(cur, 0, 8, j->stream));
    if (n > 0) {
      int64_t tiles = (n + 256 * UQ_R - 1) / (256 * UQ_R);
      int grid = (int)std::min<int64_t>(tiles, (int64_t)j->nsm * 8);
      k_probe_inner_uq<<<grid, 256, 0, j->stream>>>(ks, pview, p.filter, n, j->tv, oc, cur);
      j->stats.kernel_launches++;
    }
    // the output row count decides rb.rows (and the next append position): always needed on the host
    unsigned long long got = 0;
    TG_CUDA(cudaMemcpyAsync(&got, cur, 8, cudaMemcpyDeviceToHost, j->stream));
    TG_CUDA(cudaStreamSynchronize(j->stream));
    TG_CUDA(cudaGetLastError());
    rb.rows += (int64_t)got;
    j->stats.output_rows += (int64_t)got;
    return TG_OK;
  }
  // general path, in sub-batches
  std::vector<char> nullable;
  out_nullable(j, pview, nullable);
  if ((int)j->tmp_valid.size() != j->n_out) { j->tmp_valid.clear(); for (int i = 0; i < j->n_out; i++) j->tmp_valid.emplace_back(new DevBuf()); }
  int64_t out_start = rb.rows;
  // valid bytes are produced for the whole call, so size them after counting all sub-batches: run count+scan
  // per sub-batch, remembering totals, then write.  Sub-batching keeps the temporaries bounded.
  for (int64_t lo = 0; lo < n || (lo == 0 && n == 0); lo += kGeneralBatchRows) {
    int64_t m = std::min<int64_t>(kGeneralBatchRows, n - lo);
    if (m <= 0) break;
    DevCols sub = pview;
    if (lo) {
      if (lo % 8) return fail(TG_ERR_CUDA, "internal: sub-batch offset not byte aligned");
      for (int c = 0; c < p.ncols; c++) {
        if (sub.data[c]) sub.data[c] = reinterpret_cast<const uint8_t*>(sub.data[c]) + (size_t)lo * p.elem[c];
        if (sub.nulls[c]) sub.nulls[c] += lo / 8;
      }
    }
    KeySpec sks = ks;
    if (j->multi_key) { sks.data = reinterpret_cast<const int64_t*>(ks.data) + lo; sks.nulls = ks.nulls + lo / 8; }
    else { sks.data = sub.data[p.key_col]; sks.nulls = sub.nulls[p.key_col]; }
    TG_TRY(j->tmp_cnt.ensure(j->device, (size_t)(m + 1) * 4));
    TG_TRY(j->tmp_slot.ensure(j->device, (size_t)(m + 1) * 4));
    k_probe_count<<<grid_for(j, m, 256, 8), 256, 0, j->stream>>>(sks, sub, p.filter, j->dev_other, m, j->tv, j->probe_kind, j->tmp_cnt.as<uint32_t>(),
                                                                 j->tmp_slot.as<uint32_t>(), j->need_scan ? j->slot_used.as<uint8_t>() : nullptr);
    j->stats.kernel_launches++;
    unsigned long long total = 0;
    TG_TRY(scan_counts(j, m, &total));
    if (total) {
      TG_TRY(ensure_result(j, rb, rb.rows + (int64_t)total, true, rb.rows));
      OutCols oc{};
      fill_outspec_probe(j, oc);
      for (int c = 0; c < j->n_out; c++) {
        oc.data[c] = rb.cols[c]->p;
        oc.valid[c] = nullptr;
        if (nullable[c]) {
          TG_TRY(j->tmp_valid[c]->ensure_preserve(j->device, (size_t)(rb.rows + total) + 16, (size_t)rb.rows, j->stream));
          oc.valid[c] = j->tmp_valid[c]->as<uint8_t>();
        }
      }
      k_probe_write<<<grid_for(j, m, 256, 8), 256, 0, j->stream>>>(m, j->tmp_off.as<unsigned long long>(), j->tmp_slot.as<uint32_t>(),
                                                                   reinterpret_cast<const int64_t*>(sks.data), sks, j->tv, sub, oc, j->probe_kind,
                                                                   (unsigned long long)rb.rows, j->dev_other);
      j->stats.kernel_launches++;
      rb.rows += (int64_t)total;
      j->stats.output_rows += (int64_t)total;
    }
  }
  (void)out_start;
  if ((int)rb.cols.size() != j->n_out) TG_TRY(ensure_result(j, rb, 8, false, 0));
  TG_TRY(finish_bitmaps(j, rb, nullable));
  TG_CUDA(cudaStreamSynchronize(j->stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

// ScanRowTable after the probe side is exhausted (hash_join_v2.go:877)
static int scan_build_side(JoinImpl* j, ResultBatch& rb) {
  const Side& b = j->build;
  int64_t n = j->bcols.rows;
  if (n == 0) { if ((int)rb.cols.size() != j->n_out) TG_TRY(ensure_result(j, rb, 8, false, 0)); return TG_OK; }
  DevCols bview = j->bcols.view(b);
  TG_TRY(j->tmp_cnt.ensure(j->device, (size_t)(n + 1) * 4));
  k_build_scan_count<<<grid_for(j, n, 256, 8), 256, 0, j->stream>>>(j->row_slot.as<uint32_t>(), j->slot_used.as<uint8_t>(), n, j->scan_mode, j->tmp_cnt.as<uint32_t>());
  j->stats.kernel_launches++;
  unsigned long long total = 0;
  TG_TRY(scan_counts(j, n, &total));
  std::vector<char> nullable(j->n_out, 0);
  bool probe_is_left = j->build_is_right;
  OutCols oc{};
  oc.n = j->n_out;
  if ((int)j->tmp_valid.size() != j->n_out) { j->tmp_valid.clear(); for (int i = 0; i < j->n_out; i++) j->tmp_valid.emplace_back(new DevBuf()); }
  TG_TRY(ensure_result(j, rb, rb.rows + (int64_t)total, false, 0));
  for (int o = 0; o < j->n_out; o++) {
    bool from_left = o < j->n_lused;
    int col = from_left ? j->lused[o] : j->rused[o - j->n_lused];
    bool from_probe = from_left == probe_is_left;
    oc.spec[o].elem_len = j->out_elem[o];
    oc.spec[o].null_bit = -1;
    oc.spec[o].src = from_probe ? SRC_PROBE_COL : SRC_BUILD_WORD;
    oc.spec[o].idx = col;
    nullable[o] = from_probe || j->bcols.has_nulls[col];
    oc.data[o] = rb.cols[o]->p;
    oc.valid[o] = nullptr;
    if (nullable[o]) { TG_TRY(j->tmp_valid[o]->ensure(j->device, (size_t)total + 16)); oc.valid[o] = j->tmp_valid[o]->as<uint8_t>(); }
  }
  if (total) {
    k_build_scan_write<<<grid_for(j, n, 256, 8), 256, 0, j->stream>>>(n, j->tmp_off.as<unsigned long long>(), bview, oc, 0ull);
    j->stats.kernel_launches++;
  }
  rb.rows = (int64_t)total;
  j->stats.output_rows += (int64_t)total;
  TG_TRY(finish_bitmaps(j, rb, nullable));
  TG_CUDA(cudaStreamSynchronize(j->stream));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

static std::unique_ptr<ResultBatch> new_batch(JoinImpl* j) {
  std::lock_guard<std::mutex> lk(j->res_mu);
  if (!j->free_batches.empty()) {
    std::unique_ptr<ResultBatch> rb = std::move(j->free_batches.back());
    j->free_batches.pop_back();
    rb->rows = 0; rb->consumed = 0;
    return rb;
  }
  return std::unique_ptr<ResultBatch>(new ResultBatch());
}

static void queue_result(JoinImpl* j, std::unique_ptr<ResultBatch> rb) {
  if (rb->rows <= 0) { std::lock_guard<std::mutex> lk(j->res_mu); j->free_batches.push_back(std::move(rb)); return; }
  { std::lock_guard<std::mutex> lk(j->res_mu); j->results.push_back(std::move(rb)); }
  j->res_cv.notify_all();
}

static int flush_probe_stage(JoinImpl* j) {
  if (j->pstage.rows == 0) return TG_OK;
  TG_TRY(stage_to_device(j, j->pstage, j->probe, j->pcols_dev));
  std::unique_ptr<ResultBatch> rb = new_batch(j);
  DevCols pview = j->pcols_dev.view(j->probe);
  TG_CUDA(cudaEventRecord(j->ev0, j->stream));
  TG_TRY(probe_device(j, pview, j->pstage.rows, *rb, true));
  TG_CUDA(cudaEventRecord(j->ev1, j->stream));
  TG_CUDA(cudaStreamSynchronize(j->stream));
  float ms = 0; cudaEventElapsedTime(&ms, j->ev0, j->ev1); j->stats.probe_ms += ms;
  j->pstage.reset();
  queue_result(j, std::move(rb));
  return TG_OK;
}

}  // namespace tg

// ---------------------------------------------------------------------------------------------------
// C entry points
// ---------------------------------------------------------------------------------------------------
#define TG_LOCK(h)                                                             \
  if (!(h)) return tg::fail(TG_ERR_INVALID, "handle is NULL");                 \
  if ((h)->closed.load()) return tg::fail(TG_ERR_CANCELLED, "handle is closed"); \
  std::lock_guard<std::mutex> lock__((h)->mu);                                 \
  if ((h)->closed.load() || !(h)->impl) return tg::fail(TG_ERR_CANCELLED, "handle is closed"); \
  JoinImpl* j = (h)->impl;                                                     \
  tg::DeviceGuard guard__(j->device);                                          \
  if (!guard__.ok) return tg::fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)")

extern "C" {

int tg_join_supported(const tg_join_desc* desc) {
  tg_join shell;
  JoinImpl tmp(shell);
  return setup(&tmp, desc);
}

int tg_join_open(const tg_join_desc* desc, tg_join** out) {
  if (!out) return fail(TG_ERR_INVALID, "out is NULL");
  *out = nullptr;
  std::unique_ptr<tg_join> shell(new tg_join());
  std::unique_ptr<JoinImpl> j(new JoinImpl(*shell));
  TG_TRY(setup(j.get(), desc));
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(TG_ERR_CUDA, "no CUDA device: the GPU hash join has no CPU fallback"); }
  if (j->device < 0 || j->device >= ndev) return fail(TG_ERR_INVALID, "device ordinal out of range");
  DeviceGuard g(j->device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed");
  if (desc->stream) { j->stream = (cudaStream_t)desc->stream; j->own_stream = false; }
  else { TG_CUDA(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking)); j->own_stream = true; }
  TG_CUDA(cudaEventCreate(&j->ev0));
  TG_CUDA(cudaEventCreate(&j->ev1));
  TG_CUDA(cudaStreamCreateWithFlags(&j->d2h_stream, cudaStreamNonBlocking));
  j->nsm = device_sm_count(j->device);
  {
    // L2 fetch granularity (cudaLimitMaxL2FetchGranularity): left at the device default.  Measured (tools/sweep_probe.py,
    // profiles/r1_sweep.md): capping it at 32 B cuts the HBM bytes of the random gathers but makes the unpartitioned probe
    // SLOWER (the gathers are bound by DRAM access rate, not bytes); TG_L2_FETCH={32,64,128} overrides for experiments.
    int gran = env_int("TG_L2_FETCH", 0);
    if (gran == 32 || gran == 64 || gran == 128) { if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran) != cudaSuccess) cudaGetLastError(); }
  }
  j->bstage.init(j->build.ncols); j->bcols.init(j->build.ncols);
  j->pstage.init(j->probe.ncols); j->pcols_dev.init(j->probe.ncols);
  shell->impl = j.release();
  *out = shell.release();
  return TG_OK;
}

int tg_join_build_push(tg_join* h, const tg_chunk* chk) {
  TG_LOCK(h);
  if (j->built) return fail(TG_ERR_STATE, "build_push after build_finish");
  TG_TRY(validate_chunk(j->build, chk));
  return stage_append(j->bstage, j->build, chk);
}

int tg_join_build_push_dev(tg_join* h, const tg_chunk* chk) {
  TG_LOCK(h);
  if (j->built) return fail(TG_ERR_STATE, "build_push after build_finish");
  if (j->bstage.rows) return fail(TG_ERR_STATE, "host and device build pushes cannot be mixed");
  TG_TRY(validate_chunk(j->build, chk));
  return devchunk_append(j, chk, j->build, j->bcols);
}

int tg_join_build_finish(tg_join* h) {
  TG_LOCK(h);
  if (j->built) return fail(TG_ERR_STATE, "build_finish called twice");
  if (j->bstage.rows) { TG_TRY(stage_to_device(j, j->bstage, j->build, j->bcols)); }
  int rc = build_table(j);
  // pinned staging of the build side is no longer needed
  j->bstage.init(j->build.ncols);
  return rc;
}

int tg_join_probe_push(tg_join* h, const tg_chunk* chk) {
  TG_LOCK(h);
  if (!j->built) return fail(TG_ERR_STATE, "probe_push before build_finish");
  if (j->probe_finished.load()) return fail(TG_ERR_STATE, "probe_push after probe_finish");
  TG_TRY(validate_chunk(j->probe, chk));
  int64_t n = chunk_logical_rows(chk);
  if (n == 0) return TG_OK;
  if (!chk->sel && n >= kDirectPushRows) {
    TG_TRY(flush_probe_stage(j));
    TG_TRY(chunk_to_device(j, chk, j->probe, j->pcols_dev));
    std::unique_ptr<ResultBatch> rb = new_batch(j);
    DevCols pview = j->pcols_dev.view(j->probe);
    TG_CUDA(cudaEventRecord(j->ev0, j->stream));
    TG_TRY(probe_device(j, pview, n, *rb, true));
    TG_CUDA(cudaEventRecord(j->ev1, j->stream));
    TG_CUDA(cudaStreamSynchronize(j->stream));
    float ms = 0; cudaEventElapsedTime(&ms, j->ev0, j->ev1); j->stats.probe_ms += ms;
    queue_result(j, std::move(rb));
    return TG_OK;
  }
  TG_TRY(stage_append(j->pstage, j->probe, chk));
  if (j->pstage.rows >= kStageBatchRows) TG_TRY(flush_probe_stage(j));
  return TG_OK;
}

int tg_join_probe_finish(tg_join* h) {
  TG_LOCK(h);
  if (!j->built) return fail(TG_ERR_STATE, "probe_finish before build_finish");
  if (j->probe_finished.load()) return TG_OK;
  TG_TRY(flush_probe_stage(j));
  if (j->need_scan) {
    std::unique_ptr<ResultBatch> rb = new_batch(j);
    TG_TRY(scan_build_side(j, *rb));
    queue_result(j, std::move(rb));
  }
  { std::lock_guard<std::mutex> lk(j->res_mu); j->probe_finished.store(true); }
  j->res_cv.notify_all();
  return TG_OK;
}

static int join_next_impl(tg_join* h, tg_mut_chunk* out, int64_t max_rows, int64_t* nrows, bool wait) {
  if (!h) return fail(TG_ERR_INVALID, "handle is NULL");
  if (h->closed.load()) return fail(TG_ERR_CANCELLED, "handle is closed");
  if (!out || !nrows) return fail(TG_ERR_INVALID, "out / nrows is NULL");
  *nrows = 0;
  std::unique_lock<std::mutex> lock__(h->res_mu);
  if (h->closed.load() || !h->impl) return fail(TG_ERR_CANCELLED, "handle is closed");
  JoinImpl* j = h->impl;
  if (out->ncols != j->n_out) return fail(TG_ERR_INVALID, "output chunk column count does not match the join schema");
  for (int c = 0; c < j->n_out; c++) if (out->cols[c].elem_len != j->out_elem[c]) return fail(TG_ERR_INVALID, "output column elem_len mismatch");
  tg::DeviceGuard guard__(j->device);
  if (!guard__.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)");
  for (;;) {
    while (!j->results.empty() && j->results.front()->consumed >= j->results.front()->rows) {
      if (j->win_batch == j->results.front().get()) { j->win_batch = nullptr; j->win_lo = j->win_hi = 0; }
      j->free_batches.push_back(std::move(j->results.front()));
      j->results.pop_front();
    }
    if (!j->results.empty()) break;
    if (!wait || j->probe_finished.load()) return TG_OK;   // 0 rows: EOF iff probe_finish was called, else "push more"
    h->res_cv.wait(lock__);
    if (h->closed.load() || !h->impl) return fail(TG_ERR_CANCELLED, "handle is closed");
  }
  cudaStream_t cstream = j->d2h_stream;
  ResultBatch& rb = *j->results.front();
  int64_t want = std::min<int64_t>(std::min<int64_t>(max_rows, out->capacity_rows), rb.rows - rb.consumed);
  if (want <= 0) return TG_OK;
  // any RequiredRows >= 1 is served (LIMIT 1, MaxOneRow): the result's bit-packed NULL bitmaps are re-aligned on the host
  // when the read cursor is not on a byte boundary (see below)
  int64_t lo = rb.consumed;
  // row bytes of all columns, for the small-request window
  size_t row_bytes = 0;
  for (int c = 0; c < j->n_out; c++) row_bytes += j->out_elem[c];
  if (want >= 64 * 1024) {
    for (int c = 0; c < j->n_out; c++) {
      size_t el = j->out_elem[c];
      TG_CUDA(cudaMemcpyAsync(out->cols[c].data, rb.cols[c]->as<uint8_t>() + (size_t)lo * el, (size_t)want * el, cudaMemcpyDeviceToHost, cstream));
      j->d2h_bytes += (int64_t)want * el;
    }
  } else {
    // serve from a pinned window of up to kNextWindowRows rows fetched with one copy per column
    if (j->win_batch != &rb || lo < j->win_lo || lo + want > j->win_hi) {
      int64_t wn = std::min<int64_t>(kNextWindowRows, rb.rows - lo);
      TG_TRY(j->win.reserve((size_t)wn * row_bytes));
      size_t offb = 0;
      for (int c = 0; c < j->n_out; c++) {
        size_t el = j->out_elem[c];
        TG_CUDA(cudaMemcpyAsync(j->win.p + offb, rb.cols[c]->as<uint8_t>() + (size_t)lo * el, (size_t)wn * el, cudaMemcpyDeviceToHost, cstream));
        j->d2h_bytes += (int64_t)wn * el;
        offb += (size_t)wn * el;
      }
      TG_CUDA(cudaStreamSynchronize(cstream));
      j->win_batch = &rb; j->win_lo = lo; j->win_hi = lo + wn;
    }
    int64_t wn = j->win_hi - j->win_lo;
    size_t offb = 0;
    for (int c = 0; c < j->n_out; c++) {
      size_t el = j->out_elem[c];
      std::memcpy(out->cols[c].data, j->win.p + offb + (size_t)(lo - j->win_lo) * el, (size_t)want * el);
      offb += (size_t)wn * el;
    }
  }
  const int shift = (int)(lo & 7);
  std::vector<std::vector<uint8_t>> shifted;    // bitmaps that start inside a byte: fetched whole, shifted below
  for (int c = 0; c < j->n_out; c++) {
    size_t nb = (size_t)((want + 7) / 8);
    if (rb.bitmaps[c]->p) {
      if (!out->cols[c].null_bitmap) return fail(TG_ERR_INVALID, "output column can be NULL but the caller passed no null bitmap");
      if (shift == 0) TG_CUDA(cudaMemcpyAsync(out->cols[c].null_bitmap, rb.bitmaps[c]->as<uint8_t>() + lo / 8, nb, cudaMemcpyDeviceToHost, cstream));
      else {
        shifted.emplace_back((size_t)((shift + want + 7) / 8) + 1, (uint8_t)0);
        TG_CUDA(cudaMemcpyAsync(shifted.back().data(), rb.bitmaps[c]->as<uint8_t>() + lo / 8, shifted.back().size() - 1, cudaMemcpyDeviceToHost, cstream));
      }
      j->d2h_bytes += nb;
    } else if (out->cols[c].null_bitmap) {
      std::memset(out->cols[c].null_bitmap, 0xff, nb);
      if (want & 7) out->cols[c].null_bitmap[nb - 1] = (uint8_t)((1u << (want & 7)) - 1);
    }
  }
  TG_CUDA(cudaStreamSynchronize(cstream));
  if (shift) {
    size_t q = 0;
    for (int c = 0; c < j->n_out; c++) {
      if (!rb.bitmaps[c]->p) continue;
      const std::vector<uint8_t>& src = shifted[q++];
      size_t nb = (size_t)((want + 7) / 8);
      for (size_t b = 0; b < nb; b++) out->cols[c].null_bitmap[b] = (uint8_t)((src[b] >> shift) | (src[b + 1] << (8 - shift)));
    }
  }
  // mask the tail bits of copied bitmaps (Column.nullBitmap keeps unused bits zero)
  if (want & 7) for (int c = 0; c < j->n_out; c++) if (rb.bitmaps[c]->p) out->cols[c].null_bitmap[(want >> 3)] &= (uint8_t)((1u << (want & 7)) - 1);
  rb.consumed += want;
  *nrows = want;
  return TG_OK;
}

int tg_join_probe_rewind(tg_join* h) {
  TG_LOCK(h);
  if (!j->built) return fail(TG_ERR_STATE, "rewind before build_finish");
  if (j->need_scan) return fail(TG_ERR_UNSUPPORTED, "joins that scan the build side afterwards cannot be re-probed (used flags accumulate)");
  j->pstage.reset();
  std::lock_guard<std::mutex> lk(j->res_mu);
  while (!j->results.empty()) { j->free_batches.push_back(std::move(j->results.front())); j->results.pop_front(); }
  j->win_batch = nullptr; j->win_lo = j->win_hi = 0;
  j->probe_finished.store(false);
  return TG_OK;
}

int tg_join_next(tg_join* h, tg_mut_chunk* out, int64_t max_rows, int64_t* nrows) { return join_next_impl(h, out, max_rows, nrows, false); }
int tg_join_next_wait(tg_join* h, tg_mut_chunk* out, int64_t max_rows, int64_t* nrows) { return join_next_impl(h, out, max_rows, nrows, true); }

int tg_join_probe_dev(tg_join* h, const tg_chunk* dev_chk, int64_t* out_rows, void** out_cols, void** out_nulls) {
  TG_LOCK(h);
  if (!j->built) return fail(TG_ERR_STATE, "probe before build_finish");
  DevCols pview; int64_t n = 0;
  TG_TRY(devchunk_view(dev_chk, j->probe, pview, &n));
  if (!j->dev_result) j->dev_result.reset(new ResultBatch());
  ResultBatch& rb = *j->dev_result;
  rb.rows = 0; rb.consumed = 0;
  TG_CUDA(cudaEventRecord(j->ev0, j->stream));
  TG_TRY(probe_device(j, pview, n, rb, out_rows != nullptr));
  TG_CUDA(cudaEventRecord(j->ev1, j->stream));
  if (out_rows) {
    TG_CUDA(cudaStreamSynchronize(j->stream));
    float ms = 0; cudaEventElapsedTime(&ms, j->ev0, j->ev1); j->stats.probe_ms += ms;
    *out_rows = rb.rows;
  }
  for (int c = 0; c < j->n_out; c++) {
    if (out_cols) out_cols[c] = rb.cols[c]->p;
    if (out_nulls) out_nulls[c] = rb.bitmaps[c]->p;
  }
  return TG_OK;
}

int tg_join_probe_dev_seg(tg_join* h, const tg_chunk* dev_chk, const int64_t* seg_cnt_dev, int32_t nseg, int64_t seg_cap,
                          int64_t* out_rows, void** out_cols, void** out_nulls) {
  TG_LOCK(h);
  if (!j->built) return fail(TG_ERR_STATE, "probe before build_finish");
  if (!seg_cnt_dev || nseg < 1 || seg_cap < 1024 || seg_cap % 1024) return fail(TG_ERR_INVALID, "seg_cnt_dev required; seg_cap must be a positive multiple of 1024");
  DevCols pview; int64_t n = 0;
  TG_TRY(devchunk_view(dev_chk, j->probe, pview, &n));
  if (n != (int64_t)nseg * seg_cap) return fail(TG_ERR_INVALID, "column length must be nseg * seg_cap");
  if (n / 128 >= (1ll << 31)) return fail(TG_ERR_UNSUPPORTED, "segmented chunk too large");
  if (!j->dev_result) j->dev_result.reset(new ResultBatch());
  ResultBatch& rb = *j->dev_result;
  rb.rows = 0; rb.consumed = 0;
  SegSpec seg{reinterpret_cast<const unsigned long long*>(seg_cnt_dev), (uint32_t)(seg_cap / 128), 0, seg_cap, nullptr};
  TG_CUDA(cudaEventRecord(j->ev0, j->stream));
  TG_TRY(probe_device(j, pview, n, rb, out_rows != nullptr, &seg));
  TG_CUDA(cudaEventRecord(j->ev1, j->stream));
  if (out_rows) {
    TG_CUDA(cudaStreamSynchronize(j->stream));
    float ms = 0; cudaEventElapsedTime(&ms, j->ev0, j->ev1); j->stats.probe_ms += ms;
    *out_rows = rb.rows;
  }
  for (int c = 0; c < j->n_out; c++) {
    if (out_cols) out_cols[c] = rb.cols[c]->p;
    if (out_nulls) out_nulls[c] = rb.bitmaps[c]->p;
  }
  return TG_OK;
}

int tg_join_get_stats(tg_join* h, tg_join_stats* out) {
  TG_LOCK(h);
  if (!out) return fail(TG_ERR_INVALID, "out is NULL");
  TG_CUDA(cudaStreamSynchronize(j->stream));
  *out = j->stats;
  out->d2h_bytes = j->d2h_bytes.load();
  return TG_OK;
}

int tg_join_close(tg_join* h) {
  if (!h) return TG_OK;
  bool was = h->closed.exchange(true);
  if (was) return TG_OK;
  // wake a consumer parked in tg_join_next_wait; notifying under res_mu closes the window between its predicate check
  // and its wait (it holds res_mu across both)
  { std::lock_guard<std::mutex> rlock(h->res_mu); h->res_cv.notify_all(); }
  {
    // waits for an in-flight push and an in-flight next; later calls see `closed`
    std::lock_guard<std::mutex> lock(h->mu);
    std::lock_guard<std::mutex> rlock(h->res_mu);
    JoinImpl* j = h->impl;
    h->impl = nullptr;
    if (j) {
      DeviceGuard g(j->device);
      if (j->stream) cudaStreamSynchronize(j->stream);
      if (j->d2h_stream) { cudaStreamSynchronize(j->d2h_stream); cudaStreamDestroy(j->d2h_stream); }
      j->results.clear();
      j->free_batches.clear();
      j->dev_result.reset();
      if (j->ev0) cudaEventDestroy(j->ev0);
      if (j->ev1) cudaEventDestroy(j->ev1);
      if (j->own_stream && j->stream) cudaStreamDestroy(j->stream);
      cudaGetLastError();
      delete j;
    }
    h->res_cv.notify_all();
  }
  bury_handle(h);   // the shell stays readable for late callers; freed after kGraveyardDepth further closes
  return TG_OK;
}

}  // extern "C"
