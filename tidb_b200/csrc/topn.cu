// topn.cu — tg_topn: TopNExec (pkg/executor/sortexec/topn.go:74 TopNExec, :230 executeTopN, :325 processChildChk) on the
// device.  The reference keeps a heap of offset+count rows and compares rows with the ORDER BY items' CompareFuncs
// (chunk.GetCompareFunc: NULL sorts before every value, DESC negates).  Here:
//   1. k_topn_rank   : one 64-bit RANK per row from the FIRST item (order-preserving map of the value, inverted for DESC,
//                      NULL = smallest / largest) — smaller rank = earlier in the output;
//   2. radix select  : 8 histogram passes (k_topn_hist, 8 bits each, most significant first) find the rank of the
//                      (offset+count)-th row without sorting anything;
//   3. k_topn_collect: rows whose rank is <= that threshold (>= offset+count rows; more only on ties of the first item)
//                      are compacted, their columns gathered (k_topn_gather) and copied to the host;
//   4. host          : the few candidates are sorted with the full multi-item comparator and rows [offset, offset+count)
//                      are returned.  Ties are broken arbitrarily, as by the reference's heap.
// HBM-bound: the rank pass reads 8 B/row + bitmap, each histogram pass 8 B/row.
#include <algorithm>
#include <memory>
#include <vector>
#include "common.cuh"

namespace tg {

__device__ __forceinline__ unsigned long long rank_of(unsigned long long raw, bool is_null, int kind /*0 signed, 1 unsigned, 2 real*/, bool desc) {
  unsigned long long o;
  if (is_null) o = 0ull;                         // NULL sorts before every value (chunk.GetCompareFunc -> cmpNull)
  else if (kind == 2) { o = (raw >> 63) ? ~raw : (raw | 0x8000000000000000ull); }
  else if (kind == 1) o = raw;
  else o = raw ^ 0x8000000000000000ull;
  // NULL and the smallest value may share rank 0 (and, inverted, the largest): that only widens the candidate set; the
  // final order comes from the exact comparator on the host
  return desc ? ~o : o;
}

__global__ void __launch_bounds__(256)
k_topn_rank(const unsigned long long* __restrict__ data, const uint8_t* __restrict__ nulls, int64_t n, int kind, int desc,
            unsigned long long* __restrict__ rank) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    bool isn = nulls && !bit_not_null(nulls, i);
    unsigned long long raw = __ldcs(data + i);
    if (kind == 2 && !isn) { double d = __longlong_as_double((long long)raw); if (d != d) raw = 0xFFF8000000000000ull; }   // NaN below everything (Go cmp.Compare)
    rank[i] = rank_of(raw, isn, kind, desc != 0);
  }
}

// histogram of byte `shift/8` over the rows whose higher bytes equal `prefix`
__global__ void __launch_bounds__(256)
k_topn_hist(const unsigned long long* __restrict__ rank, int64_t n, unsigned long long prefix, int shift, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const unsigned long long himask = shift >= 56 ? 0ull : (~0ull << (shift + 8));
  for (; i < n; i += stride) {
    unsigned long long r = rank[i];
    if ((r & himask) == (prefix & himask)) atomicAdd(&s_h[(r >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_h[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
k_topn_collect(const unsigned long long* __restrict__ rank, int64_t n, unsigned long long threshold, unsigned long long cap,
               unsigned long long* __restrict__ cursor, long long* __restrict__ idx) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    if (rank[i] <= threshold) {
      unsigned long long o = atomicAdd(cursor, 1ull);
      if (o < cap) idx[o] = i;
    }
  }
}

__global__ void __launch_bounds__(256)
k_topn_gather(const unsigned long long* __restrict__ data, const uint8_t* __restrict__ nulls, const long long* __restrict__ idx, int64_t m,
              unsigned long long* __restrict__ out, uint8_t* __restrict__ out_valid) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < m; i += stride) {
    const long long r = idx[i];
    out[i] = data[r];
    out_valid[i] = (nulls && !bit_not_null(nulls, r)) ? 0 : 1;
  }
}

static int kind_of(int tp, uint32_t flag) {
  if (tp == TG_TYPE_DOUBLE) return 2;
  if (is_int_family(tp)) return (flag & TG_FLAG_UNSIGNED) ? 1 : 0;
  if (tp == TG_TYPE_DATE || tp == TG_TYPE_DATETIME || tp == TG_TYPE_TIMESTAMP) return 1;   // packed CoreTime compares as uint64 (types/time.go:646)
  return -1;
}

}  // namespace tg

using namespace tg;

extern "C" {

int tg_topn(int device, int on_device, const tg_chunk* chk, const int32_t* col_types, const uint32_t* col_flags,
            const tg_sort_item* items, int32_t n_items, int64_t offset, int64_t count, tg_mut_chunk* out, int64_t* nrows, void* stream) {
  if (!chk || !col_types || !items || !out || !nrows) return fail(TG_ERR_INVALID, "chk / col_types / items / out / nrows is NULL");
  *nrows = 0;
  if (n_items < 1 || n_items > 8) return fail(TG_ERR_UNSUPPORTED, "1..8 ORDER BY items are offloaded");
  if (offset < 0 || count < 0) return fail(TG_ERR_INVALID, "negative offset / count");
  if (chk->sel) return fail(TG_ERR_UNSUPPORTED, "TopN input must not carry a sel vector");
  const int nc = chk->ncols;
  if (nc < 1 || nc > TG_MAX_COLS || out->ncols != nc) return fail(TG_ERR_INVALID, "1..16 columns; the output chunk has the child's schema");
  std::vector<int> kinds(nc);
  for (int c = 0; c < nc; c++) {
    if (chk->cols[c].elem_len != 8) return fail(TG_ERR_UNSUPPORTED, "TopN is offloaded for 8-byte columns only");
    kinds[c] = kind_of(col_types[c], col_flags ? col_flags[c] : 0);
  }
  for (int q = 0; q < n_items; q++) {
    if (items[q].col < 0 || items[q].col >= nc) return fail(TG_ERR_INVALID, "ORDER BY column out of range");
    if (kinds[items[q].col] < 0) return fail(TG_ERR_UNSUPPORTED, "ORDER BY column type is not offloaded (int family / double / time)");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(TG_ERR_CUDA, "no CUDA device: TopN has no CPU fallback"); }
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = chk->cols[0].length;
  const int64_t want = std::min<int64_t>(n, offset + count);
  if (want <= offset || n == 0 || count == 0) return TG_OK;
  // device-resident columns
  std::vector<std::unique_ptr<DevBuf>> hold;
  std::vector<const unsigned long long*> dcol(nc);
  std::vector<const uint8_t*> dnul(nc, nullptr);
  for (int c = 0; c < nc; c++) {
    if (chk->cols[c].length != n) return fail(TG_ERR_INVALID, "chunk columns have different lengths");
    if (on_device) { dcol[c] = reinterpret_cast<const unsigned long long*>(chk->cols[c].data); dnul[c] = chk->cols[c].null_bitmap; continue; }
    hold.emplace_back(new DevBuf());
    TG_TRY(hold.back()->ensure(device, (size_t)n * 8 + 16));
    TG_CUDA(cudaMemcpyAsync(hold.back()->p, chk->cols[c].data, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    dcol[c] = hold.back()->as<unsigned long long>();
    if (chk->cols[c].null_bitmap) {
      hold.emplace_back(new DevBuf());
      size_t nb = (size_t)((n + 7) / 8);
      TG_TRY(hold.back()->ensure(device, nb + 16));
      TG_CUDA(cudaMemcpyAsync(hold.back()->p, chk->cols[c].null_bitmap, nb, cudaMemcpyHostToDevice, st));
      dnul[c] = hold.back()->as<uint8_t>();
    }
  }
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)device_sm_count(device) * 8);
  DevBuf rank, scratch, idx;
  TG_TRY(rank.ensure(device, (size_t)n * 8 + 16));
  TG_TRY(scratch.ensure(device, 257 * 8));
  const int c0 = items[0].col;
  k_topn_rank<<<grid, 256, 0, st>>>(dcol[c0], dnul[c0], n, kinds[c0], items[0].desc, rank.as<unsigned long long>());
  // radix select: the rank of the `want`-th smallest row
  unsigned long long prefix = 0, remaining = (unsigned long long)want;
  unsigned long long hist[256];
  for (int shift = 56; shift >= 0; shift -= 8) {
    TG_CUDA(cudaMemsetAsync(scratch.p, 0, 256 * 8, st));
    k_topn_hist<<<grid, 256, 0, st>>>(rank.as<unsigned long long>(), n, prefix, shift, scratch.as<unsigned long long>());
    TG_CUDA(cudaMemcpyAsync(hist, scratch.p, 256 * 8, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    int b = 0;
    for (; b < 256; b++) { if (hist[b] >= remaining) break; remaining -= hist[b]; }
    if (b == 256) return fail(TG_ERR_CUDA, "internal: TopN radix select ran out of rows");
    prefix |= (unsigned long long)b << shift;
  }
  // candidates: every row at or below the threshold rank
  unsigned long long* cursor = scratch.as<unsigned long long>() + 256;
  unsigned long long cap = (unsigned long long)want + 65536;
  unsigned long long m = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    TG_TRY(idx.ensure(device, (size_t)cap * 8 + 16));
    TG_CUDA(cudaMemsetAsync(cursor, 0, 8, st));
    k_topn_collect<<<grid, 256, 0, st>>>(rank.as<unsigned long long>(), n, prefix, cap, cursor, idx.as<long long>());
    TG_CUDA(cudaMemcpyAsync(&m, cursor, 8, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    if (m <= cap) break;
    cap = m;   // many rows tie with the threshold on the first item: take them all, the host comparator decides
  }
  // gather the candidates' columns and bring them to the host
  std::vector<std::vector<unsigned long long>> hv(nc, std::vector<unsigned long long>((size_t)m));
  std::vector<std::vector<uint8_t>> hn(nc, std::vector<uint8_t>((size_t)m));
  DevBuf gcol, gval;
  TG_TRY(gcol.ensure(device, (size_t)m * 8 + 16));
  TG_TRY(gval.ensure(device, (size_t)m + 16));
  const int ggrid = (int)std::min<int64_t>(((int64_t)m + 255) / 256, (int64_t)device_sm_count(device) * 8);
  for (int c = 0; c < nc; c++) {
    k_topn_gather<<<ggrid, 256, 0, st>>>(dcol[c], dnul[c], idx.as<long long>(), (int64_t)m, gcol.as<unsigned long long>(), gval.as<uint8_t>());
    TG_CUDA(cudaMemcpyAsync(hv[c].data(), gcol.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaMemcpyAsync(hn[c].data(), gval.p, (size_t)m, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
  }
  TG_CUDA(cudaGetLastError());
  // exact multi-item comparator (sortexec compareRow: per item CompareFunc, NULL first, DESC negated)
  std::vector<int64_t> order((size_t)m);
  for (size_t i = 0; i < order.size(); i++) order[i] = (int64_t)i;
  auto cmp_item = [&](int q, int64_t a, int64_t b) -> int {
    const int c = items[q].col;
    const bool an = !hn[c][(size_t)a], bn = !hn[c][(size_t)b];
    int r;
    if (an || bn) r = an == bn ? 0 : (an ? -1 : 1);
    else {
      const unsigned long long x = hv[c][(size_t)a], y = hv[c][(size_t)b];
      if (kinds[c] == 2) {
        double dx, dy; std::memcpy(&dx, &x, 8); std::memcpy(&dy, &y, 8);
        const bool xn = dx != dx, yn = dy != dy;
        r = xn ? (yn ? 0 : -1) : (yn ? 1 : (dx < dy ? -1 : (dx > dy ? 1 : 0)));
      } else if (kinds[c] == 1) r = x < y ? -1 : (x > y ? 1 : 0);
      else r = (long long)x < (long long)y ? -1 : ((long long)x > (long long)y ? 1 : 0);
    }
    return items[q].desc ? -r : r;
  };
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    for (int q = 0; q < n_items; q++) { int r = cmp_item(q, a, b); if (r) return r < 0; }
    return false;
  });
  const int64_t take = std::min<int64_t>(want - offset, std::min<int64_t>(out->capacity_rows, (int64_t)m - offset));
  if (want - offset > out->capacity_rows) return fail(TG_ERR_CAPACITY, "TopN output chunk is smaller than `count`");
  for (int c = 0; c < nc; c++) {
    if (out->cols[c].elem_len != 8) return fail(TG_ERR_INVALID, "output column elem_len mismatch");
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(out->cols[c].data);
    uint8_t* nb = out->cols[c].null_bitmap;
    if (nb) std::memset(nb, 0, (size_t)((take + 7) / 8));
    for (int64_t i = 0; i < take; i++) {
      const int64_t r = order[(size_t)(offset + i)];
      const bool valid = hn[c][(size_t)r] != 0;
      dst[i] = valid ? hv[c][(size_t)r] : 0ull;
      if (nb) { if (valid) nb[i >> 3] |= (uint8_t)(1u << (i & 7)); }
      else if (!valid) return fail(TG_ERR_INVALID, "output column can be NULL but the caller passed no null bitmap");
    }
  }
  *nrows = take;
  return TG_OK;
}

}  // extern "C"
