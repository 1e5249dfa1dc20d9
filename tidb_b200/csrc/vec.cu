// vec.cu — VecEval* kernels: column-at-a-time scalar builtins over 8-byte chunk columns.
//
// Replaces (pkg/expression): builtinLTIntSig.vecEvalInt & the other five integer comparisons
// (builtin_compare_vec.go:524-561, vecCompareInt :619), the real comparisons
// (builtin_compare_vec_generated.go:54), builtinArithmetic{Plus,Minus,Multiply}{Int,Real}Sig
// (builtin_arithmetic_vec.go:856, :365, :646, :1011, :496, :300, :40) and
// expression.VectorizedFilter (chunk_executor.go:413).  All are streaming kernels bounded by HBM:
// 8 B (+8 B) read and 8 B written per row, null bitmaps merged bytewise (Column.MergeNulls column.go:906).
#include <memory>
#include "common.cuh"

namespace tg {

struct VArg { const void* data; const uint8_t* nulls; };

// Streaming layout: lane l of a warp owns rows base+l, base+32+l, ... (VEC_ITEMS per thread, all loads issued before
// use) so every load/store instruction is one fully coalesced 256-byte request; the 32 validity bits of a warp-row are
// one ballot, written as one aligned 32-bit word of the result bitmap.
#define VEC_ITEMS 4
__device__ __forceinline__ bool arg_valid(const VArg& a, int64_t i) { return !a.nulls || bit_not_null(a.nulls, i); }

// write the 32 validity bits of rows [wbase, wbase+32) (wbase % 32 == 0); tail rows are masked off
__device__ __forceinline__ void store_valid_word(uint8_t* rnulls, int64_t wbase, int64_t n, unsigned bal, int lane) {
  if (lane == 0 && wbase < n) {
    int64_t rem = n - wbase;
    if (rem >= 32) *reinterpret_cast<uint32_t*>(rnulls + (wbase >> 3)) = bal;
    else {
      bal &= (1u << rem) - 1u;
      for (int b = 0; b < (int)((rem + 7) / 8); b++) rnulls[(wbase >> 3) + b] = (uint8_t)(bal >> (8 * b));
    }
  }
}

template <bool REAL>
__global__ void __launch_bounds__(256)
k_vec_compare(int op, int ua, int ub, VArg a, VArg b, long long bc_i, double bc_f, int64_t n,
              long long* __restrict__ result, uint8_t* __restrict__ rnulls) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t tile = 32 * VEC_ITEMS;
  for (int64_t base = warp * tile; base < n; base += nwarps * tile) {
    unsigned long long x[VEC_ITEMS], y[VEC_ITEMS];
    bool in[VEC_ITEMS];
#pragma unroll
    for (int j = 0; j < VEC_ITEMS; j++) {
      int64_t i = base + j * 32 + lane;
      in[j] = i < n;
      x[j] = in[j] ? __ldcs(reinterpret_cast<const unsigned long long*>(a.data) + i) : 0ull;
      y[j] = (in[j] && b.data) ? __ldcs(reinterpret_cast<const unsigned long long*>(b.data) + i) : 0ull;
    }
#pragma unroll
    for (int j = 0; j < VEC_ITEMS; j++) {
      int64_t i = base + j * 32 + lane;
      bool valid = in[j] && arg_valid(a, i) && arg_valid(b, i);
      int c;
      if (REAL) c = cmp_real(__longlong_as_double((long long)x[j]), b.data ? __longlong_as_double((long long)y[j]) : bc_f);
      else c = cmp_int((long long)x[j], ua != 0, b.data ? (long long)y[j] : bc_i, ub != 0);
      if (in[j]) __stcs(result + i, apply_cmp(op, c) ? 1ll : 0ll);
      unsigned bal = __ballot_sync(0xffffffffu, valid);
      store_valid_word(rnulls, base + j * 32, n, bal, lane);
    }
  }
}

// builtinArithmeticMinusIntSig.overflowCheck (builtin_arithmetic.go:491), forceToSigned = false
__device__ __forceinline__ bool minus_overflow(bool lu, bool ru, long long a, long long b) {
  bool is_signed = !lu && !ru;
  long long res = (long long)((unsigned long long)a - (unsigned long long)b);
  unsigned long long ua = (unsigned long long)a, ub = (unsigned long long)b;
  bool resUnsigned = false;
  if (lu) {
    if (ru) { if (ua < ub) { if (res >= 0) return true; } else resUnsigned = true; }
    else {
      if (b >= 0) { if (ua > ub) resUnsigned = true; }
      else { if (~0ull - ua < (unsigned long long)(-(unsigned long long)b)) return true; resUnsigned = true; }
    }
  } else {
    if (ru) { if ((unsigned long long)a - 0x8000000000000000ull < ub) return true; }
    else { if (a > 0 && b < 0) resUnsigned = true; else if (a < 0 && b > 0 && res >= 0) return true; }
  }
  return (!is_signed && !resUnsigned && res < 0) || (is_signed && resUnsigned && (unsigned long long)res > 0x7fffffffffffffffull);
}

__global__ void __launch_bounds__(256)
k_vec_arith_int(int op, int lu, int ru, VArg a, VArg b, long long bc, int64_t n, long long* __restrict__ result,
                uint8_t* __restrict__ rnulls, int* overflow) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t tile = 32 * VEC_ITEMS;
  const long long MAXI = 0x7fffffffffffffffll, MINI = (long long)0x8000000000000000ull;
  for (int64_t base = warp * tile; base < n; base += nwarps * tile) {
    long long xs[VEC_ITEMS], ys[VEC_ITEMS];
    bool in[VEC_ITEMS];
#pragma unroll
    for (int j = 0; j < VEC_ITEMS; j++) {
      int64_t i = base + j * 32 + lane;
      in[j] = i < n;
      xs[j] = in[j] ? __ldcs(reinterpret_cast<const long long*>(a.data) + i) : 0ll;
      ys[j] = in[j] ? (b.data ? __ldcs(reinterpret_cast<const long long*>(b.data) + i) : bc) : 0ll;
    }
#pragma unroll
    for (int j = 0; j < VEC_ITEMS; j++) {
      int64_t i = base + j * 32 + lane;
      long long lh = xs[j], rh = ys[j];
      bool ovf;
      long long r;
      if (op == TG_ARITH_PLUS) {
        if (lu && ru) ovf = (unsigned long long)lh > ~0ull - (unsigned long long)rh;
        else if (lu && !ru) ovf = (rh < 0 && (unsigned long long)(-(unsigned long long)rh) > (unsigned long long)lh) || (rh > 0 && (unsigned long long)lh > ~0ull - (unsigned long long)rh);
        else if (!lu && ru) ovf = (lh < 0 && (unsigned long long)(-(unsigned long long)lh) > (unsigned long long)rh) || (lh > 0 && (unsigned long long)rh > ~0ull - (unsigned long long)lh);
        else ovf = (lh > 0 && rh > MAXI - lh) || (lh < 0 && rh < MINI - lh);
        r = (long long)((unsigned long long)lh + (unsigned long long)rh);
      } else if (op == TG_ARITH_MINUS) {
        ovf = minus_overflow(lu != 0, ru != 0, lh, rh);
        r = (long long)((unsigned long long)lh - (unsigned long long)rh);
      } else if (lu || ru) {
        unsigned long long x = (unsigned long long)lh, y = (unsigned long long)rh, res = x * y;
        ovf = x != 0 && res / x != y;
        r = (long long)res;
      } else {
        long long tmp = (long long)((unsigned long long)lh * (unsigned long long)rh);
        bool special = tmp == MINI && lh == -1;
        ovf = special || (lh != 0 && tmp / lh != rh);
        r = tmp;
      }
      bool valid = in[j] && arg_valid(a, i) && arg_valid(b, i);
      if (ovf) { if (valid) atomicExch(overflow, 1); r = 0; }
      if (in[j]) __stcs(result + i, r);
      unsigned bal = __ballot_sync(0xffffffffu, valid);
      store_valid_word(rnulls, base + j * 32, n, bal, lane);
    }
  }
}

__global__ void __launch_bounds__(256)
k_vec_arith_real(int op, VArg a, VArg b, double bc, int64_t n, double* __restrict__ result, uint8_t* __restrict__ rnulls,
                 int* overflow) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t tile = 32 * VEC_ITEMS;
  for (int64_t base = warp * tile; base < n; base += nwarps * tile) {
    double xs[VEC_ITEMS], ys[VEC_ITEMS];
    bool in[VEC_ITEMS];
#pragma unroll
    for (int j = 0; j < VEC_ITEMS; j++) {
      int64_t i = base + j * 32 + lane;
      in[j] = i < n;
      xs[j] = in[j] ? __ldcs(reinterpret_cast<const double*>(a.data) + i) : 0.0;
      ys[j] = in[j] ? (b.data ? __ldcs(reinterpret_cast<const double*>(b.data) + i) : bc) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < VEC_ITEMS; j++) {
      int64_t i = base + j * 32 + lane;
      double r;
      bool ovf;
      if (op == TG_ARITH_PLUS) { r = xs[j] + ys[j]; ovf = !isfinite(r); }          // mathutil.IsFinite
      else if (op == TG_ARITH_MINUS) { r = xs[j] - ys[j]; ovf = !isfinite(r); }
      else { r = xs[j] * ys[j]; ovf = isinf(r); }                                    // math.IsInf only (:51-58)
      bool valid = in[j] && arg_valid(a, i) && arg_valid(b, i);
      if (ovf && valid) atomicExch(overflow, 1);
      if (in[j]) __stcs(result + i, r);
      unsigned bal = __ballot_sync(0xffffffffu, valid);
      store_valid_word(rnulls, base + j * 32, n, bal, lane);
    }
  }
}

// selected[physical row] = every CNF item is non-NULL true and the row is in sel (if any)
__global__ void __launch_bounds__(256)
k_vec_filter(DevCols cols, DevFilter f, const long long* __restrict__ sel, int64_t nsel, int64_t nphys,
             uint8_t* __restrict__ selected, unsigned long long* count) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  int64_t n = sel ? nsel : nphys;
  for (; i < n; i += stride) {
    int64_t p = sel ? sel[i] : i;
    bool s = eval_filter(f, cols, p);
    selected[p] = s ? 1 : 0;
    local += s;
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// a column argument made device-resident (copies host buffers when on_device == 0)
struct ArgDev {
  DevBuf data, nulls;
  VArg v{nullptr, nullptr};
  int load(int device, int on_device, const tg_column* c, cudaStream_t st) {
    if (!c) return TG_OK;
    if (c->elem_len != 8) return fail(TG_ERR_UNSUPPORTED, "VecEval kernels take 8-byte columns");
    if (on_device) { v.data = c->data; v.nulls = c->null_bitmap; return TG_OK; }
    size_t bytes = (size_t)c->length * 8, nb = (size_t)((c->length + 7) / 8);
    TG_TRY(data.ensure(device, bytes + 16));
    if (bytes) TG_CUDA(cudaMemcpyAsync(data.p, c->data, bytes, cudaMemcpyHostToDevice, st));
    v.data = data.p;
    if (c->null_bitmap) {
      TG_TRY(nulls.ensure(device, nb + 16));
      if (nb) TG_CUDA(cudaMemcpyAsync(nulls.p, c->null_bitmap, nb, cudaMemcpyHostToDevice, st));
      v.nulls = nulls.as<uint8_t>();
    }
    return TG_OK;
  }
};

static int vgrid(int device, int64_t n_threads) {
  int64_t need = (n_threads + 255) / 256, cap = (int64_t)device_sm_count(device) * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

template <typename Launch>
static int run_binary(int device, int on_device, const tg_column* a, const tg_column* b, void* result, uint8_t* rnulls,
                      void* stream, bool has_overflow, Launch launch) {
  if (!a || !result || !rnulls) return fail(TG_ERR_INVALID, "a / result / result_nulls is NULL");
  if (b && b->length != a->length) return fail(TG_ERR_INVALID, "argument columns have different lengths");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(TG_ERR_CUDA, "no CUDA device: VecEval kernels have no CPU fallback"); }
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n = a->length;
  ArgDev da, db;
  TG_TRY(da.load(device, on_device, a, st));
  TG_TRY(db.load(device, on_device, b, st));
  DevBuf dres, dnul;
  static std::mutex flag_mu;
  static int* ovf_flag[16];   // one overflow flag per device, allocated once and intentionally never freed
  std::lock_guard<std::mutex> flag_lock(flag_mu);   // VecEval calls are serialised per process
  int*& dovf = ovf_flag[device & 15];
  void* res_dev = result; uint8_t* nul_dev = rnulls;
  size_t nb = (size_t)((n + 7) / 8);
  if (!on_device) {
    TG_TRY(dres.ensure(device, (size_t)n * 8 + 16)); TG_TRY(dnul.ensure(device, nb + 16));
    res_dev = dres.p; nul_dev = dnul.as<uint8_t>();
  }
  if (!dovf) TG_CUDA(cudaMalloc(reinterpret_cast<void**>(&dovf), 16));
  TG_CUDA(cudaMemsetAsync(dovf, 0, 4, st));
  if (n > 0) launch(vgrid(device, (n + VEC_ITEMS - 1) / VEC_ITEMS), st, da.v, db.v, n, res_dev, nul_dev, dovf);
  int ovf = 0;
  if (has_overflow) TG_CUDA(cudaMemcpyAsync(&ovf, dovf, 4, cudaMemcpyDeviceToHost, st));
  if (!on_device && n > 0) {
    TG_CUDA(cudaMemcpyAsync(result, res_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaMemcpyAsync(rnulls, nul_dev, nb, cudaMemcpyDeviceToHost, st));
  }
  TG_CUDA(cudaStreamSynchronize(st));
  TG_CUDA(cudaGetLastError());
  if (ovf) return fail(TG_ERR_OVERFLOW, "ErrOverflow: value is out of range in arithmetic VecEval kernel");
  return TG_OK;
}

}  // namespace tg

using namespace tg;

extern "C" {

int tg_vec_compare_int(int device, int on_device, int op, int a_unsigned, int b_unsigned, const tg_column* a,
                       const tg_column* b, int64_t b_const, int64_t* result, uint8_t* result_nulls, void* stream) {
  return run_binary(device, on_device, a, b, result, result_nulls, stream, false,
                    [&](int grid, cudaStream_t st, VArg va, VArg vb, int64_t n, void* r, uint8_t* rn, int*) {
                      k_vec_compare<false><<<grid, 256, 0, st>>>(op, a_unsigned, b_unsigned, va, vb, (long long)b_const, 0.0, n,
                                                                 reinterpret_cast<long long*>(r), rn);
                    });
}

int tg_vec_compare_real(int device, int on_device, int op, const tg_column* a, const tg_column* b, double b_const,
                        int64_t* result, uint8_t* result_nulls, void* stream) {
  return run_binary(device, on_device, a, b, result, result_nulls, stream, false,
                    [&](int grid, cudaStream_t st, VArg va, VArg vb, int64_t n, void* r, uint8_t* rn, int*) {
                      k_vec_compare<true><<<grid, 256, 0, st>>>(op, 0, 0, va, vb, 0, b_const, n, reinterpret_cast<long long*>(r), rn);
                    });
}

int tg_vec_arith_int(int device, int on_device, int op, int a_unsigned, int b_unsigned, const tg_column* a,
                     const tg_column* b, int64_t b_const, int64_t* result, uint8_t* result_nulls, void* stream) {
  return run_binary(device, on_device, a, b, result, result_nulls, stream, true,
                    [&](int grid, cudaStream_t st, VArg va, VArg vb, int64_t n, void* r, uint8_t* rn, int* ovf) {
                      k_vec_arith_int<<<grid, 256, 0, st>>>(op, a_unsigned, b_unsigned, va, vb, (long long)b_const, n,
                                                            reinterpret_cast<long long*>(r), rn, ovf);
                    });
}

int tg_vec_arith_real(int device, int on_device, int op, const tg_column* a, const tg_column* b, double b_const,
                      double* result, uint8_t* result_nulls, void* stream) {
  return run_binary(device, on_device, a, b, result, result_nulls, stream, true,
                    [&](int grid, cudaStream_t st, VArg va, VArg vb, int64_t n, void* r, uint8_t* rn, int* ovf) {
                      k_vec_arith_real<<<grid, 256, 0, st>>>(op, va, vb, b_const, n, reinterpret_cast<double*>(r), rn, ovf);
                    });
}

int tg_vec_filter(int device, int on_device, const tg_chunk* chk, const tg_filter_item* items, int32_t n_items,
                  uint8_t* selected, int64_t* n_selected, void* stream) {
  if (!chk || !selected) return fail(TG_ERR_INVALID, "chunk / selected is NULL");
  if (n_items < 0 || n_items > TG_MAX_FILTER) return fail(TG_ERR_UNSUPPORTED, "at most 8 CNF filter items are offloaded");
  if (chk->ncols <= 0 || chk->ncols > TG_MAX_COLS) return fail(TG_ERR_UNSUPPORTED, "chunk must have 1..16 columns");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(TG_ERR_CUDA, "no CUDA device: VecEval kernels have no CPU fallback"); }
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  DevFilter f{}; f.n = n_items;
  std::vector<char> needed(chk->ncols, 0);
  for (int i = 0; i < n_items; i++) {
    const tg_filter_item& it = items[i];
    if (it.lhs_col < 0 || it.lhs_col >= chk->ncols || it.rhs_col >= chk->ncols) return fail(TG_ERR_INVALID, "filter column out of range");
    f.items[i] = it; needed[it.lhs_col] = 1; if (it.rhs_col >= 0) needed[it.rhs_col] = 1;
  }
  int64_t nphys = chk->cols[0].length;
  std::vector<std::unique_ptr<ArgDev>> args;
  DevCols cols{};
  for (int c = 0; c < chk->ncols; c++) {
    args.emplace_back(new ArgDev());
    cols.elem_len[c] = chk->cols[c].elem_len;
    if (!needed[c]) continue;
    TG_TRY(args[c]->load(device, on_device, &chk->cols[c], st));
    cols.data[c] = args[c]->v.data; cols.nulls[c] = args[c]->v.nulls;
  }
  DevBuf dsel_idx, dselected, dcount;
  const long long* sel_dev = reinterpret_cast<const long long*>(chk->sel);
  uint8_t* selected_dev = selected;
  if (!on_device) {
    if (chk->sel) {
      TG_TRY(dsel_idx.ensure(device, (size_t)chk->nsel * 8 + 16));
      TG_CUDA(cudaMemcpyAsync(dsel_idx.p, chk->sel, (size_t)chk->nsel * 8, cudaMemcpyHostToDevice, st));
      sel_dev = dsel_idx.as<long long>();
    }
    TG_TRY(dselected.ensure(device, (size_t)nphys + 16));
    selected_dev = dselected.as<uint8_t>();
  }
  TG_TRY(dcount.ensure(device, 16));
  TG_CUDA(cudaMemsetAsync(dcount.p, 0, 8, st));
  TG_CUDA(cudaMemsetAsync(selected_dev, 0, (size_t)nphys, st));
  int64_t n = chk->sel ? chk->nsel : nphys;
  if (n > 0) k_vec_filter<<<vgrid(device, n), 256, 0, st>>>(cols, f, sel_dev, chk->nsel, nphys, selected_dev, dcount.as<unsigned long long>());
  unsigned long long cnt = 0;
  TG_CUDA(cudaMemcpyAsync(&cnt, dcount.p, 8, cudaMemcpyDeviceToHost, st));
  if (!on_device && nphys) TG_CUDA(cudaMemcpyAsync(selected, selected_dev, (size_t)nphys, cudaMemcpyDeviceToHost, st));
  TG_CUDA(cudaStreamSynchronize(st));
  TG_CUDA(cudaGetLastError());
  if (n_selected) *n_selected = (int64_t)cnt;
  return TG_OK;
}

}  // extern "C"
