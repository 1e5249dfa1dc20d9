// partition.cu — key-hash repartition for the multi-GPU exchange.
//
// The reference's analogue is the MPP ExchangeSender with ExchangeType HashPartition
// (pkg/planner/core/operator/physicalop/physical_exchange_sender.go:115; executed by TiFlash) and, in
// process, partitionHashSplitter.split (pkg/executor/shuffle.go:450).  Here: count rows per destination
// (k_partition_count), then one scatter kernel that regroups each 2048-row tile by destination in shared
// memory and writes every destination's rows as one contiguous run — either into local per-partition
// regions (tg_partition_by_key, followed by an NCCL all-to-all) or straight into the peers' receive
// buffers over NVLink (tg_partition_exchange: the repartition step and its all-to-all in ONE kernel; the
// stores to peer memory are 1 KB runs, and the transfer overlaps the regrouping tile by tile).
#include "partition_kernels.cuh"
#include <mutex>

namespace tg {

static int check_parts(int32_t nparts, int32_t ncols) {
  if (nparts < 1 || nparts > TG_MAX_PARTS) return fail(TG_ERR_UNSUPPORTED, "1..16 partitions");
  if (ncols < 1 || ncols > TG_PART_MAX_COLS) return fail(TG_ERR_UNSUPPORTED, "1..8 columns per repartition call");
  return TG_OK;
}

static int pgrid(int device, int64_t n, int per_block, int per_sm) {
  int64_t need = (n + per_block - 1) / per_block, cap = (int64_t)device_sm_count(device) * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace tg

using namespace tg;

extern "C" {

int32_t tg_partition_of_key(int64_t key, int32_t nparts) { return (int32_t)part_of(hash64((uint64_t)key), (uint32_t)nparts); }

int tg_partition_count(int device, const int64_t* key_dev, int64_t rows, int32_t nparts, int64_t* part_counts_dev, void* stream) {
  TG_TRY(check_parts(nparts, 1));
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)");
  cudaStream_t st = (cudaStream_t)stream;
  TG_CUDA(cudaMemsetAsync(part_counts_dev, 0, (size_t)nparts * 8, st));
  TG_TRY(launch_partition_count<false>(device, st, reinterpret_cast<const long long*>(key_dev), nullptr, rows, (uint32_t)nparts,
                                       reinterpret_cast<unsigned long long*>(part_counts_dev), nullptr));
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

int tg_partition_by_key(int device, const int64_t* key_dev, const uint8_t* key_nulls_dev, int64_t rows, int32_t nparts,
                        int32_t ncols, const void* const* src_cols_dev, void* const* dst_cols_dev,
                        int64_t* part_offsets_dev, void* stream) {
  TG_TRY(check_parts(nparts, ncols));
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf scratch;
  TG_TRY(scratch.ensure(device, (size_t)TG_MAX_PARTS * 8 * 2 + 64));
  unsigned long long* counts = scratch.as<unsigned long long>();
  unsigned long long* cursors = counts + TG_MAX_PARTS;
  TG_CUDA(cudaMemsetAsync(counts, 0, (size_t)TG_MAX_PARTS * 16, st));
  const long long* key = reinterpret_cast<const long long*>(key_dev);
  TG_TRY(launch_partition_count<false>(device, st, key, key_nulls_dev, rows, (uint32_t)nparts, counts, nullptr));
  k_partition_offsets<<<1, 32, 0, st>>>(counts, (uint32_t)nparts, reinterpret_cast<long long*>(part_offsets_dev), cursors);
  PartDst d{};
  d.nparts = nparts; d.ncols = ncols;
  for (int c = 0; c < ncols; c++) { d.src[c] = src_cols_dev[c]; for (int p = 0; p < nparts; p++) d.dst[p][c] = dst_cols_dev[c]; }
  d.dst_base = reinterpret_cast<const long long*>(part_offsets_dev);
  TG_TRY(launch_partition_scatter<false>(device, st, key, key_nulls_dev, rows, d, cursors, nullptr));
  TG_CUDA(cudaGetLastError());
  TG_CUDA(cudaStreamSynchronize(st));   // scratch is freed on return
  return TG_OK;
}

int tg_partition_exchange(int device, const int64_t* key_dev, int64_t rows, int32_t nparts, int32_t ncols,
                          const void* const* src_cols_dev, void* const* recv_cols_peer, const int64_t* part_counts_dev,
                          const int64_t* recv_base_dev, void* stream) {
  (void)part_counts_dev;
  TG_TRY(check_parts(nparts, ncols));
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf scratch;
  TG_TRY(scratch.ensure(device, (size_t)TG_MAX_PARTS * 8 + 64));
  unsigned long long* cursors = scratch.as<unsigned long long>();
  TG_CUDA(cudaMemsetAsync(cursors, 0, (size_t)TG_MAX_PARTS * 8, st));
  PartDst d{};
  d.nparts = nparts; d.ncols = ncols;
  for (int c = 0; c < ncols; c++) { d.src[c] = src_cols_dev[c]; for (int p = 0; p < nparts; p++) d.dst[p][c] = recv_cols_peer[p * ncols + c]; }
  d.dst_base = reinterpret_cast<const long long*>(recv_base_dev);
  TG_TRY(launch_partition_scatter<false>(device, st, reinterpret_cast<const long long*>(key_dev), nullptr, rows, d, cursors, nullptr));
  TG_CUDA(cudaGetLastError());
  TG_CUDA(cudaStreamSynchronize(st));
  return TG_OK;
}

static __global__ void k_zero_cf(unsigned long long* sent, unsigned long long* overflow, long long* bases, long long base) {
  if (threadIdx.x < TG_MAX_PARTS) { sent[threadIdx.x] = 0; bases[threadIdx.x] = base; }
  (void)overflow;   // sticky: zeroed once by the owner, so an overflow of ANY step is still visible when the host looks
}

int tg_partition_exchange_cf(int device, const int64_t* key_dev, int64_t rows, int32_t nparts, int32_t ncols,
                             const void* const* src_cols_dev, void* const* recv_cols_peer, int64_t region_base,
                             int64_t region_cap, int64_t* sent_rows_dev, uint64_t* overflow_dev, void* stream) {
  TG_TRY(check_parts(nparts, ncols));
  if (ncols > 4) return fail(TG_ERR_UNSUPPORTED, "count-free exchange moves at most 4 columns per call");
  if (!scatter_bulk_enabled()) return fail(TG_ERR_UNSUPPORTED, "count-free exchange needs the bulk-store scatter kernel (TG_SCATTER_BULK=0 disables it)");
  if (!sent_rows_dev || !overflow_dev || region_cap <= 0) return fail(TG_ERR_INVALID, "sent_rows_dev / overflow_dev / region_cap are required");
  if (src_cols_dev[0] != (const void*)key_dev) return fail(TG_ERR_INVALID, "src_cols_dev[0] must be the key column");
  for (int c = 0; c < ncols; c++) if (!ptr_aligned16(src_cols_dev[c])) return fail(TG_ERR_UNSUPPORTED, "source columns must be 16-byte aligned");
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)");
  cudaStream_t st = (cudaStream_t)stream;
  // per-device scratch for the (identical) region bases: lives as long as the process, so the call never synchronises
  static std::mutex mu;
  static long long* bases_of[64];
  long long* bases = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!bases_of[device & 63]) TG_CUDA(cudaMalloc(&bases_of[device & 63], TG_MAX_PARTS * 8 * 4));
    bases = bases_of[device & 63];
  }
  // NOTE: `bases` is rewritten per call on the caller's stream; concurrent calls on different streams of one device with
  // different region_base values must not overlap (the exchange is one call per step)
  unsigned long long* cursors = reinterpret_cast<unsigned long long*>(sent_rows_dev);
  k_zero_cf<<<1, 32, 0, st>>>(cursors, reinterpret_cast<unsigned long long*>(overflow_dev), bases, (long long)region_base);
  PartDst d{};
  d.nparts = nparts; d.ncols = ncols;
  for (int c = 0; c < ncols; c++) { d.src[c] = src_cols_dev[c]; for (int p = 0; p < nparts; p++) d.dst[p][c] = recv_cols_peer[p * ncols + c]; }
  d.dst_base = bases; d.capacity = region_cap; d.overflow = reinterpret_cast<unsigned long long*>(overflow_dev);
  const int64_t TILE = 1024;
  const int64_t n_main = rows / TILE * TILE;
  if (n_main > 0) TG_TRY(launch_partition_scatter<false>(device, st, reinterpret_cast<const long long*>(key_dev), nullptr, n_main, d, cursors, nullptr));
  if (n_main < rows) {
    // the last < 1024 rows: LSU kernel, one CTA, same cursors and capacity
    PartDst tail = d;
    for (int c = 0; c < ncols; c++) tail.src[c] = reinterpret_cast<const unsigned long long*>(src_cols_dev[c]) + n_main;
    k_partition_scatter<false><<<1, PT_BLOCK, 0, st>>>(reinterpret_cast<const long long*>(tail.src[0]), nullptr, rows - n_main, tail, cursors);
  }
  TG_CUDA(cudaGetLastError());
  return TG_OK;
}

}  // extern "C"
