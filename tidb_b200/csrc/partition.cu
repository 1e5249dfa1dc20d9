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
#include "common.cuh"

namespace tg {

#define TG_MAX_PARTS 16
#define TG_PART_MAX_COLS 8
#define PT_BLOCK 256
#define PT_ITEMS 8
#define PT_TILE (PT_BLOCK * PT_ITEMS)

struct PartDst {
  int32_t nparts, ncols;
  const void* src[TG_PART_MAX_COLS];
  void* dst[TG_MAX_PARTS][TG_PART_MAX_COLS];   // column base per destination
  // row offset inside the destination buffers where this launch starts writing, per destination
  const long long* dst_base;                   // device array [nparts]
};

__device__ __forceinline__ uint32_t row_part(const long long* key, const uint8_t* nulls, int64_t i, uint32_t nparts) {
  if (nulls && !bit_not_null(nulls, i)) return part_of(mix64((uint64_t)i), nparts);   // NULL keys never join: spread them
  return part_of(mix64((uint64_t)key[i]), nparts);
}

__global__ void __launch_bounds__(256)
k_partition_count(const long long* __restrict__ key, const uint8_t* __restrict__ nulls, int64_t n, uint32_t nparts,
                  unsigned long long* __restrict__ counts) {
  __shared__ unsigned long long s_cnt[TG_MAX_PARTS];
  if (threadIdx.x < TG_MAX_PARTS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned int local[TG_MAX_PARTS];
#pragma unroll
  for (int p = 0; p < TG_MAX_PARTS; p++) local[p] = 0;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t p = row_part(key, nulls, i, nparts);
#pragma unroll
    for (int q = 0; q < TG_MAX_PARTS; q++) local[q] += (p == (uint32_t)q);
  }
#pragma unroll
  for (int p = 0; p < TG_MAX_PARTS; p++) {
    unsigned int v = local[p];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[p], (unsigned long long)v);
  }
  __syncthreads();
  if (threadIdx.x < nparts && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

// exclusive prefix of counts → part_offsets[nparts+1]; also seeds the scatter cursors
__global__ void k_partition_offsets(const unsigned long long* counts, uint32_t nparts, long long* part_offsets,
                                    unsigned long long* cursors) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (uint32_t p = 0; p < nparts; p++) { part_offsets[p] = run; cursors[p] = 0; run += (long long)counts[p]; }
    part_offsets[nparts] = run;
  }
}

__global__ void __launch_bounds__(PT_BLOCK)
k_partition_scatter(const long long* __restrict__ key, const uint8_t* __restrict__ nulls, int64_t n, PartDst d,
                    unsigned long long* __restrict__ cursors) {
  __shared__ unsigned long long s_val[PT_TILE];
  __shared__ uint32_t s_cnt[TG_MAX_PARTS], s_off[TG_MAX_PARTS + 1];
  __shared__ unsigned long long s_gbase[TG_MAX_PARTS];
  const int lane = threadIdx.x & 31;
  const uint32_t P = (uint32_t)d.nparts;
  const int64_t ntiles = (n + PT_TILE - 1) / PT_TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * PT_TILE;
    if (threadIdx.x < TG_MAX_PARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    // phase 1: destination of every row + rank inside (tile, destination), warp-aggregated
    uint32_t part[PT_ITEMS], rank[PT_ITEMS];
#pragma unroll
    for (int j = 0; j < PT_ITEMS; j++) {
      int64_t i = base + (int64_t)j * PT_BLOCK + threadIdx.x;
      bool in = i < n;
      uint32_t p = in ? row_part(key, nulls, i, P) : 0xffffffffu;
      unsigned peers = __match_any_sync(0xffffffffu, p);
      uint32_t r = 0;
      if (in) {
        int leader = __ffs(peers) - 1;
        uint32_t wbase = 0;
        if (lane == leader) wbase = atomicAdd(&s_cnt[p], (uint32_t)__popc(peers));
        wbase = __shfl_sync(peers, wbase, leader);
        r = wbase + __popc(peers & ((1u << lane) - 1));
      }
      part[j] = p; rank[j] = r;
    }
    __syncthreads();
    // phase 2: reserve a contiguous run per destination in the global cursors
    if (threadIdx.x == 0) {
      uint32_t run = 0;
      for (uint32_t p = 0; p < P; p++) { s_off[p] = run; run += s_cnt[p]; }
      s_off[P] = run;
    }
    if (threadIdx.x < P) {
      uint32_t c = s_cnt[threadIdx.x];
      s_gbase[threadIdx.x] = (c ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)c) : 0ull) + (unsigned long long)d.dst_base[threadIdx.x];
    }
    __syncthreads();
    const uint32_t tile_rows = s_off[P];
    // phase 3: per column, regroup through shared memory and write coalesced runs
    for (int c = 0; c < d.ncols; c++) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(d.src[c]);
#pragma unroll
      for (int j = 0; j < PT_ITEMS; j++) {
        int64_t i = base + (int64_t)j * PT_BLOCK + threadIdx.x;
        if (i < n) s_val[s_off[part[j]] + rank[j]] = __ldcs(src + i);
      }
      __syncthreads();
      for (uint32_t sidx = threadIdx.x; sidx < tile_rows; sidx += PT_BLOCK) {
        uint32_t p = 0;
        while (sidx >= s_off[p + 1]) p++;   // ≤ nparts steps
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(d.dst[p][c]);
        dst[s_gbase[p] + (sidx - s_off[p])] = s_val[sidx];
      }
      __syncthreads();
    }
  }
}

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

int32_t tg_partition_of_key(int64_t key, int32_t nparts) { return (int32_t)part_of(mix64((uint64_t)key), (uint32_t)nparts); }

int tg_partition_count(int device, const int64_t* key_dev, int64_t rows, int32_t nparts, int64_t* part_counts_dev, void* stream) {
  TG_TRY(check_parts(nparts, 1));
  DeviceGuard g(device);
  if (!g.ok) return fail(TG_ERR_CUDA, "cudaSetDevice failed (no usable CUDA device)");
  cudaStream_t st = (cudaStream_t)stream;
  TG_CUDA(cudaMemsetAsync(part_counts_dev, 0, (size_t)nparts * 8, st));
  if (rows > 0)
    k_partition_count<<<pgrid(device, rows, 256, 8), 256, 0, st>>>(reinterpret_cast<const long long*>(key_dev), nullptr, rows, (uint32_t)nparts,
                                                                   reinterpret_cast<unsigned long long*>(part_counts_dev));
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
  if (rows > 0) k_partition_count<<<pgrid(device, rows, 256, 8), 256, 0, st>>>(key, key_nulls_dev, rows, (uint32_t)nparts, counts);
  k_partition_offsets<<<1, 32, 0, st>>>(counts, (uint32_t)nparts, reinterpret_cast<long long*>(part_offsets_dev), cursors);
  PartDst d{};
  d.nparts = nparts; d.ncols = ncols;
  for (int c = 0; c < ncols; c++) { d.src[c] = src_cols_dev[c]; for (int p = 0; p < nparts; p++) d.dst[p][c] = dst_cols_dev[c]; }
  d.dst_base = reinterpret_cast<const long long*>(part_offsets_dev);
  if (rows > 0) k_partition_scatter<<<pgrid(device, rows, PT_TILE, 4), PT_BLOCK, 0, st>>>(key, key_nulls_dev, rows, d, cursors);
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
  if (rows > 0)
    k_partition_scatter<<<pgrid(device, rows, PT_TILE, 4), PT_BLOCK, 0, st>>>(reinterpret_cast<const long long*>(key_dev), nullptr, rows, d, cursors);
  TG_CUDA(cudaGetLastError());
  TG_CUDA(cudaStreamSynchronize(st));
  return TG_OK;
}

}  // extern "C"
