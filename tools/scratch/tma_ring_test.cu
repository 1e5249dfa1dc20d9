// standalone check of the cp.async.bulk ring used by k_probe_inner_u1_tma: identity copy of two columns
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../tidb_b200/csrc/tma.cuh"
using namespace tg;
template <int STAGES, int MODE>
__global__ void __launch_bounds__(256) k_ring(const unsigned long long* a, const unsigned long long* b, long long ntiles,
                                               unsigned long long* oa, unsigned long long* ob) {
  constexpr int T = 1024, COLS = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* ring = reinterpret_cast<unsigned long long*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * COLS * T * 8);
  const int tid = threadIdx.x;
  const unsigned long long pol = l2_policy_evict_first();
  if (tid == 0) { for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1); mbar_fence_init(); }
  __syncthreads();
  auto issue = [&](long long it) {
    long long tile = (long long)blockIdx.x + it * gridDim.x;
    if (tile >= ntiles) return;
    int s = (int)(it % STAGES);
    unsigned long long* st = ring + (size_t)s * COLS * T;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(COLS * T * 8));
    bulk_g2s(st, a + tile * T, T * 8, &full[s], pol);
    bulk_g2s(st + T, b + tile * T, T * 8, &full[s], pol);
  };
  if (tid == 0) for (int it = 0; it < STAGES; it++) issue(it);
  for (long long it = 0;; it++) {
    const long long tile = (long long)blockIdx.x + it * gridDim.x;
    if (tile >= ntiles) break;
    const int s = (int)(it % STAGES);
    mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
    const unsigned long long* st = ring + (size_t)s * COLS * T;
    unsigned long long k[4], p[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { k[j] = st[j * 256 + tid]; p[j] = st[T + j * 256 + tid]; }
    if (MODE == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) issue(it + STAGES);
    // simulate work of varying length
    unsigned long long acc = 0;
    for (int r = 0; r < (int)((k[0] >> 60) & 15) * 8; r++) acc += __ldcg(a + ((k[1] + r * 977) % (ntiles * T)));
#pragma unroll
    for (int j = 0; j < 4; j++) { oa[tile * T + j * 256 + tid] = k[j] + (acc & 0); ob[tile * T + j * 256 + tid] = p[j]; }
  }
}
template <int STAGES, int MODE> int run(const char* name, long long ntiles, int grid) {
  size_t n = ntiles * 1024;
  std::vector<unsigned long long> ha(n), hb(n);
  for (size_t i = 0; i < n; i++) { ha[i] = i * 0x9E3779B97F4A7C15ull; hb[i] = i; }
  unsigned long long *a, *b, *oa, *ob;
  cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&oa, n * 8); cudaMalloc(&ob, n * 8);
  cudaMemcpy(a, ha.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(b, hb.data(), n * 8, cudaMemcpyHostToDevice);
  size_t smem = (size_t)STAGES * 2 * 1024 * 8 + STAGES * 8 + 16;
  cudaFuncSetAttribute(k_ring<STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long bad = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaMemset(oa, 0, n * 8); cudaMemset(ob, 0, n * 8);
    k_ring<STAGES, MODE><<<grid, 256, smem>>>(a, b, ntiles, oa, ob);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<unsigned long long> ra(n), rb(n);
    cudaMemcpy(ra.data(), oa, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(rb.data(), ob, n * 8, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n; i++) if (ra[i] != ha[i] || rb[i] != hb[i]) { if (bad < 5) printf("  %s mismatch at %zu (tile %zu, in-tile %zu)\n", name, i, i / 1024, i % 1024); bad++; }
  }
  printf("%s: stages=%d mode=%d ntiles=%lld grid=%d -> %lld mismatches\n", name, STAGES, MODE, ntiles, grid, bad);
  cudaFree(a); cudaFree(b); cudaFree(oa); cudaFree(ob);
  return bad != 0;
}
int main() {
  int rc = 0;
  rc |= run<4, 0>("plain", 2441, 444);
  rc |= run<4, 1>("proxyfence", 2441, 444);
  rc |= run<2, 0>("plain2", 2441, 444);
  rc |= run<4, 0>("plain_big", 20000, 444);
  return rc;
}
