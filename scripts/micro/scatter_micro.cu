// Microbenchmark: cost of the random gathers / reductions of splat_feat (K3) as a function of the
// address pattern and the lane <-> point mapping.  Not part of the product; results in profiles/.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_max_f16x4(uint2* addr, uint2 v) {
  asm volatile("red.global.max.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// MODE 0: gather only (sum into sink), 1: RED.MAX.F16x4 only, 2: gather then RED, 3: REDG.MIN.32 only
// MAP 0: thread owns 4 consecutive points, MAP 1: instruction k covers 32 consecutive points
template <int MODE, int MAP>
__global__ void k(const uint32_t* __restrict__ idx, const uint32_t* zb, uint2* fb, uint32_t* zb_w, int n, int active_pct, uint32_t* sink) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int warp_base = (t & ~31) * 4, lane = t & 31;
  uint32_t id[4];
  bool on[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = MAP == 0 ? t * 4 + j : warp_base + j * 32 + lane;
    on[j] = p < n && (int)((p * 2654435761u) >> 24) * 100 < active_pct * 256;
    id[j] = p < n ? idx[p] : 0;
    if (id[j] == 0xFFFFFFFFu) { on[j] = false; id[j] = 0; }
  }
  uint32_t acc = 0;
  uint32_t z[4];
  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = on[j] ? __ldcg(zb + id[j]) : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += z[j];
  }
  if (MODE == 1 || MODE == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (on[j] && (MODE == 1 || z[j] != 12345u)) red_max_f16x4(fb + id[j], make_uint2(id[j] & 0x3c003c00u, 0x3c00u));
  }
  if (MODE == 3) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (on[j]) atomicMin(zb_w + id[j], id[j] ^ 0x5555u);
  }
  if (acc == 0xdeadbeefu) *sink = acc;
}

template <int MODE, int MAP>
float run(const uint32_t* idx, const uint32_t* zb, uint2* fb, uint32_t* zbw, int n, int pct, uint32_t* sink, int reps = 20) {
  const int threads = 128, blocks = (n / 4 + threads - 1) / threads;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) k<MODE, MAP><<<blocks, threads>>>(idx, zb, fb, zbw, n, pct, sink);
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) k<MODE, MAP><<<blocks, threads>>>(idx, zb, fb, zbw, n, pct, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps * 1e3f;
}

int main() {
  const int H = 512, W = 1024, J = 8, HW = H * W, n = J * HW;
  std::vector<uint32_t> idx(n);
  std::mt19937 rng(1);
  uint32_t *d_idx, *d_zb, *d_zbw, *d_sink;
  uint2* d_fb;
  cudaMalloc(&d_idx, n * 4); cudaMalloc(&d_zb, n * 4); cudaMalloc(&d_zbw, n * 4); cudaMalloc(&d_fb, (size_t)n * 8); cudaMalloc(&d_sink, 4);
  cudaMemset(d_zb, 1, n * 4); cudaMemset(d_zbw, 0xFF, n * 4); cudaMemset(d_fb, 0, (size_t)n * 8);
  const char* names[] = {"identity", "jitter1 (+-1 px row/col)", "jitter3", "random within job", "real D_room (oracle indices)", "real D_rand (oracle indices)"};
  for (int pat = 0; pat < 6; ++pat) {
    if (pat >= 4) {
      // target pixel of every source point of one 512x1024 pano as the oracle computes it (-1 = rejected),
      // written by tests/tools/dump_scatter_indices.py; replicated over the 8 jobs
      FILE* f = fopen(pat == 4 ? "scripts/micro/bin/idx_room.i32" : "scripts/micro/bin/idx_rand.i32", "rb");
      if (!f) continue;
      std::vector<int32_t> one(HW);
      if (fread(one.data(), 4, HW, f) != (size_t)HW) { fclose(f); continue; }
      fclose(f);
      for (int j = 0; j < J; ++j)
        for (int i = 0; i < HW; ++i) idx[(size_t)j * HW + i] = one[i] < 0 ? 0xFFFFFFFFu : (uint32_t)(j * HW + one[i]);
    } else
    for (int j = 0; j < J; ++j)
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          int rr = r, cc = c;
          if (pat == 1 || pat == 2) {
            const int a = pat == 1 ? 1 : 3;
            rr = std::min(H - 1, std::max(0, r + (int)(rng() % (2 * a + 1)) - a));
            cc = (c + (int)(rng() % (2 * a + 1)) - a + W) % W;
          } else if (pat == 3) {
            rr = rng() % H; cc = rng() % W;
          }
          idx[(size_t)j * HW + r * W + c] = j * HW + rr * W + cc;
        }
    cudaMemcpy(d_idx, idx.data(), n * 4, cudaMemcpyHostToDevice);
    printf("pattern %s, %d points\n", names[pat], n);
    for (int pct : {100, 75, 50}) {
      printf("  active %3d%%: gather  map0 %6.1f us map1 %6.1f us | red.f16x4 map0 %6.1f map1 %6.1f | gather+red map0 %6.1f map1 %6.1f | redg.min32 map0 %6.1f map1 %6.1f\n", pct,
             run<0, 0>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink), run<0, 1>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink),
             run<1, 0>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink), run<1, 1>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink),
             run<2, 0>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink), run<2, 1>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink),
             run<3, 0>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink), run<3, 1>(d_idx, d_zb, d_fb, d_zbw, n, pct, d_sink));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
