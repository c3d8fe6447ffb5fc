// kernels.cuh -- sm_100a kernels of the geometric guidance path.
//
// Fused path (se3ds_reproject), three launches per job chunk, no point cloud in memory:
//   K2 splat_depth_kernel : RGB-D planes -> unproject -> +src -tgt (-> rotate) -> project ->
//                           REDG.MIN into the z-buffer: the 64-bit packed (depth bits | point
//                           index) key when winner indices are wanted, the 32-bit depth bits
//                           otherwise; per point (pixel, depth) goes to a scratch for K3.  The
//                           projection is a certified MUFU fast path with a canonical (IEEE)
//                           fallback for the few points it cannot certify.
//                           A thread owns 4 consecutive source pixels (128-bit loads / stores);
//                           the kernel is bound by instruction issue, its time follows its
//                           instruction count (rotation is a separate instantiation for that reason).
//   K3 splat_feat_kernel  : tolerance test d < dmin + 0.1 against the final z-buffer; every
//                           surviving point REDG.MAX.F16x4 its RGB into the feature buffer (one
//                           8-byte vector reduction = the reference's per-channel scatter-max);
//                           rejected points are block-reduced into the reject bin.  Instruction k
//                           of a warp covers 32 consecutive source pixels, so the lanes of a
//                           gather / reduction that share a cache line are merged.
//   K4 resolve_kernel     : per target pixel, streaming: z-buffer + feature buffer (+ bin on the
//                           owner pixel) -> proj_image / proj_depth / proj_mask / winner, written
//                           once (RGB staged through shared memory for 512-byte store
//                           instructions, 256-bit loads); re-arms the touched z-buffer / feature
//                           entries.
// The three kernels are chained with programmatic dependent launch (pdl_enter); the streams that are
// read once carry L2 evict_first hints.
// Compat path (se3ds_unproject_equirect / se3ds_project_cloud) works on materialised clouds with
// float32 features of any channel count; the resampling kernels (rotate_pano, perspective <->
// equirect, tf-style resize, tfa-style bilinear gather) are at the end of the file.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "canon_math.cuh"

namespace se3ds {

constexpr int kThreads = 128;

// Programmatic dependent launch: the three kernels of a chunk (and of consecutive calls) are
// launched with programmatic stream serialization, so the blocks of the next kernel are scheduled
// into the tail of the running one.  pdl_enter() first lets the *next* grid start launching, then
// waits until the *previous* grid has completed and its memory is visible -- every dependent
// access (z-buffer, feature buffer, scratch, bins) comes after it.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}
// Profiling (se3ds_ws_profile mode 2): every block stamps the end of its kernel; the latest stamp of a
// kernel minus the latest stamp of the kernel before it is that kernel's share of the pipelined step
// (programmatic dependent launch stays on, unlike with events between the launches).
constexpr int kStampSlots = 64;  // the blocks of a kernel spread their stamps over this many words (same-address atomics serialise)
__device__ __forceinline__ void stamp_end(const unsigned long long* stamps_const, int k) {
  unsigned long long* stamps = const_cast<unsigned long long*>(stamps_const);
  if (stamps != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned b = blockIdx.x + blockIdx.y * gridDim.x + blockIdx.z * gridDim.x * gridDim.y;
    atomicMax(stamps + k * kStampSlots + (b % kStampSlots), t);
  }
}
constexpr unsigned long long kZArmed = 0xFFFFFFFFFFFFFFFFull;

// scratch word: pixel index in bits 0..27, flags above
constexpr uint32_t kScPixMask = 0x0FFFFFFFu;
constexpr uint32_t kScInvalid = 1u << 28;   // rejected before the tolerance test (no pixel)
constexpr uint32_t kScDepthInv = 1u << 29;  // depth-invalid source pixel: feature = unproject_void
constexpr uint32_t kScDropped = 1u << 30;   // removed by the compaction (FILTER_VOID)

// Reject bin of one reference call (all-zero = armed): zneg = ~ordered(min depth), f = max feature.
struct Bin {
  uint32_t zneg;
  int f[3];
  uint32_t own1;  // depth bits + 1 of the owner pixel's own winner (0 = none), where the bin is applied later
  uint32_t pad[27];  // one replica per 128-byte line
};
// A bin is kept in kBinReplicas copies, 128 bytes apart; a block updates copy blockIdx % kBinReplicas and the
// consumer (the owner pixel's thread, the patch / export kernels) folds them.  Loads and atomics on ONE address
// from every block of a kernel serialise at one L2 slice (measured on the compat path: 70 % of a kernel's
// stalls); spread over 16 lines they do not.
constexpr int kBinReplicas = 16;
__device__ __forceinline__ Bin* bin_replica(Bin* bin) {
  return bin + ((blockIdx.x + blockIdx.y * 7u + blockIdx.z * 3u) % kBinReplicas);
}
// folds the replicas into `out` (zneg: max, f: max, own1: replica 0 only) and re-arms them
__device__ __forceinline__ Bin bin_fold_and_rearm(Bin* bin) {
  Bin r{0u, {0, 0, 0}, bin->own1, {}};
  for (int i = 0; i < kBinReplicas; ++i) {
    r.zneg = max(r.zneg, bin[i].zneg);
    r.f[0] = max(r.f[0], bin[i].f[0]); r.f[1] = max(r.f[1], bin[i].f[1]); r.f[2] = max(r.f[2], bin[i].f[2]);
    bin[i].zneg = 0u; bin[i].f[0] = 0; bin[i].f[1] = 0; bin[i].f[2] = 0; bin[i].own1 = 0u;
  }
  return r;
}

struct FusedParams {
  const void* rgb;
  const float* depth;
  const float* src_pos;
  const float* tgt_pos;
  const float* tgt_rot;  // (N,P,3,3) row-major or nullptr (translation only, as the reference)
  const float* tab;  // sin_e[H], cos_e[H], sin_h[W], cos_h[W]
  unsigned long long* zbuf;  // KEY64: depth_bits << 32 | point_index << 1 | depth_invalid
  uint32_t* zbuf32;          // !KEY64: depth bits only (no winner index requested)
  uint2* fbuf;
  uint32_t* sc_flat;
  float* sc_rad;
  Bin* bins;  // J bins of kBinReplicas copies each (per call: only the first bin is used)
  float* out_image;
  float* out_depth;
  float* out_mask;
  int* out_winner;
  int N, S, P, H, W, HW;
  int SC;          // frame slots per item in rgb / depth / src_pos (>= S: a frame ring that is partly filled)
  int n0, p0, PC;  // chunk: items n0.., poses p0..p0+PC-1; blockIdx.z = (n-n0)*PC + (p-p0)
  int mh;          // int(H * mask_proportion)
  int mask_frames;
  int uv, pv;
  unsigned flags;
  float depth_scale;
  int finalize_bins;  // 1: the bin of an owner pixel is complete when K4 runs
  float* bin_out;     // export mode: no owner pixel, the bin is handed to the caller
  float inv_depth_scale;  // RN(1 / depth_scale), computed on the host with an IEEE division
  FastProj fast;          // certified fast projection (canon_math.cuh)
  int prefilter_z, prefilter_f;  // read-before-reduce filters (several frames per job, L2-resident buffer)
  unsigned long long* stamps;  // profiling: end-of-kernel %globaltimer stamps {K2, K3, K4} of this chunk, or nullptr
  unsigned long long* dbg;  // verify mode: {points, certified, certified-but-wrong, max |dfx| bits<<32 | max |dfy| bits}
};

__device__ __forceinline__ bool row_masked(const FusedParams& q, int s, int row) {
  return s < q.mask_frames && (row < q.mh || row > q.H - q.mh);
}

__device__ __forceinline__ int3 load_rgb1(const uint8_t* rgb, size_t pix) {
  const uint8_t* p = rgb + pix * 3;
  return make_int3(p[0], p[1], p[2]);
}
__device__ __forceinline__ int3 load_rgb1(const int* rgb, size_t pix) {
  const int* p = rgb + pix * 3;
  return make_int3(p[0], p[1], p[2]);
}

// L2 eviction-priority hint (createpolicy + .L2::cache_hint): the streams that are read exactly once
// (input depth and colour, the per-point scratch) are marked evict_first so that they do not push the
// z-buffer / feature buffer out of L2 between the kernels (measured: K3 27.0 -> 24.8 us).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t ldcg_u32_stream(const void* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32_stream(const void* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ldg_u128_stream(const void* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u8_stream(const void* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// A pixel's raw colour held in registers: uint8 sources stay packed in one word until they are used.
template <typename RGB_T> struct RawRGB;
template <> struct RawRGB<uint8_t> {
  uint32_t v = 0;  // r | g << 8 | b << 16
  __device__ __forceinline__ void load(const uint8_t* rgb, size_t pix, uint64_t pol) {
    const uint8_t* p = rgb + pix * 3;
    v = ldg_u8_stream(p, pol) | (ldg_u8_stream(p + 1, pol) << 8) | (ldg_u8_stream(p + 2, pol) << 16);
  }
  __device__ __forceinline__ int3 get() const { return make_int3(v & 255u, (v >> 8) & 255u, v >> 16); }
  __device__ __forceinline__ uint32_t packed() const { return v; }
};
template <> struct RawRGB<int> {
  int3 v = {0, 0, 0};
  __device__ __forceinline__ void load(const int* rgb, size_t pix, uint64_t pol) {
    const int* p = rgb + pix * 3;
    v = make_int3((int)ldg_u32_stream(p, pol), (int)ldg_u32_stream(p + 1, pol), (int)ldg_u32_stream(p + 2, pol));
  }
  __device__ __forceinline__ int3 get() const { return v; }
  __device__ __forceinline__ uint32_t packed() const { return 0u; }  // (only uint8 colours have a packed form)
};

// Feature of a source pixel as the reference sees it after mask_pano + unprojection
// (pano_utils.py:225,262-265): depth-invalid -> unproject_void, masked row -> -1, else RGB.
__device__ __forceinline__ int3 point_feat(const FusedParams& q, bool dinv, bool masked, int3 raw) {
  if (dinv) return make_int3(q.uv, q.uv, q.uv);
  if (masked) return make_int3(-1, -1, -1);
  return raw;
}

__device__ __forceinline__ uint2 pack_f16x4(int3 f) {
  const uint32_t r = __half_as_ushort(__int2half_rn(f.x));
  const uint32_t g = __half_as_ushort(__int2half_rn(f.y));
  const uint32_t b = __half_as_ushort(__int2half_rn(f.z));
  return make_uint2(r | (g << 16), b);
}
// the same for three bytes r | g << 8 | b << 16 without conversions: 0x6400 | x is 1024 + x in float16 (ulp 1), and
// subtracting 1024 is exact -- two byte permutes and two packed subtractions instead of three I2F and two permutes
__device__ __forceinline__ uint2 pack_f16x4_u8(uint32_t c) {
  const __half2 k = __halves2half2(__ushort_as_half((unsigned short)0x6400), __ushort_as_half((unsigned short)0x6400));
  const uint32_t rg = __byte_perm(c, 0x64u, 0x4140), b0 = __byte_perm(c, 0x64u, 0x4542);
  const __half2 hrg = __hsub2(*reinterpret_cast<const __half2*>(&rg), k), hb0 = __hsub2(*reinterpret_cast<const __half2*>(&b0), k);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&hrg), *reinterpret_cast<const uint32_t*>(&hb0));
}
__device__ __forceinline__ float3 unpack_f16x4(uint2 v) {
  return make_float3(__half2float(__ushort_as_half((unsigned short)(v.x & 0xffff))),
                     __half2float(__ushort_as_half((unsigned short)(v.x >> 16))),
                     __half2float(__ushort_as_half((unsigned short)(v.y & 0xffff))));
}
__device__ __forceinline__ void red_max_f16x4(uint2* addr, uint2 v) {
  asm volatile("red.global.max.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// Reject bin.  Every rejected point of a call lands on flat index 0 in the reference.  Here every
// block reduces its rejected points (REDUX per warp, shared memory across warps) and one thread
// issues the REDG only if it can change the bin: same-address atomics serialise at one L2 slice
// (~50-100 ns each), a plain read does not, and bins only grow, so a stale read is a safe filter.
__device__ __forceinline__ void bin_update_z(Bin* bin, uint32_t zneg) {
  if (zneg > __ldcg(&bin->zneg)) atomicMax(&bin->zneg, zneg);
}
__device__ __forceinline__ void bin_max_feat_block(Bin* bin, bool has, int3 f) {
  __shared__ int sm_f[kThreads / 32][3];
  const int r = __reduce_max_sync(0xffffffffu, has ? f.x : 0);
  const int g = __reduce_max_sync(0xffffffffu, has ? f.y : 0);
  const int b = __reduce_max_sync(0xffffffffu, has ? f.z : 0);
  if ((threadIdx.x & 31) == 0) { sm_f[threadIdx.x >> 5][0] = r; sm_f[threadIdx.x >> 5][1] = g; sm_f[threadIdx.x >> 5][2] = b; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int m = sm_f[0][threadIdx.x];
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) m = max(m, sm_f[i][threadIdx.x]);
    if (m > 0 && m > __ldcg(&bin->f[threadIdx.x])) atomicMax(&bin->f[threadIdx.x], m);
  }
}

// Launch geometry shared by K2 / K3: blockIdx.y = source row, blockIdx.z = local job * S + frame, a
// block covers kThreads * PPT consecutive columns and a warp 32 * PPT of them.  Point k of a lane
// sits at column col0 + 32 * k, i.e. every load, store and -- what matters -- every reduction
// instruction of a warp covers 32 *consecutive* source pixels: neighbouring source pixels mostly
// land on neighbouring target pixels, and the memory system merges the lanes of one instruction
// that fall into the same line (measured, scripts/micro/scatter_micro.cu: a thread owning 4
// consecutive points instead costs 1.2-1.5x more for the same REDG / gather traffic on the room
// workload, 2.5x on an identity map).
struct SrcIdx {
  int lj, s, n, p, job, row, col0;
};
// K3 uses that mapping (STRIDE = 32).  K2 is bound by instruction issue, its reductions are hidden
// behind the projection math: it keeps 4 consecutive pixels per thread (STRIDE = 1) for the 128-bit
// depth / table loads and scratch stores (c3: K2 699 us against 735 us with the strided mapping).
// The scratch is indexed by source pixel, so the two kernels need not agree.
template <int PPT, int STRIDE>
__device__ __forceinline__ SrcIdx src_index(const FusedParams& q) {
  SrcIdx i;
  const int z = blockIdx.z;
  if (q.S == 1) { i.lj = z; i.s = 0; } else { i.lj = z / q.S; i.s = z - i.lj * q.S; }
  if (q.PC == 1) { i.n = q.n0 + i.lj; i.p = q.p0; } else { const int a = i.lj / q.PC; i.n = q.n0 + a; i.p = q.p0 + (i.lj - a * q.PC); }
  i.job = i.n * q.P + i.p;
  i.row = blockIdx.y;
  if constexpr (STRIDE == 1) i.col0 = (blockIdx.x * kThreads + threadIdx.x) * PPT;
  else i.col0 = blockIdx.x * (kThreads * PPT) + (threadIdx.x >> 5) * (32 * PPT) + (threadIdx.x & 31);
  return i;
}


// ------------------------------------------------------------------------------------------
// K2: fused unproject + translate + project + depth splat
// ------------------------------------------------------------------------------------------
// Launch geometry: blockIdx.z = local job * S + frame, blockIdx.x = block of kThreads * 4 columns,
// blockIdx.y = row group: a block walks the rows blockIdx.y + k * gridDim.y (k = 0 .. rows_per_block),
// so the per-thread set-up (index math, poses, the column sines / cosines) is paid once per ~28 points
// and the masked and unmasked rows of a pano are spread evenly over the blocks.  A thread owns 4
// consecutive columns (VEC: one 128-bit depth load and two 128-bit scratch stores per row); the depth
// of the next row is loaded while the current one is processed.  The host sizes rows_per_block so
// that the grid is about one resident wave (8 blocks per SM).
//
// The kernel is bound by instruction issue (ncu: 70 % issue-active, no pipe above 50 %), so the per-point path is
// kept short (~120 SASS instructions on an unmasked row, 127 per point over the whole launch):
//  * rad: the two-fma refinement of MUFU.RSQ (fast_rad, = __fsqrt_rn for normal operands); its
//    reciprocal square root doubles as 1 / rad for the elevation.
//  * pixel: certified fast projection in pixel units (project_pixel_fast).
//  * Points the fast path cannot certify (a few in a thousand), and points whose squared radius is
//    zero / denormal / huge / not finite, are pushed on a warp-private stack in shared memory (one
//    vote per point on the common path) and projected canonically 32 at a time by the whole warp
//    (drain): no divergence in the slow path, no block barrier anywhere in the kernel.
// FAST: uint8 RGB with project_void == -1 (every reference caller) and, if the compaction is on,
// unproject_void == -1: a raw colour can then never equal a void class, so the fate of a point
// (0 dropped / 1 rejected: only its depth feeds the reject bin / 2 projected) follows from its depth
// validity and the row mask alone; otherwise the colours are read and compared.
// PROJ: 0 = every point takes the canonical path (through the stack), 1 = certified fast path,
// 2 = verify (both projections for every point, disagreements counted into q.dbg).
// KEY64: the z-buffer holds the 64-bit packed (depth | point index) key -- the deterministic winner
// (nearest depth, lowest index).  When the caller does not ask for winner indices the same minimum
// depth comes from a 32-bit key (depth bits only): half the z-buffer traffic, identical guidance.
constexpr int kStackCap = 160;  // < 32 entries survive a row, a row pushes at most 128 per warp

// PLAIN (host-checked): FAST, VEC, every thread of the grid inside the row (W % 512 == 0), no compaction and
// unproject_void == -1 -- what SE3DSModel / eval_metric run: a point is projected iff its depth is valid and its row
// is not masked, and rejected otherwise; nothing is ever dropped.  Same results, ~10 % fewer instructions per point.
template <typename RGB_T, bool VEC, bool FAST, int PROJ, bool KEY64, bool ROT, bool PLAIN>
__global__ void __launch_bounds__(kThreads, 8) splat_depth_kernel(const FusedParams q) {
  static_assert(!PLAIN || (FAST && VEC), "PLAIN is a specialisation of the FAST, vectorised kernel");
  // K2 only reads caller inputs until it touches the z-buffer / scratch / bins.  If the caller
  // guarantees that those inputs were not produced by the kernel launched just before this call
  // (SE3DS_FLAG_INPUTS_READY), the wait for the previous grid -- normally the resolve that re-arms
  // the z-buffer -- is postponed until after the first row's projection math; otherwise it comes first.
  pdl_launch_dependents();
  bool waited = !(q.flags & SE3DS_FLAG_INPUTS_READY);
  if (waited) pdl_wait();
  constexpr int kWarps = kThreads / 32;
  __shared__ float4 sq[kWarps][kStackCap];  // X, Y, Z, bits: source pixel | projected << 30 | depth-valid << 31
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int zz = blockIdx.z;
  int lj, s;
  if (q.S == 1) { lj = zz; s = 0; } else { lj = zz / q.S; s = zz - lj * q.S; }
  int n, p;
  if (q.PC == 1) { n = q.n0 + lj; p = q.p0; } else { const int a = lj / q.PC; n = q.n0 + a; p = q.p0 + (lj - a * q.PC); }
  const int job = n * q.P + p;
  const int col0 = (blockIdx.x * kThreads + threadIdx.x) * 4;
  const int W = q.W, H = q.H;
  bool act[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) act[k] = PLAIN || col0 + (VEC ? 0 : k) < W;  // VEC: W % 4 == 0, the four points are active together

  Bin* bin = bin_replica(q.bins + (size_t)((q.flags & SE3DS_FLAG_BIN_PER_JOB) ? job : 0) * kBinReplicas);
  // 32-bit element offsets into the chunk's z-buffer and scratch (the host keeps a chunk below 2^31 elements): the
  // address of a reduction is one add and one widening multiply-add on a base taken from the constant bank
  const uint32_t zoff = (uint32_t)lj * (uint32_t)q.HW;
  unsigned long long* const zb = q.zbuf;
  uint32_t* const zb32 = q.zbuf32;
  const uint32_t sc_frame = ((uint32_t)lj * (uint32_t)q.S + (uint32_t)s) * (uint32_t)q.HW;
  const uint32_t idx_frame = (uint32_t)(s * q.HW);
  const size_t frame = (size_t)(n * q.SC + s) * q.HW;
  const float* dframe = q.depth + frame;
  const float *sin_e = q.tab, *cos_e = q.tab + H, *sin_h = q.tab + 2 * H, *cos_h = sin_h + W;
  const uint64_t stream_pol = l2_policy_evict_first();

  float sh[4] = {}, ch[4] = {};
  if constexpr (VEC) {
    if (act[0]) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sin_h + col0));
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(cos_h + col0));
      sh[0] = s4.x; sh[1] = s4.y; sh[2] = s4.z; sh[3] = s4.w;
      ch[0] = c4.x; ch[1] = c4.y; ch[2] = c4.z; ch[3] = c4.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (act[k]) { sh[k] = __ldg(sin_h + col0 + k); ch[k] = __ldg(cos_h + col0 + k); }
  }
  const float* sp = q.src_pos + (size_t)(n * q.SC + s) * 3;
  const float* tp = q.tgt_pos + (size_t)job * 3;
  const float sx = __ldg(sp), sy = __ldg(sp + 1), sz = __ldg(sp + 2);
  const float tx = __ldg(tp), ty = __ldg(tp + 1), tz = __ldg(tp + 2);
  float rot[ROT ? 9 : 1];  // ROT: full SE(3) target pose (q.tgt_rot != nullptr), a separate instantiation
  if constexpr (ROT) {
#pragma unroll
    for (int i = 0; i < 9; ++i) rot[i] = __ldg(q.tgt_rot + (size_t)job * 9 + i);
  }
  const bool filt = q.flags & SE3DS_FLAG_FILTER_VOID;
  const bool prefilter = q.prefilter_z != 0;
  // FAST: fate of a depth-invalid point (feature = unproject_void everywhere, pano_utils.py:225)
  const int a_void = filt ? 0 : (q.uv != -1 ? 2 : 1);

  uint32_t minb = 0x7fffffffu;  // smallest depth bits among this thread's rejected points (depths are > 0 here)
  int wq = 0;                   // entries on the warp's stack (warp-uniform)

  // canonical projection of up to 32 listed points by the whole warp (all entries of a block belong to its
  // job-frame)
  auto drain_list = [&](const float4* list, int m) {
    if (lane < m) {
      const float4 e = list[lane];
      const uint32_t meta = __float_as_uint(e.w);
      const int pix = (int)(meta & kScPixMask);
      const bool dvalid = meta >> 31, isproj = (meta >> 30) & 1u;
      const float rad = canon_rad(e.x, e.y, e.z);
      const int tpix = isproj ? project_pixel_rad(e.x, e.y, e.z, H, W, rad) : -1;
      const uint32_t dflag = dvalid ? 0u : kScDepthInv;
      uint32_t word;
      if (tpix >= 0) {
        word = (uint32_t)tpix | dflag;
        if constexpr (KEY64) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(rad) << 32) | ((idx_frame + (uint32_t)pix) << 1) | (dvalid ? 0u : 1u);
          if (!prefilter || key < __ldcg(zb + (zoff + (uint32_t)tpix))) atomicMin(zb + (zoff + (uint32_t)tpix), key);
        } else {
          const uint32_t key = __float_as_uint(rad);
          if (!prefilter || key < __ldcg(zb32 + (zoff + (uint32_t)tpix))) atomicMin(zb32 + (zoff + (uint32_t)tpix), key);
        }
      } else {  // rejected: only its depth feeds the reject bin (point_cloud_utils.py:146-159)
        word = kScInvalid | dflag;
        // a NaN depth never lowers a minimum (scatter-min ignores it); said explicitly -- the bit pattern of a NaN
        // that the compiler re-materialises is not something to build an ordering on
        const uint32_t zneg = (rad != rad) ? 0u : ~f32_ordered(rad);
        if (zneg) bin_update_z(bin, zneg);
      }
      __stcg(q.sc_flat + (sc_frame + (uint32_t)pix), word);
      __stcg(q.sc_rad + (sc_frame + (uint32_t)pix), rad);
    }
  };
  auto drain = [&]() {  // the top min(wq, 32) entries of the warp's stack
    __syncwarp();
    const int m = min(wq, 32);
    drain_list(&sq[wid][wq - m], m);
    wq -= m;
    __syncwarp();
  };

  const int row_step = gridDim.y;
  int row = blockIdx.y;
  // depth of the first row
  float dn[4] = {};
  auto load_depth = [&](int r) {
    if constexpr (VEC) {
      if (act[0]) {
        const uint4 dv = ldg_u128_stream(dframe + (size_t)r * W + col0, stream_pol);
        dn[0] = __uint_as_float(dv.x); dn[1] = __uint_as_float(dv.y); dn[2] = __uint_as_float(dv.z); dn[3] = __uint_as_float(dv.w);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (act[k]) dn[k] = __uint_as_float(ldg_u32_stream(dframe + (size_t)r * W + col0 + k, stream_pol));
    }
  };
  if (row < H) load_depth(row);
  for (; row < H; row += row_step) {
    float d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = dn[k];
    if (row + row_step < H) load_depth(row + row_step);
    const float se = __ldg(sin_e + row), ce = __ldg(cos_e + row);
    const bool masked = row_masked(q, s, row);
    const int pix0 = row * W + col0;
    // FAST: fate of a depth-valid point of this row (masked rows hold -1 features, pano_utils.py:262-265)
    const int a_row = masked ? (filt ? 0 : 1) : 2;
    RawRGB<RGB_T> raw[4];
    if constexpr (!FAST) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (act[k]) raw[k].load(static_cast<const RGB_T*>(q.rgb), frame + pix0 + k, stream_pol);
    }
    uint32_t scf[4];  // scratch word; a point is splatted below iff it carries a pixel (neither kScInvalid nor kScDropped)
    float scr[4];
    // LEAN rows (FAST only): no point of the row is projected (a masked row whose depth-invalid points
    // are not projected either) -- only the depths of the rejected points are needed, for the reject bin.
    auto points = [&](auto lean_tag) {
      constexpr bool LEAN = decltype(lean_tag)::value;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // pano_utils.py:220-236
      const bool dvalid = d[k] > 0.0f && d[k] < 1.0f;
      const float rad0 = __fmul_rn(__fmul_rn(d[k], q.depth_scale), dvalid ? 1.0f : 0.0f);
      const float t = __fmul_rn(rad0, se);
      // models.py:225-226 then :273-275 -- two roundings
      float X = __fsub_rn(__fadd_rn(__fmul_rn(t, ch[k]), sx), tx);
      float Y = __fsub_rn(__fadd_rn(__fmul_rn(t, sh[k]), sy), ty);
      float Z = __fsub_rn(__fadd_rn(__fmul_rn(rad0, ce), sz), tz);
      if constexpr (ROT) {  // into the target camera frame: row-wise fma(r2, z, fma(r1, y, r0 * x))
        const float a = X, b = Y, c = Z;
        X = __fmaf_rn(rot[2], c, __fmaf_rn(rot[1], b, __fmul_rn(rot[0], a)));
        Y = __fmaf_rn(rot[5], c, __fmaf_rn(rot[4], b, __fmul_rn(rot[3], a)));
        Z = __fmaf_rn(rot[8], c, __fmaf_rn(rot[7], b, __fmul_rn(rot[6], a)));
      }
      int action = 0;  // 0 dropped (compaction, or past the end of the row), 1 rejected (void feature), 2 projected
      if constexpr (PLAIN) {
        // nothing is dropped; a masked row (LEAN) rejects all its points, an unmasked one its depth-invalid points
      } else if constexpr (FAST) {
        action = dvalid ? a_row : a_void;
        if (!act[k]) action = 0;
      } else {
        const int3 f = point_feat(q, !dvalid, masked, raw[k].get());
        const bool dropped = filt && f.x == q.uv && f.y == q.uv && f.z == q.uv;
        const bool fv = f.x != q.pv && f.y != q.pv && f.z != q.pv;
        action = (dropped || !act[k]) ? 0 : (fv ? 2 : 1);
      }
      const bool proj = PLAIN ? (!LEAN && dvalid) : (!LEAN && action == 2);
      const bool rej = PLAIN ? (LEAN || !dvalid) : (action == 1);
      const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(X, X), __fmul_rn(Y, Y)), __fmul_rn(Z, Z));
      float rs;
      bool sq_ok;
      const float rad = fast_rad(r2, rs, sq_ok);
      int tpix = 0, frow = 0;
      float fx = 0.f;
      bool certain = false;
      if constexpr (!LEAN && PROJ != 0) certain = project_pixel_fast(X, Y, Z, rs, H, W, q.fast, tpix, fx, frow) && sq_ok;
      if constexpr (PROJ == 2 && !LEAN) {
        if (proj) {
          const float radx = canon_rad(X, Y, Z);
          const int exact = project_pixel_rad(X, Y, Z, H, W, radx);
          atomicAdd(q.dbg, 1ull);
          if (certain) {
            atomicAdd(q.dbg + 1, 1ull);
            if (exact != tpix || __float_as_uint(radx) != __float_as_uint(rad)) atomicAdd(q.dbg + 2, 1ull);
          }
          if (exact >= 0 && fabsf(Z) < 0.984375f * radx) {  // how far the fast column falls outside the canonical pixel
            // (the fast row is a certified candidate, not a coordinate: it only counts when it was certified)
            const float cx = (float)(exact % W);
            const float ex = fmaxf(fmaxf(cx - fx, fx - (cx + 1.0f)), 0.0f);
            const float ey = (certain && frow != exact / W) ? 1.0f : 0.0f;
            atomicMax(q.dbg + 3, ((unsigned long long)__float_as_uint(ex) << 32) | __float_as_uint(ey));
          }
          certain = false;  // the results are the canonical ones
        }
      }
      const uint32_t dflag = dvalid ? 0u : kScDepthInv;
      scr[k] = rad;
      // a projected point the fast path could not certify carries no pixel yet: the drain writes its final word
      scf[k] = ((proj && certain) ? (uint32_t)tpix : (rej ? kScInvalid : kScDropped)) | dflag;
      minb = min(minb, (rej && sq_ok) ? __float_as_uint(rad) : 0x7fffffffu);
      // the rest goes to the warp's stack: projected but uncertified, or a radius the fast square root
      // does not cover; drain() finishes those points (pixel, splat, scratch, reject bin)
      const bool defer = (proj && !certain) || (rej && !sq_ok);
      if (__any_sync(0xffffffffu, defer)) {
        const unsigned dmask = __ballot_sync(0xffffffffu, defer);
        if (defer)
          sq[wid][wq + __popc(dmask & ((1u << lane) - 1u))] =
              make_float4(X, Y, Z, __uint_as_float((uint32_t)(pix0 + k) | (proj ? 1u << 30 : 0u) | (dvalid ? 1u << 31 : 0u)));
        wq += __popc(dmask);
      }
    }
    };
    if (PLAIN ? masked : (FAST && a_row != 2 && a_void != 2)) points(std::true_type{});
    else points(std::false_type{});
    if (!waited) { pdl_wait(); waited = true; }  // from here on: z-buffer, scratch and bins of this workspace
    if (prefilter) {
      // multi-frame jobs: most points arrive at a pixel that already holds something nearer; a plain read
      // (the entry only decreases, so a stale value is a safe filter) saves the reduction.  The four
      // filter reads are issued together.
      unsigned long long cur[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool hit = !(scf[k] & (kScInvalid | kScDropped));
        if constexpr (KEY64) cur[k] = hit ? __ldcg(zb + (zoff + (scf[k] & kScPixMask))) : 0ull;
        else cur[k] = hit ? (unsigned long long)__ldcg(zb32 + (zoff + (scf[k] & kScPixMask))) : 0ull;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (KEY64) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(scr[k]) << 32) | ((idx_frame + (uint32_t)(pix0 + k)) << 1) | ((scf[k] & kScDepthInv) ? 1u : 0u);
          if (!(scf[k] & (kScInvalid | kScDropped)) && key < cur[k]) atomicMin(zb + (zoff + (scf[k] & kScPixMask)), key);
        } else {
          const uint32_t key = __float_as_uint(scr[k]);
          if (!(scf[k] & (kScInvalid | kScDropped)) && key < (uint32_t)cur[k]) atomicMin(zb32 + (zoff + (scf[k] & kScPixMask)), key);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (scf[k] & (kScInvalid | kScDropped)) continue;
        if constexpr (KEY64) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(scr[k]) << 32) | ((idx_frame + (uint32_t)(pix0 + k)) << 1) | ((scf[k] & kScDepthInv) ? 1u : 0u);
          atomicMin(zb + (zoff + (scf[k] & kScPixMask)), key);
        } else {
          atomicMin(zb32 + (zoff + (scf[k] & kScPixMask)), __float_as_uint(scr[k]));
        }
      }
    }
    // K3 skips the masked rows altogether when their features (-1 / unproject_void) cannot raise a maximum
    if (!(masked && q.uv <= 0)) {
      if constexpr (VEC) {
        if (act[0]) {
          __stcg(reinterpret_cast<uint4*>(q.sc_flat + (sc_frame + (uint32_t)pix0)), make_uint4(scf[0], scf[1], scf[2], scf[3]));
          __stcg(reinterpret_cast<float4*>(q.sc_rad + (sc_frame + (uint32_t)pix0)), make_float4(scr[0], scr[1], scr[2], scr[3]));
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (act[k]) {
            __stcg(q.sc_flat + (sc_frame + (uint32_t)(pix0 + k)), scf[k]);
            __stcg(q.sc_rad + (sc_frame + (uint32_t)(pix0 + k)), scr[k]);
          }
      }
    }
    while (wq >= 32) drain();
  }
  if (!waited) pdl_wait();
  // (a block-wide drain by the last warp was tried: fewer instructions, but the canonical projection is a long
  // dependent chain and serialising it at the end of every block made the kernel 6 us slower)
  while (wq > 0) drain();
  // reject bin: smallest depth of this warp's rejected points (the drains add their own)
  const uint32_t wmin = __reduce_min_sync(0xffffffffu, minb);
  if (lane == 0 && wmin != 0x7fffffffu) bin_update_z(bin, 0x7fffffffu - wmin);
  stamp_end(q.stamps, 0);
}

// ------------------------------------------------------------------------------------------
// K3: tolerance test + per-channel max of the surviving features
// ------------------------------------------------------------------------------------------
template <typename RGB_T, int PPT, bool KEY64>
__global__ void __launch_bounds__(kThreads, 16) splat_feat_kernel(const FusedParams q) {
  pdl_enter();
  constexpr int kLaneStride = PPT == 1 ? 1 : 32;
  const SrcIdx ix = src_index<PPT, kLaneStride>(q);
  // The feature buffer and the reject bin start at 0 (output_void_class) and only take maxima, so a
  // point whose channels are all <= 0 changes nothing.  A masked row holds only -1 / unproject_void
  // features: the whole block (= one row segment) has nothing to do.
  if (row_masked(q, ix.s, ix.row) && q.uv <= 0) return;
  Bin* bin = bin_replica(q.bins + (size_t)((q.flags & SE3DS_FLAG_BIN_PER_JOB) ? ix.job : 0) * kBinReplicas);
  bool bin_has = false;
  int3 bin_f = make_int3(0, 0, 0);
  if (ix.col0 < q.W) {
    const int pix0 = ix.row * q.W + ix.col0;  // point k: pixel pix0 + kLaneStride * k
    const size_t frame = (size_t)(ix.n * q.SC + ix.s) * q.HW;
    const uint32_t sc0 = ((uint32_t)ix.lj * (uint32_t)q.S + (uint32_t)ix.s) * (uint32_t)q.HW + (uint32_t)pix0;
    uint32_t scf[PPT];
    float scr[PPT];
    RawRGB<RGB_T> raw[PPT];
    const uint64_t stream_pol = l2_policy_evict_first();  // last use of the scratch, only use of the colours
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (ix.col0 + kLaneStride * k < q.W) {
        scf[k] = ldcg_u32_stream(q.sc_flat + (sc0 + (uint32_t)(kLaneStride * k)), stream_pol);
        scr[k] = __uint_as_float(ldcg_u32_stream(q.sc_rad + (sc0 + (uint32_t)(kLaneStride * k)), stream_pol));
        raw[k].load(static_cast<const RGB_T*>(q.rgb), frame + pix0 + kLaneStride * k, stream_pol);
      } else {  // past the end of the row
        scf[k] = kScDropped; scr[k] = 0.0f;
      }
    }
    // 32-bit element offsets (a chunk stays below 2^32 elements, see plan_chunks): base pointers from the constant bank
    const uint32_t zoff = (uint32_t)ix.lj * (uint32_t)q.HW;
    const unsigned long long* const zb = q.zbuf;
    const uint32_t* const zb32 = q.zbuf32;
    uint2* const fb = q.fbuf;
    // issue the z-buffer gathers first, then consume (only the depth half of the key is needed)
    uint32_t zbits[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const bool haspix = !(scf[k] & (kScDropped | kScInvalid));
      if constexpr (KEY64) zbits[k] = haspix ? (uint32_t)(__ldcg(zb + (zoff + (scf[k] & kScPixMask))) >> 32) : 0xFFFFFFFFu;
      else zbits[k] = haspix ? __ldcg(zb32 + (zoff + (scf[k] & kScPixMask))) : 0xFFFFFFFFu;
    }
    const bool masked = row_masked(q, ix.s, ix.row);
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const bool live = !(scf[k] & kScDropped);
      const bool haspix = !(scf[k] & (kScDropped | kScInvalid));
      // point_cloud_utils.py:168-169: depth < min_depth + 0.1 (min_depth includes the init fill)
      const float zmin = fminf(__uint_as_float(zbits[k]), q.depth_scale);  // armed bits are a NaN: fminf -> depth_scale
      const bool keep = haspix && scr[k] < __fadd_rn(zmin, 0.1f);
      // the feature is the raw colour unless the depth was invalid or the row is masked (pano_utils.py:225,262-265)
      const bool plain = std::is_same<RGB_T, uint8_t>::value && !masked && !(scf[k] & kScDepthInv);
      if (keep) {
        uint2 packed;
        int3 f = make_int3(0, 0, 0);
        if (plain) {
          packed = pack_f16x4_u8(raw[k].packed());
        } else {
          f = point_feat(q, scf[k] & kScDepthInv, masked, raw[k].get());
          packed = pack_f16x4(f);
        }
        // With several source frames many points share a pixel and most of them cannot raise the
        // maximum any more: a plain read (the buffer only grows, a stale value is a safe filter)
        // saves the reduction.  With one frame, or a workspace that does not fit in L2, the read
        // costs more than it saves (measured: c3 -25 %, c5 +19 % for K3), so the host decides.
        bool need = true;
        if (q.prefilter_f) {
          if (plain) f = raw[k].get();
          const float3 cur = unpack_f16x4(__ldcg(fb + (zoff + (scf[k] & kScPixMask))));
          need = (float)f.x > cur.x || (float)f.y > cur.y || (float)f.z > cur.z;
        }
        if (need) red_max_f16x4(fb + (zoff + (scf[k] & kScPixMask)), packed);
      } else if (live) {  // rejected: its feature goes to the reject bin
        const int3 f = point_feat(q, scf[k] & kScDepthInv, masked, raw[k].get());
        bin_has = true;
        bin_f.x = max(bin_f.x, f.x); bin_f.y = max(bin_f.y, f.y); bin_f.z = max(bin_f.z, f.z);
      }
    }
  }
  bin_max_feat_block(bin, bin_has, bin_f);
  stamp_end(q.stamps, 1);
}

// ------------------------------------------------------------------------------------------
// K4: gather-resolve -> guidance tensors.  blockIdx.y = target row, blockIdx.z = local job.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float clip01_div255(float v) {
  // models.py:290-291: clip_by_value(rgb / 255, 0, 1).  v / 255 as q0 = v*y, r = v - 255*q0 (exact,
  // fma), q = q0 + r*y with y = RN(1/255): bit-identical to IEEE division for every float32
  // (exhaustively verified over all mantissas, tests/test_canon_math.py).
  const float y = 0x1.010102p-8f;
  const float q0 = __fmul_rn(v, y);
  const float q1 = __fmaf_rn(__fmaf_rn(-255.0f, q0, v), y, q0);
  return fminf(fmaxf(q1, 0.0f), 1.0f);
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): one full 32 B sector per lane.
__device__ __forceinline__ void ldcg_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void st_256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// Loads the z-buffer keys and feature maxima of four consecutive pixels (flat entry e, multiple of 4).
// !KEY64: depth bits go to the upper half, the index bits are all ones.
template <bool KEY64>
__device__ __forceinline__ void resolve_load4(const FusedParams& q, size_t e, unsigned long long (&key)[4], uint2 (&fv)[4]) {
  if constexpr (KEY64) {
    uint4 a, b;
    ldcg_256(q.zbuf + e, a, b);
    key[0] = ((unsigned long long)a.y << 32) | a.x; key[1] = ((unsigned long long)a.w << 32) | a.z;
    key[2] = ((unsigned long long)b.y << 32) | b.x; key[3] = ((unsigned long long)b.w << 32) | b.z;
  } else {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(q.zbuf32 + e));
    key[0] = ((unsigned long long)a.x << 32) | 0xFFFFFFFFu; key[1] = ((unsigned long long)a.y << 32) | 0xFFFFFFFFu;
    key[2] = ((unsigned long long)a.z << 32) | 0xFFFFFFFFu; key[3] = ((unsigned long long)a.w << 32) | 0xFFFFFFFFu;
  }
  uint4 c, d;
  ldcg_256(q.fbuf + e, c, d);
  fv[0] = make_uint2(c.x, c.y); fv[1] = make_uint2(c.z, c.w);
  fv[2] = make_uint2(d.x, d.y); fv[3] = make_uint2(d.z, d.w);
}

// One target pixel: z-buffer key + feature maxima -> (depth, mask, rgb, winner), including the owner
// pixel of a reject bin (point_cloud_utils.py:160-162, models.py:282-293).
// Compact colour of a pixel: the raw per-channel maxima clamped to [0, 255] as bytes (r | g << 8 | b << 16).
// clip(x / 255, 0, 1) of the float32 contract is a function of exactly that byte (the maxima are
// integers >= 0), so expand_guidance_kernel reproduces proj_image bit for bit.
__device__ __forceinline__ uint32_t pack_rgb_u8(float3 f) {
  const uint32_t r = (uint32_t)min(max(__float2int_rz(f.x), 0), 255), g = (uint32_t)min(max(__float2int_rz(f.y), 0), 255),
                 b = (uint32_t)min(max(__float2int_rz(f.z), 0), 255);
  return r | (g << 8) | (b << 16);
}

template <bool KEY64, bool COMPACT = false>
__device__ __forceinline__ void resolve_pixel(const FusedParams& q, int job, int pix, unsigned long long key, uint2 fvk,
                                              float& od, float& om, float* oi, int& ow, uint32_t* packed = nullptr) {
  const bool per_job = q.flags & SE3DS_FLAG_BIN_PER_JOB;
  const bool has = key != kZArmed;
  const float radw = __uint_as_float((uint32_t)(key >> 32));
  float zmin = has ? fminf(radw, q.depth_scale) : q.depth_scale;
  float3 f = unpack_f16x4(fvk);  // per-channel max of every point that passed the tolerance test
  ow = (KEY64 && has && radw <= q.depth_scale) ? (int)((uint32_t)key >> 1) : -1;  // winner indices exist with 64-bit keys only
  if ((pix == 0) && (per_job || job == 0)) {  // owner pixel of a reject bin
    // A rejected point that is nearer than every valid point of this pixel takes the pixel's depth
    // (scatter-min over all points, point_cloud_utils.py:157-159): then no valid point is the winner.
    Bin* bin = q.bins + (size_t)(per_job ? job : 0) * kBinReplicas;
    const uint32_t own1 = (KEY64 && ow >= 0) ? (uint32_t)(key >> 32) + 1u : 0u;
    if (q.bin_out != nullptr) {
      bin->own1 = own1;  // export mode: the reduced bin is applied by se3ds_apply_bin
    } else if (q.finalize_bins) {
      const Bin r = bin_fold_and_rearm(bin);
      if (r.zneg) {
        const float bz = f32_unordered(~r.zneg);
        zmin = fminf(zmin, bz);
        if (KEY64 && bz < radw) ow = -1;
      }
      f.x = fmaxf(f.x, (float)r.f[0]); f.y = fmaxf(f.y, (float)r.f[1]); f.z = fmaxf(f.z, (float)r.f[2]);
    } else {
      // more chunks will still add to the global bin: park this pixel's own values in it,
      // patch_owner_kernel finishes the pixel after the last chunk.
      atomicMax(&bin->zneg, ~f32_ordered(zmin));
      atomicMax(&bin->f[0], (int)f.x); atomicMax(&bin->f[1], (int)f.y); atomicMax(&bin->f[2], (int)f.z);
      bin->own1 = own1;
    }
  }
  const float depth = div_rcp(fminf(fmaxf(zmin, 0.0f), q.depth_scale), q.depth_scale, q.inv_depth_scale);
  od = depth;
  if constexpr (COMPACT) {
    *packed = pack_rgb_u8(f);
    return;
  }
  if (q.flags & SE3DS_FLAG_RAW_FEATURES) {
    oi[0] = f.x; oi[1] = f.y; oi[2] = f.z;
  } else {
    oi[0] = clip01_div255(f.x); oi[1] = clip01_div255(f.y); oi[2] = clip01_div255(f.z);
  }
  om = (depth > 0.0f && depth < 1.0f && f.x != -1.0f && f.y != -1.0f && f.z != -1.0f) ? 1.0f : 0.0f;
}

// Re-arms only what was touched (a touched feature buffer entry implies a touched z-buffer entry).
template <bool KEY64>
__device__ __forceinline__ void resolve_rearm4(const FusedParams& q, size_t e, const unsigned long long (&key)[4]) {
  if (key[0] == kZArmed && key[1] == kZArmed && key[2] == kZArmed && key[3] == kZArmed) return;
  const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), zero = make_uint4(0, 0, 0, 0);
  if constexpr (KEY64) st_256(q.zbuf + e, ones, ones);
  else *reinterpret_cast<uint4*>(q.zbuf32 + e) = ones;
  st_256(q.fbuf + e, zero, zero);
}

// COMPACT (SE3DS_FLAG_COMPACT_OUT): out_image is a uint8 (J,H,W,3) plane of clamped maxima, out_mask is not
// written (mask = 0 < depth < 1): 7 instead of 20 bytes per pixel leave the kernel.
template <int PPT, bool KEY64, bool COMPACT>
__global__ void __launch_bounds__(kThreads, KEY64 ? 10 : 12) resolve_kernel(const FusedParams q) {
  pdl_enter();
  constexpr bool STAGED = PPT == 4 && !COMPACT;  // RGB stores go through shared memory (below)
  __shared__ float4 stage[STAGED ? kThreads * 3 : 1];
  __shared__ __align__(16) uint32_t stage8[(PPT == 4 && COMPACT) ? kThreads * 3 : 1];  // the same for the packed colours
  const int lj = blockIdx.z;
  int n, p;
  if (q.PC == 1) { n = q.n0 + lj; p = q.p0; } else { const int a = lj / q.PC; n = q.n0 + a; p = q.p0 + (lj - a * q.PC); }
  const int job = n * q.P + p;
  const int row = blockIdx.y;
  const int col0 = (blockIdx.x * kThreads + threadIdx.x) * PPT;
  if constexpr (PPT == 4) {
    if (((blockIdx.x * kThreads + (threadIdx.x & ~31)) + 32) * PPT > q.W) {  // ragged warp: plain path
      if (col0 >= q.W) return;
    }
  } else {
    if (col0 >= q.W) return;
  }
  const int pix0 = row * q.W + col0;
  const size_t e = (size_t)lj * q.HW + pix0;
  const size_t o = (size_t)job * q.HW + pix0;
  if constexpr (PPT == 4) {
    unsigned long long key[4];
    uint2 fv[4];
    resolve_load4<KEY64>(q, e, key, fv);
    float od[4], om[4], oi[12];
    int ow[4];
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) resolve_pixel<KEY64, COMPACT>(q, job, pix0 + k, key[k], fv[k], od[k], om[k], oi + 3 * k, ow[k], pk + k);
    __stcs(reinterpret_cast<float4*>(q.out_depth + o), make_float4(od[0], od[1], od[2], od[3]));
    if constexpr (COMPACT) {  // 12 bytes of colour per thread, 384 contiguous bytes per warp
      uint32_t* im8 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(q.out_image) + o * 3);
      const uint32_t w0 = pk[0] | (pk[1] << 24), w1 = (pk[1] >> 8) | (pk[2] << 16), w2 = (pk[2] >> 16) | (pk[3] << 8);
      if (((blockIdx.x * kThreads + (threadIdx.x & ~31)) + 32) * PPT <= q.W) {
        // a full warp: through shared memory, so that each store instruction writes 128 contiguous bytes (whole
        // sectors -- what a store that leaves the GPU through NVLink, e.g. to a multicast mapping, wants)
        const unsigned lane = threadIdx.x & 31u;
        uint32_t* wstage = stage8 + (threadIdx.x >> 5) * 96;
        wstage[lane * 3 + 0] = w0;
        wstage[lane * 3 + 1] = w1;
        wstage[lane * 3 + 2] = w2;
        __syncwarp();
        uint32_t* im0 = im8 - lane * 3;
        if ((reinterpret_cast<uintptr_t>(im0) & 15u) == 0) {  // 24 lanes x 16 bytes
          if (lane < 24) __stcs(reinterpret_cast<uint4*>(im0) + lane, reinterpret_cast<const uint4*>(wstage)[lane]);
        } else {
          __stcs(im0 + lane, wstage[lane]);
          __stcs(im0 + 32 + lane, wstage[32 + lane]);
          __stcs(im0 + 64 + lane, wstage[64 + lane]);
        }
      } else {
        __stcs(im8, w0);
        __stcs(im8 + 1, w1);
        __stcs(im8 + 2, w2);
      }
    } else {
    __stcs(reinterpret_cast<float4*>(q.out_mask + o), make_float4(om[0], om[1], om[2], om[3]));
    float4* im = reinterpret_cast<float4*>(q.out_image + o * 3);
    bool direct = true;
    if constexpr (STAGED) {
      // a full warp holds 128 consecutive pixels of one row: pass the RGB (48 B per lane) through
      // shared memory so that each store instruction writes 512 contiguous bytes
      const unsigned lane = threadIdx.x & 31u;
      if (((blockIdx.x * kThreads + (threadIdx.x & ~31)) + 32) * PPT <= q.W) {
        direct = false;
        float4* wstage = stage + (threadIdx.x >> 5) * 96;
        wstage[lane * 3 + 0] = make_float4(oi[0], oi[1], oi[2], oi[3]);
        wstage[lane * 3 + 1] = make_float4(oi[4], oi[5], oi[6], oi[7]);
        wstage[lane * 3 + 2] = make_float4(oi[8], oi[9], oi[10], oi[11]);
        __syncwarp();
        float4* im0 = im - lane * 3;
        __stcs(im0 + lane, wstage[lane]);
        __stcs(im0 + 32 + lane, wstage[32 + lane]);
        __stcs(im0 + 64 + lane, wstage[64 + lane]);
      }
    }
    if (direct) {
      __stcs(im, make_float4(oi[0], oi[1], oi[2], oi[3]));
      __stcs(im + 1, make_float4(oi[4], oi[5], oi[6], oi[7]));
      __stcs(im + 2, make_float4(oi[8], oi[9], oi[10], oi[11]));
    }
    }
    if constexpr (KEY64)
      if (q.out_winner) __stcs(reinterpret_cast<int4*>(q.out_winner + o), make_int4(ow[0], ow[1], ow[2], ow[3]));
    resolve_rearm4<KEY64>(q, e, key);
  } else {
    const unsigned long long key = KEY64 ? q.zbuf[e] : (((unsigned long long)q.zbuf32[e] << 32) | 0xFFFFFFFFu);
    float od, om, oi[3];
    int ow;
    uint32_t pk;
    resolve_pixel<KEY64, COMPACT>(q, job, pix0, key, q.fbuf[e], od, om, oi, ow, &pk);
    q.out_depth[o] = od;
    if constexpr (COMPACT) {
      uint8_t* im8 = reinterpret_cast<uint8_t*>(q.out_image) + o * 3;
      im8[0] = (uint8_t)pk; im8[1] = (uint8_t)(pk >> 8); im8[2] = (uint8_t)(pk >> 16);
    } else {
      q.out_mask[o] = om;
      for (int c = 0; c < 3; ++c) q.out_image[o * 3 + c] = oi[c];
    }
    if constexpr (KEY64)
      if (q.out_winner) q.out_winner[o] = ow;
    if constexpr (KEY64) q.zbuf[e] = kZArmed; else q.zbuf32[e] = 0xFFFFFFFFu;
    q.fbuf[e] = make_uint2(0, 0);
  }
  stamp_end(q.stamps, 2);  // thread 0 of a block never leaves early (its column is inside the image)
}

// After the last chunk of a multi-chunk call with the global bin: finish job 0's pixel (0,0).
__global__ void patch_owner_kernel(const FusedParams q) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Bin r = bin_fold_and_rearm(q.bins);
  const float zmin = r.zneg ? f32_unordered(~r.zneg) : q.depth_scale;
  const float3 f = make_float3((float)r.f[0], (float)r.f[1], (float)r.f[2]);
  if (q.out_winner && r.own1 && zmin < __uint_as_float(r.own1 - 1u)) q.out_winner[0] = -1;  // a rejected point is nearer
  const float depth = div_rcp(fminf(fmaxf(zmin, 0.0f), q.depth_scale), q.depth_scale, q.inv_depth_scale);
  q.out_depth[0] = depth;
  if (q.flags & SE3DS_FLAG_COMPACT_OUT) {
    const uint32_t pk = pack_rgb_u8(f);
    uint8_t* im8 = reinterpret_cast<uint8_t*>(q.out_image);
    im8[0] = (uint8_t)pk; im8[1] = (uint8_t)(pk >> 8); im8[2] = (uint8_t)(pk >> 16);
    return;
  }
  const bool raw = q.flags & SE3DS_FLAG_RAW_FEATURES;
  q.out_image[0] = raw ? f.x : clip01_div255(f.x); q.out_image[1] = raw ? f.y : clip01_div255(f.y);
  q.out_image[2] = raw ? f.z : clip01_div255(f.z);
  q.out_mask[0] = (depth > 0.0f && depth < 1.0f && f.x != -1.0f && f.y != -1.0f && f.z != -1.0f) ? 1.0f : 0.0f;
}

// Export mode (multi-GPU): hand the call's bin to the caller as (min depth | +inf, max R, G, B).
__global__ void export_bin_kernel(const FusedParams q) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Bin r = bin_fold_and_rearm(q.bins);
  q.bin_out[0] = r.zneg ? f32_unordered(~r.zneg) : __int_as_float(0x7f800000);
  q.bin_out[1] = (float)r.f[0]; q.bin_out[2] = (float)r.f[1]; q.bin_out[3] = (float)r.f[2];
  q.bin_out[4] = r.own1 ? __uint_as_float(r.own1 - 1u) : __int_as_float(0x7f800000);  // depth of the owner pixel's own winner
}

// Applies a (reduced) bin to pixel (0,0) of the first job of finished guidance tensors.
__global__ void apply_bin_kernel(const float* bin, float depth_scale, bool raw, bool compact, float* image, float* depth, float* mask, int32_t* winner) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (winner && bin[0] < bin[4]) winner[0] = -1;  // a rejected point is nearer than the pixel's own winner
  const float d = fminf(depth[0], __fdiv_rn(fminf(fmaxf(bin[0], 0.0f), depth_scale), depth_scale));
  depth[0] = d;
  if (compact) {  // uint8 colours: max of the clamped maxima
    uint8_t* im8 = reinterpret_cast<uint8_t*>(image);
    for (int c = 0; c < 3; ++c) im8[c] = (uint8_t)max((int)im8[c], min(max(__float2int_rz(bin[1 + c]), 0), 255));
    return;
  }
  float f[3];
  // raw: the outputs hold raw per-channel maxima (SE3DS_FLAG_RAW_FEATURES), else clip(x / 255, 0, 1)
  for (int c = 0; c < 3; ++c) { f[c] = fmaxf(image[c], raw ? bin[1 + c] : clip01_div255(bin[1 + c])); image[c] = f[c]; }
  mask[0] = (d > 0.0f && d < 1.0f && f[0] != -1.0f && f[1] != -1.0f && f[2] != -1.0f) ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------------------------------
// Compat path kernels (materialised clouds, float32 features, any channel count)
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T cast_void(double v) { return (T)v; }

// utils/pano_utils.py:245-265
template <typename T>
__global__ void mask_pano_kernel(const T* __restrict__ in, T* __restrict__ out, long long total, int H,
                                 long long row_elems, int mh, T value) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)((i / row_elems) % H);
    out[i] = (row >= mh && row <= H - mh) ? in[i] : value;
  }
}

// utils/pano_utils.py:164-242: xyz1 (N,4,HW) planar + filtered features
template <typename TI, typename TO>
__global__ void unproject_kernel(const TI* __restrict__ feats, const float* __restrict__ depth,
                                 const float* __restrict__ tab, int N, int H, int W, int C,
                                 float depth_scale, TO void_value, float* __restrict__ xyz1,
                                 TO* __restrict__ feats_out) {
  const int HW = H * W;
  const long long total = (long long)N * HW;
  const float *sin_e = tab, *cos_e = tab + H, *sin_h = tab + 2 * H, *cos_h = sin_h + W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW), pix = (int)(i - (long long)b * HW);
    const int row = pix / W, col = pix - row * W;
    const float d = depth[i];
    const bool dvalid = d > 0.0f && d < 1.0f;
    const float rad0 = __fmul_rn(__fmul_rn(d, depth_scale), dvalid ? 1.0f : 0.0f);
    const float t = __fmul_rn(rad0, sin_e[row]);
    float* o = xyz1 + (size_t)b * 4 * HW + pix;
    o[0] = __fmul_rn(t, cos_h[col]);
    o[HW] = __fmul_rn(t, sin_h[col]);
    o[2 * (size_t)HW] = __fmul_rn(rad0, cos_e[row]);
    o[3 * (size_t)HW] = 1.0f;
    for (int c = 0; c < C; ++c) feats_out[i * C + c] = dvalid ? (TO)feats[i * C + c] : void_value;
  }
}

struct CloudParams {
  const float* coords;  // (N,4,M)
  const void* feats;    // (N,M,C)
  unsigned long long* zbuf;  // KEY64: depth bits << 32 | point index << 1
  uint32_t* zbuf32;          // !KEY64: depth bits
  uint2* fbuf;               // F16 mode: per-channel maxima of small integer features (float16 x 4, armed 0)
  float* fbuf32;             // F16 mode: side accumulator (N*HW*3, armed void_out) for values float16 cannot hold
  uint32_t* used32;          // F16 mode: != 0 once any point went to the side accumulator
  uint32_t* sc_flat;
  float* sc_rad;
  uint32_t* bin;  // [0] = ~ordered(min depth), [1+c] = ordered(max feature c); all-zero = armed
  float* depth_out;
  float* feats_out;  // (N,H,W,C); legacy mode: pre-filled with output_void_class and accumulated in place
  int* winner_out;
  long long M;
  int N, C, H, W, HW, mode;
  float void_in, void_out, depth_scale;
  FastProj fast;
};

__device__ __forceinline__ float load_feat(const uint8_t* f, size_t i) { return (float)f[i]; }
__device__ __forceinline__ float load_feat(const int* f, size_t i) { return (float)f[i]; }
__device__ __forceinline__ float load_feat(const float* f, size_t i) { return f[i]; }

// float max through integer atomics (works for mixed signs; -0.0 never replaces +0.0)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// The compat path (utils/pano_utils.py:117-161 / utils/point_cloud_utils.py:90-183 on a materialised cloud),
// three kernels chained with programmatic dependent launch:
//   cloud_depth : projection (mode 0: certified fast projection with the canonical one as the fallback of the
//                 lanes it cannot certify; mode 1: the coordinates are already "transformed") + REDG.MIN of the
//                 64-bit (depth | index) key when winner indices are wanted, of the 32-bit depth otherwise
//   cloud_feat  : tolerance test + per-channel max.  F16 mode (3 channels of integer features, output void 0:
//                 every RGB caller of the reference): one 8-byte float16x4 reduction per point, like the fused
//                 path; a value float16 cannot hold exactly goes to a float32 side accumulator instead, and the
//                 resolve merges the two (a maximum can be taken in pieces).  Otherwise: float32 atomics into
//                 the pre-filled output.
//   cloud_resolve: depth, features, winner; re-arms what was touched.
template <typename T, bool KEY64>
__global__ void __launch_bounds__(kThreads) cloud_depth_kernel(const CloudParams q) {
  pdl_enter();
  const int b = blockIdx.y;
  // Persistent grid-stride loop: the warp keeps the largest reject-bin value it has seen in a register and
  // goes back to memory only when one of its points exceeds it (same-address loads from every warp of a
  // one-point-per-thread grid serialise at one L2 slice: 70 % of the kernel's stalls, profiles/r02_*).
  uint32_t bin_seen = 0u;
  const long long span = (long long)gridDim.x * kThreads;
  for (long long m0 = blockIdx.x * (long long)kThreads; m0 < q.M; m0 += span) {
    const long long m = m0 + threadIdx.x;
    const bool on = m < q.M;
    float pz = 0.f;
    int tpix = -1;
    bool fvalid = true;
    if (on) {
      const float* cb = q.coords + (size_t)b * 4 * q.M;
      const float x = __ldg(cb + m), y = __ldg(cb + q.M + m), z = __ldg(cb + 2 * q.M + m);
      const size_t f0 = ((size_t)b * q.M + m) * q.C;
      for (int c = 0; c < q.C; ++c) fvalid &= load_feat(static_cast<const T*>(q.feats), f0 + c) != q.void_in;
      if (q.mode == 0) {
        const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        float rs, fx;
        bool ok;
        int frow;
        pz = fast_rad(r2, rs, ok);
        const bool certain = project_pixel_fast(x, y, z, rs, q.H, q.W, q.fast, tpix, fx, frow) && ok;
        if (!certain) {  // a few lanes in a thousand: the canonical projection
          pz = canon_rad(x, y, z);
          tpix = project_pixel_rad(x, y, z, q.H, q.W, pz);
        }
      } else {
        pz = z;
        tpix = pixel_of(x, y, z, q.H, q.W);
      }
    }
    const bool valid = on && fvalid && tpix >= 0;
    if (valid) {
      if constexpr (KEY64) atomicMin(q.zbuf + (size_t)b * q.HW + tpix, ((unsigned long long)__float_as_uint(pz) << 32) | ((uint32_t)m << 1));
      else atomicMin(q.zbuf32 + (size_t)b * q.HW + tpix, __float_as_uint(pz));
    }
    // reject bin: min depth of every rejected point (utils/point_cloud_utils.py:150-159)
    const bool rej = on && !valid;
    if (__any_sync(0xffffffffu, rej)) {
      const uint32_t v = __reduce_max_sync(0xffffffffu, (rej && pz == pz) ? ~f32_ordered(pz) : 0u);  // a NaN depth never lowers the minimum
      if (v > bin_seen) {
        if ((threadIdx.x & 31) == 0) {
          bin_seen = __ldcg(q.bin);
          if (v > bin_seen) { atomicMax(q.bin, v); bin_seen = v; }
        }
        bin_seen = __shfl_sync(0xffffffffu, bin_seen, 0);
      }
    }
    if (on) {
      q.sc_flat[(size_t)b * q.M + m] = valid ? (uint32_t)tpix : kScInvalid;
      q.sc_rad[(size_t)b * q.M + m] = pz;
    }
  }
}

template <typename T, bool KEY64, bool F16>
__global__ void __launch_bounds__(kThreads) cloud_feat_kernel(const CloudParams q) {
  pdl_enter();
  const int b = blockIdx.y;
  const T* feats = static_cast<const T*>(q.feats);
  uint32_t bin_seen[3] = {0u, 0u, 0u};  // register copy of the reject bin's first channels (see cloud_depth_kernel)
  const long long span = (long long)gridDim.x * kThreads;
  for (long long m0 = blockIdx.x * (long long)kThreads; m0 < q.M; m0 += span) {
    const long long m = m0 + threadIdx.x;
    const bool on = m < q.M;
    bool keep = false;
    uint32_t fl = kScInvalid;
    if (on) {
      fl = q.sc_flat[(size_t)b * q.M + m];
      const float rad = q.sc_rad[(size_t)b * q.M + m];
      if (!(fl & kScInvalid)) {
        float zmin;
        if constexpr (KEY64) zmin = __uint_as_float((uint32_t)(__ldcg(q.zbuf + (size_t)b * q.HW + fl) >> 32));
        else zmin = __uint_as_float(__ldcg(q.zbuf32 + (size_t)b * q.HW + fl));
        keep = rad < __fadd_rn(fminf(zmin, q.depth_scale), 0.1f);  // armed bits are a NaN: fminf -> depth_scale
      }
    }
    const size_t f0 = ((size_t)b * q.M + (on ? m : 0)) * q.C;
    if (keep) {
      if constexpr (F16) {
        const int3 f = make_int3((int)feats[f0], (int)feats[f0 + 1], (int)feats[f0 + 2]);
        if (max(max(abs(f.x), abs(f.y)), abs(f.z)) <= 2048) {
          if (f.x > 0 || f.y > 0 || f.z > 0) red_max_f16x4(q.fbuf + (size_t)b * q.HW + fl, pack_f16x4(f));  // <= 0 cannot raise a maximum over 0
        } else {  // float16 would round it: the float32 side accumulator
          float* dst = q.fbuf32 + ((size_t)b * q.HW + fl) * 3;
          atomic_max_f32(dst, (float)f.x); atomic_max_f32(dst + 1, (float)f.y); atomic_max_f32(dst + 2, (float)f.z);
          *q.used32 = 1u;
        }
      } else {
        float* dst = q.feats_out + ((size_t)b * q.HW + fl) * q.C;
        for (int c = 0; c < q.C; ++c) atomic_max_f32(dst + c, load_feat(feats, f0 + c));
      }
    }
    // reject bin: per-channel maximum of every rejected point (flat index 0 of the reference), reduced per warp;
    // memory is consulted only when a value exceeds what this warp has already seen there
    const bool rej = on && !keep;
    if (__any_sync(0xffffffffu, rej)) {
      for (int c = 0; c < q.C; ++c) {
        const uint32_t v = __reduce_max_sync(0xffffffffu, rej ? f32_ordered(load_feat(feats, f0 + c)) : 0u);
        const uint32_t seen = c < 3 ? bin_seen[c] : 0u;
        if (v > seen) {
          uint32_t now = v;
          if ((threadIdx.x & 31) == 0) {
            now = __ldcg(q.bin + 1 + c);
            if (v > now) { atomicMax(q.bin + 1 + c, v); now = v; }
          }
          now = __shfl_sync(0xffffffffu, now, 0);
          if (c < 3) bin_seen[c] = now;
        }
      }
    }
  }
}

template <bool KEY64, bool F16>
__global__ void __launch_bounds__(kThreads) cloud_resolve_kernel(const CloudParams q) {
  pdl_enter();
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const long long total = (long long)q.N * q.HW;
  if (i >= total) return;
  unsigned long long key;
  if constexpr (KEY64) key = q.zbuf[i];
  else key = ((unsigned long long)q.zbuf32[i] << 32) | 0xFFFFFFFFu;
  const bool has = key != kZArmed;
  const float radw = __uint_as_float((uint32_t)(key >> 32));
  float zmin = has ? fminf(radw, q.depth_scale) : q.depth_scale;
  float f[3] = {0.f, 0.f, 0.f};
  if constexpr (F16) {
    const uint2 fv = q.fbuf[i];
    const float3 u = unpack_f16x4(fv);
    f[0] = u.x; f[1] = u.y; f[2] = u.z;
    if (*q.used32) {  // some value of this call did not fit float16: merge (and re-arm) the side accumulator
      for (int c = 0; c < 3; ++c) {
        f[c] = fmaxf(f[c], q.fbuf32[i * 3 + c]);
        q.fbuf32[i * 3 + c] = q.void_out;
      }
    }
    if (fv.x | fv.y) q.fbuf[i] = make_uint2(0u, 0u);
  }
  if (i == 0) {
    if (q.bin[0]) zmin = fminf(zmin, f32_unordered(~q.bin[0]));
    for (int c = 0; c < q.C; ++c) {
      if (q.bin[1 + c]) {
        const float bf = f32_unordered(q.bin[1 + c]);
        if constexpr (F16) f[c] = fmaxf(f[c], bf);
        else q.feats_out[c] = fmaxf(q.feats_out[c], bf);
      }
      q.bin[1 + c] = 0u;
    }
    q.bin[0] = 0u;
  }
  if constexpr (F16) {
    q.feats_out[i * 3] = f[0]; q.feats_out[i * 3 + 1] = f[1]; q.feats_out[i * 3 + 2] = f[2];
  }
  q.depth_out[i] = __fdiv_rn(fminf(fmaxf(zmin, 0.0f), q.depth_scale), q.depth_scale);
  if constexpr (KEY64) {
    if (q.winner_out) q.winner_out[i] = (has && radw <= q.depth_scale) ? (int)((uint32_t)key >> 1) : -1;
    if (has) q.zbuf[i] = kZArmed;
  } else {
    if (has) q.zbuf32[i] = 0xFFFFFFFFu;
  }
}

// re-arms the `used32` flag after the resolve (one thread; launched only in F16 mode)
__global__ void cloud_clear_flag_kernel(uint32_t* flag) {
  pdl_enter();
  if (threadIdx.x == 0 && blockIdx.x == 0) *flag = 0u;
}

// tensorflow_addons.image.interpolate_bilinear (called at utils/pano_utils.py:339,412,472):
// grid (B,H,W,C) f32, queries (B,Nq,2) f32 -> out (B,Nq,C).  xy = 1: queries are (x, y), else (y, x).
// Floors clamp to [0, size-2], alphas to [0,1]; a_y * (bottom - top) + top, top = a_x * (tr - tl) + tl.
__global__ void __launch_bounds__(kThreads) interpolate_bilinear_kernel(const float* __restrict__ grid,
                                                                       const float* __restrict__ queries, int B, int H,
                                                                       int W, int C, long long Nq, int xy,
                                                                       float* __restrict__ out) {
  const long long total = (long long)B * Nq;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int b = (int)(i / Nq);
    const float2 qq = __ldg(reinterpret_cast<const float2*>(queries) + i);
    const float qy = xy ? qq.y : qq.x, qx = xy ? qq.x : qq.y;
    const float fy = fminf(fmaxf(0.0f, floorf(qy)), (float)(H - 2)), fx = fminf(fmaxf(0.0f, floorf(qx)), (float)(W - 2));
    const float ay = fminf(fmaxf(0.0f, __fsub_rn(qy, fy)), 1.0f), ax = fminf(fmaxf(0.0f, __fsub_rn(qx, fx)), 1.0f);
    const int y0 = (int)fy, x0 = (int)fx;
    const float* tl = grid + (((size_t)b * H + y0) * W + x0) * C;
    const float* bl = tl + (size_t)W * C;
    float* o = out + (size_t)i * C;
    for (int c = 0; c < C; ++c) {
      const float vtl = __ldg(tl + c), vtr = __ldg(tl + C + c), vbl = __ldg(bl + c), vbr = __ldg(bl + C + c);
      const float top = __fadd_rn(__fmul_rn(ax, __fsub_rn(vtr, vtl)), vtl);
      const float bot = __fadd_rn(__fmul_rn(ax, __fsub_rn(vbr, vbl)), vbl);
      o[c] = __fadd_rn(__fmul_rn(ay, __fsub_rn(bot, top)), top);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused resampling kernels: query coordinates are computed in-kernel and fed to
// the same tfa-style bilinear sample, one thread per output pixel.
// ------------------------------------------------------------------------------------------
struct Mat3 {
  float m[9];
};

// tf.linspace value i of n in float32: exact end points, start + delta * i inside
__device__ __forceinline__ float linspace_at(float start, float stop, int n, int i) {
  if (i == 0 || n == 1) return start;
  if (i == n - 1) return stop;
  return __fadd_rn(start, __fmul_rn(__fdiv_rn(__fsub_rn(stop, start), (float)(n - 1)), (float)i));
}

// equirectangular_pixel_rays (utils/pano_utils.py:92-114) for pixel (r, c) of an oh x 2*oh image
__device__ __forceinline__ float3 pixel_ray(int oh, int r, int c) {
  const float kPi = 3.14159274101257324f;
  const float heading = linspace_at(-kPi, kPi, 2 * oh, c), pitch = linspace_at(0.0f, kPi, oh, r);
  const float sp = sinf(pitch);
  return make_float3(__fmul_rn(sp, sinf(heading)), -cosf(pitch), __fmul_rn(sp, cosf(heading)));
}

// tfa interpolate_bilinear of one query on an image with an optional virtual 1-pixel constant
// border (pad = 1: the image behaves like tf.pad(image, 1, constant_values=pad_value)).
__device__ __forceinline__ void bilinear_sample(const float* __restrict__ img, int H, int W, int C, float qy, float qx,
                                                int pad, float pad_value, float* __restrict__ o) {
  const int PH = H + 2 * pad, PW = W + 2 * pad;
  const float fy = fminf(fmaxf(0.0f, floorf(qy)), (float)(PH - 2)), fx = fminf(fmaxf(0.0f, floorf(qx)), (float)(PW - 2));
  const float ay = fminf(fmaxf(0.0f, __fsub_rn(qy, fy)), 1.0f), ax = fminf(fmaxf(0.0f, __fsub_rn(qx, fx)), 1.0f);
  const int y0 = (int)fy - pad, x0 = (int)fx - pad;
  for (int c = 0; c < C; ++c) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + (k >> 1), x = x0 + (k & 1);
      v[k] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + ((size_t)y * W + x) * C + c) : pad_value;
    }
    const float top = __fadd_rn(__fmul_rn(ax, __fsub_rn(v[1], v[0])), v[0]);
    const float bot = __fadd_rn(__fmul_rn(ax, __fsub_rn(v[3], v[2])), v[2]);
    o[c] = __fadd_rn(__fmul_rn(ay, __fsub_rn(bot, top)), top);
  }
}

__device__ __forceinline__ float3 mat3_mul(const float* m, float3 v) {
  return make_float3(__fadd_rn(__fadd_rn(__fmul_rn(m[0], v.x), __fmul_rn(m[1], v.y)), __fmul_rn(m[2], v.z)),
                     __fadd_rn(__fadd_rn(__fmul_rn(m[3], v.x), __fmul_rn(m[4], v.y)), __fmul_rn(m[5], v.z)),
                     __fadd_rn(__fadd_rn(__fmul_rn(m[6], v.x), __fmul_rn(m[7], v.y)), __fmul_rn(m[8], v.z)));
}

// rotate_pano (utils/pano_utils.py:306-341): pano (N,H,W,C), matrix (N,3,3) -> (N,OH,2*OH,C)
__global__ void __launch_bounds__(kThreads) rotate_pano_kernel(const float* __restrict__ pano, const float* __restrict__ matrix,
                                                              int N, int H, int W, int C, int OH, float* __restrict__ out) {
  const float kPi = 3.14159274101257324f, kTwoPi = 6.28318548202514648f;
  const int OW = 2 * OH;
  const long long total = (long long)N * OH * OW;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int c = (int)(i % OW), r = (int)((i / OW) % OH), b = (int)(i / ((long long)OW * OH));
    const float3 v = mat3_mul(matrix + (size_t)b * 9, pixel_ray(OH, r, c));
    const float pitch = acosf(-v.y), heading = atan2f(v.x, v.z);
    const float qx = __fmul_rn(__fadd_rn(__fdiv_rn(heading, kTwoPi), 0.5f), (float)(W - 1));
    const float qy = __fmul_rn(__fdiv_rn(pitch, kPi), (float)(H - 1));
    bilinear_sample(pano + (size_t)b * H * W * C, H, W, C, qy, qx, 0, 0.0f, out + (size_t)i * C);
  }
}

// project_perspective_image (utils/pano_utils.py:344-417): image (h,w,C) -> equirect (OH,2*OH,C).
// w2i = world-to-image transform; pad = 1 for the 'constant' / 'mean' modes (virtual border of
// pad_value), 0 for 'reflect' (no border, clamped sampling).
__global__ void __launch_bounds__(kThreads) persp_to_equirect_kernel(const float* __restrict__ image, Mat3 w2i, int H, int W,
                                                                    int C, int OH, int pad, float pad_value, int round_nearest,
                                                                    float* __restrict__ out) {
  const int OW = 2 * OH;
  const long long total = (long long)OH * OW;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int c = (int)(i % OW), r = (int)(i / OW);
    const float3 p = mat3_mul(w2i.m, pixel_ray(OH, r, c));
    float qx = -1.0f, qy = -1.0f;
    if (p.z > 0.0f) { qx = __fdiv_rn(p.x, p.z); qy = __fdiv_rn(p.y, p.z); }
    if (round_nearest) { qx = rintf(qx); qy = rintf(qy); }  // tf.math.round: half to even
    if (pad) { qx = __fadd_rn(qx, 1.0f); qy = __fadd_rn(qy, 1.0f); }
    bilinear_sample(image, H, W, C, qy, qx, pad, pad_value, out + (size_t)i * C);
  }
}

// get_perspective_from_equirectangular_image (utils/pano_utils.py:443-476):
// equirect (EH,EW,C) -> perspective (PH,PW,C); kinv_t = inv(K)^T, rot = rotation matrix, applied as
// row vectors ((x, y, 1) @ kinv_t) @ rot like the reference.
__global__ void __launch_bounds__(kThreads) equirect_to_persp_kernel(const float* __restrict__ image, Mat3 kinv_t, Mat3 rot,
                                                                    int EH, int EW, int C, int PH, int PW,
                                                                    float* __restrict__ out) {
  const float kPi = 3.14159274101257324f, kTwoPi = 6.28318548202514648f;
  const long long total = (long long)PH * PW;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const float x = (float)(i % PW), y = (float)(i / PW);
    // row vector times matrix: out_j = sum_k v_k * M[k][j]
    const float* a = kinv_t.m;
    const float3 t = make_float3(__fadd_rn(__fadd_rn(__fmul_rn(x, a[0]), __fmul_rn(y, a[3])), a[6]),
                                 __fadd_rn(__fadd_rn(__fmul_rn(x, a[1]), __fmul_rn(y, a[4])), a[7]),
                                 __fadd_rn(__fadd_rn(__fmul_rn(x, a[2]), __fmul_rn(y, a[5])), a[8]));
    const float* m = rot.m;
    float3 v = make_float3(__fadd_rn(__fadd_rn(__fmul_rn(t.x, m[0]), __fmul_rn(t.y, m[3])), __fmul_rn(t.z, m[6])),
                           __fadd_rn(__fadd_rn(__fmul_rn(t.x, m[1]), __fmul_rn(t.y, m[4])), __fmul_rn(t.z, m[7])),
                           __fadd_rn(__fadd_rn(__fmul_rn(t.x, m[2]), __fmul_rn(t.y, m[5])), __fmul_rn(t.z, m[8])));
    const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
    v.x = __fdiv_rn(v.x, norm); v.y = __fdiv_rn(v.y, norm); v.z = __fdiv_rn(v.z, norm);
    const float lon = atan2f(v.x, v.z), lat = asinf(v.y);
    const float qx = __fmul_rn(__fadd_rn(__fdiv_rn(lon, kTwoPi), 0.5f), (float)(EW - 1));
    const float qy = __fmul_rn(__fadd_rn(__fdiv_rn(lat, kPi), 0.5f), (float)(EH - 1));
    bilinear_sample(image, EH, EW, C, qy, qx, 0, 0.0f, out + (size_t)i * C);
  }
}

// equirectangular_pixel_rays as a tensor (3, OH * 2*OH)
__global__ void __launch_bounds__(kThreads) pixel_rays_kernel(int OH, float* __restrict__ out) {
  const int OW = 2 * OH;
  const long long total = (long long)OH * OW;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const float3 v = pixel_ray(OH, (int)(i / OW), (int)(i % OW));
    out[i] = v.x; out[total + i] = v.y; out[2 * total + i] = v.z;
  }
}

// tf.image.resize with half_pixel_centers (TF 2.x), used by equirectangular_to_pointcloud when
// size_mult != 1 (utils/pano_utils.py:203-208).  NEAREST keeps the dtype:
// in = min(floor((out + 0.5) * scale), in_size - 1); BILINEAR returns float32:
// in = (out + 0.5) * scale - 0.5, lower = max(floor(in), 0), upper = min(ceil(in), in_size - 1).
template <typename TI, typename TO, bool BILINEAR>
__global__ void __launch_bounds__(kThreads) resize_kernel(const TI* __restrict__ in, int N, int H, int W, int C, int OH,
                                                         int OW, TO* __restrict__ out) {
  const float sy = __fdiv_rn((float)H, (float)OH), sx = __fdiv_rn((float)W, (float)OW);
  const long long total = (long long)N * OH * OW;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), b = (int)(i / ((long long)OW * OH));
    const TI* img = in + (size_t)b * H * W * C;
    TO* o = out + (size_t)i * C;
    if constexpr (!BILINEAR) {
      const int iy = min((int)floorf(__fmul_rn(__fadd_rn((float)oy, 0.5f), sy)), H - 1);
      const int ix = min((int)floorf(__fmul_rn(__fadd_rn((float)ox, 0.5f), sx)), W - 1);
      for (int c = 0; c < C; ++c) o[c] = (TO)img[((size_t)iy * W + ix) * C + c];
    } else {
      const float py = __fsub_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), sy), 0.5f), px = __fsub_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), sx), 0.5f);
      const float fy = floorf(py), fx = floorf(px);
      const int y0 = max((int)fy, 0), y1 = min((int)ceilf(py), H - 1), x0 = max((int)fx, 0), x1 = min((int)ceilf(px), W - 1);
      const float ly = __fsub_rn(py, fy), lx = __fsub_rn(px, fx);
      for (int c = 0; c < C; ++c) {
        const float tl = (float)img[((size_t)y0 * W + x0) * C + c], tr = (float)img[((size_t)y0 * W + x1) * C + c];
        const float bl = (float)img[((size_t)y1 * W + x0) * C + c], br = (float)img[((size_t)y1 * W + x1) * C + c];
        const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
        const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
        o[c] = (TO)__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
      }
    }
  }
}

// get_filtered_coords_and_feats (utils/point_cloud_utils.py:32-87, legacy perspective unprojection):
// xs / ys = float32(tf.linspace(-1, 1, n) evaluated in float64), xyz = (xs*d, ys*d, d, 1) * mask,
// then inv(K) @ xyz with K = diag(k, k, 1, 1); features are zeroed where the depth is invalid.
template <typename T>
__global__ void __launch_bounds__(kThreads) filtered_coords_kernel(const T* __restrict__ feats, const float* __restrict__ depth,
                                                                  int N, int H, int W, int C, float depth_scale, float kinv_x,
                                                                  float kinv_y, float* __restrict__ xyz, float* __restrict__ out) {
  const int HW = H * W;
  const long long total = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int b = (int)(i / HW), pix = (int)(i - (long long)b * HW);
    const int r = pix / W, c = pix - r * W;
    const float xs = (c == W - 1 || W == 1) ? (W == 1 ? -1.0f : 1.0f) : (float)(-1.0 + (2.0 / (double)(W - 1)) * (double)c);
    const float ys = (r == H - 1 || H == 1) ? (H == 1 ? -1.0f : 1.0f) : (float)(-1.0 + (2.0 / (double)(H - 1)) * (double)r);
    const float d = __fmul_rn(depth[i], depth_scale);
    const bool valid = d > 0.0f && d < depth_scale;
    const float m = valid ? 1.0f : 0.0f;
    float* o = xyz + (size_t)b * 4 * HW + pix;
    o[0] = __fmul_rn(kinv_x, __fmul_rn(__fmul_rn(xs, d), m));
    o[HW] = __fmul_rn(kinv_y, __fmul_rn(__fmul_rn(ys, d), m));
    o[2 * (size_t)HW] = __fmul_rn(d, m);
    o[3 * (size_t)HW] = m;
    for (int ch = 0; ch < C; ++ch) out[i * C + ch] = valid ? (float)feats[i * C + ch] : 0.0f;
  }
}

// Feedback of a generated frame into the memory (trainers/gan_manager.py:539-542,
// utils/eval_metric.py:227-230): clip_by_value(cast(image * 255, int32), -1, 255); the cast truncates
// toward zero (NaN / out of range -> INT_MIN as on x86, then clipped to -1).  The destination is a frame
// slot of a ring: item n of the dense input goes to out + n * out_item_stride.
__global__ void __launch_bounds__(kThreads) quantize_rgb_kernel(const float* __restrict__ image, long long per_item, int n,
                                                               int* __restrict__ out, long long out_item_stride) {
  const long long total = per_item * n;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long b = i / per_item, e = i - b * per_item;
    out[b * out_item_stride + e] = min(max(cast_i32(__fmul_rn(image[i], 255.0f)), -1), 255);
  }
}

// Compact guidance (uint8 colours + float32 depth) -> the float32 contract of the reference
// (models/models.py:282-293): image = clip(x / 255, 0, 1) with the exact division sequence of the resolve
// kernel, mask = 0 < depth < 1 (the colour condition of the mask never fails: the feature buffer starts at
// the output void class 0 and only takes maxima).  4 pixels per thread when vec (px_per_job % 4 == 0 and
// aligned pointers).  job_map (or nullptr = identity) sends source job s to destination job job_map[s]
// (a gather buffer that arrived piece by piece is put into job order on the way); depth_out (or nullptr)
// receives the depth plane in destination order.
__global__ void __launch_bounds__(kThreads) expand_guidance_kernel(const uint8_t* __restrict__ rgb8, const float* __restrict__ depth,
                                                                  long long njobs, long long px_per_job, int vec,
                                                                  const int* __restrict__ job_map, float* __restrict__ image,
                                                                  float* __restrict__ depth_out, float* __restrict__ mask) {
  const long long units_per_job = vec ? px_per_job >> 2 : px_per_job;
  const long long total = njobs * units_per_job;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long sj = i / units_per_job, u = i - sj * units_per_job;
    const long long dj = job_map ? (long long)__ldg(job_map + sj) : sj;
    if (vec) {
      const long long sq = sj * units_per_job + u, dq = dj * units_per_job + u;  // quad index in source / destination
      const uint32_t* c4 = reinterpret_cast<const uint32_t*>(rgb8) + 3 * sq;    // 12 bytes = 4 pixels
      const uint32_t w[3] = {__ldg(c4), __ldg(c4 + 1), __ldg(c4 + 2)};
      const float4 d = __ldg(reinterpret_cast<const float4*>(depth) + sq);
      float o[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) o[k] = clip01_div255((float)((w[k >> 2] >> (8 * (k & 3))) & 255u));
      float4* im = reinterpret_cast<float4*>(image) + 3 * dq;
      __stcs(im, make_float4(o[0], o[1], o[2], o[3]));
      __stcs(im + 1, make_float4(o[4], o[5], o[6], o[7]));
      __stcs(im + 2, make_float4(o[8], o[9], o[10], o[11]));
      if (depth_out) __stcs(reinterpret_cast<float4*>(depth_out) + dq, d);
      __stcs(reinterpret_cast<float4*>(mask) + dq, make_float4((d.x > 0.0f && d.x < 1.0f) ? 1.0f : 0.0f, (d.y > 0.0f && d.y < 1.0f) ? 1.0f : 0.0f,
                                                               (d.z > 0.0f && d.z < 1.0f) ? 1.0f : 0.0f, (d.w > 0.0f && d.w < 1.0f) ? 1.0f : 0.0f));
    } else {
      const long long sp = sj * px_per_job + u, dp = dj * px_per_job + u;
      for (int c = 0; c < 3; ++c) image[dp * 3 + c] = clip01_div255((float)rgb8[sp * 3 + c]);
      const float d = depth[sp];
      if (depth_out) depth_out[dp] = d;
      mask[dp] = (d > 0.0f && d < 1.0f) ? 1.0f : 0.0f;
    }
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------------------------------
// inference/perturbation_utils.py:23-71, one block per candidate offset
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) proportion_invalid_kernel(const float* __restrict__ offsets,
                                                                     const float* __restrict__ depth, int H, int W,
                                                                     float pad, float depth_scale,
                                                                     float* __restrict__ out) {
  const float ox = offsets[blockIdx.x * 3], oy = offsets[blockIdx.x * 3 + 1], oz = offsets[blockIdx.x * 3 + 2];
  const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz)));
  float heading = canon_atan2f(-ox, -oy);
  if (heading < 0.0f) heading = __fadd_rn(heading, CANON_TWO_PI);
  const float hprop = __fdiv_rn(heading, CANON_TWO_PI);
  // delta_xy: float32 squares and sum, float64 sqrt, back to float32 (perturbation_utils.py:46-47)
  const float dxy = (float)sqrt((double)__fadd_rn(__fmul_rn(oy, oy), __fmul_rn(ox, ox)));
  float elev = canon_atan2f(dxy, -oz);
  if (elev < 0.0f) elev = __fadd_rn(elev, CANON_PI_HI);
  const float eprop = __fdiv_rn(elev, CANON_PI_HI);
  const int hs = (int)__fmul_rn(hprop, (float)W), es = (int)__fmul_rn(eprop, (float)H);
  const int tw = (int)(30.0 / 360.0 * W), th = (int)(60.0 / 180.0 * H);
  const int r0 = max(0, es - th), r1 = min(H, es + th), c0 = max(0, hs - tw), c1 = min(W, hs + tw);
  const int rw = max(0, c1 - c0), rh = max(0, r1 - r0);
  const float thr = __fadd_rn(dist, pad);
  int cnt = 0;
  for (int i = threadIdx.x; i < rw * rh; i += kThreads) {
    const int r = r0 + i / rw, c = c0 + i % rw;
    cnt += __fmul_rn(depth[(size_t)r * W + c], depth_scale) < thr;
  }
  __shared__ int sm[kThreads / 32];
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < kThreads / 32; ++i) tot += sm[i];
    // np.mean of a bool array: float64 count / size (nan for an empty window, like numpy)
    out[blockIdx.x] = (float)((double)tot / (double)(rw * rh));
  }
}

}  // namespace se3ds
