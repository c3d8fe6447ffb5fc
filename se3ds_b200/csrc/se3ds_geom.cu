// se3ds_geom.cu -- C ABI (include/se3ds_geom.h) over the sm_100a kernels in kernels.cuh.
// Host-side duties only: argument validation with the reference's error conditions, workspace
// (z-buffer / feature buffer / scratch / bins / angle tables), job chunking, kernel launches.
// There is no CPU fallback: without a CUDA device every entry point fails with SE3DS_ERR_CUDA.
#include "se3ds_geom.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <utility>
#include <vector>

#include "kernels.cuh"

using namespace se3ds;

namespace {

thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

#define GUARD(device)        \
  DeviceGuard guard_(device); \
  CU(guard_.err)

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(SE3DS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                    \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Every entry point runs on the device of its workspace (or of its tensors) and leaves the caller's
// current device as it found it -- a process that drives several GPUs (or PyTorch's own device
// bookkeeping) must not see it change behind its back.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && device >= 0 && device != prev) err = cudaSetDevice(device);
    else if (device < 0 || device == prev) prev = -1;  // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// device that owns a device pointer (-1: not a device pointer / unknown: stay on the current device)
int device_of(const void* p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

struct TableEntry {
  int h, w;
  float margin;  // certification margin scale the row table was built for
  float* dev;    // sin/cos of the row elevations (2 H), of the column headings (2 W), row certification table (2 H)
};

constexpr size_t kMaxStampChunks = 2048;
constexpr size_t kStampWords = 3 * kStampSlots;  // per chunk: three kernels x kStampSlots
constexpr size_t kDefaultMaxBytes = (size_t)2 << 30;
constexpr size_t kDefaultChunkBytes = (size_t)112 << 20;

}  // namespace

struct se3ds_ws {
  int device = 0, sm_count = 148;
  size_t max_bytes = kDefaultMaxBytes, chunk_bytes = kDefaultChunkBytes;
  DevBuf zbuf, zbuf32, fbuf, scf, scr, bins, cbin, cf32, cflag;  // cf32 / cflag: float32 side accumulator of the compat path
  std::vector<TableEntry> tables;
  bool dirty = false;  // a pass was enqueued but its resolve (which re-arms) was not
  // staging of the host-buffer entry point
  DevBuf s_rgb, s_depth, s_src, s_tgt, s_img, s_dep, s_msk, s_win;
  cudaStream_t hstream = nullptr, h2d_stream = nullptr, d2h_stream = nullptr;
  bool host_pending = false;  // an SE3DS_FLAG_HOST_ASYNC call is in flight on those streams
  std::vector<cudaEvent_t> pipe_ev;  // 2 per batch item: inputs on device, outputs computed
  float margin_scale = 1.0e-6f;  // certification margin: dx = W * scale, dy = 2 * H * scale pixels
  bool pdl = true;  // programmatic dependent launch between the fused kernels
  // concurrent chunk lanes: lane 0 is the caller's stream, lanes 1.. are these (fork / join by events)
  static constexpr int kMaxLanes = 4;
  int lanes = 2;
  long long min_lane_points = 1ll << 20;  // a lane must get at least this many source points per chunk ...
  int min_lane_chunks = 2;                // ... and this many chunks of the call
  cudaStream_t lane_stream[kMaxLanes - 1] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[kMaxLanes - 1] = {};
  int proj_mode = 1;  // 0 canonical only, 1 certified fast path (default), 2 verify
  DevBuf dbg;
  // measurement hooks
  int profile = 0;  // 0 off, 1 cudaEvents between the launches (no PDL), 2 end-of-kernel stamps (pipeline as it runs)
  std::vector<cudaEvent_t> ev_pool;  // groups of 4 events per profiled chunk
  size_t ev_used = 0;
  DevBuf stamps;  // mode 2: kMaxStampChunks x 3 end-of-kernel %globaltimer values
  size_t stamp_used = 0;
  unsigned long long launches = 0;
};

namespace {

size_t ws_total(const se3ds_ws* ws) {
  size_t t = ws->zbuf.cap + ws->zbuf32.cap + ws->fbuf.cap + ws->scf.cap + ws->scr.cap + ws->bins.cap + ws->cbin.cap + ws->cf32.cap + ws->cflag.cap +
             ws->s_rgb.cap + ws->s_depth.cap + ws->s_src.cap + ws->s_tgt.cap + ws->s_img.cap +
             ws->s_dep.cap + ws->s_msk.cap + ws->s_win.cap;
  for (const auto& e : ws->tables) t += (size_t)(4 * e.h + 2 * e.w) * sizeof(float);
  return t;
}

// Job chunking: a chunk's z-buffer + feature buffer + scratch (16 + 8 S bytes per target pixel) should
// sit in L2.  The chunks are dealt round-robin to `lanes` concurrent streams which share that budget; a
// call that cannot give every lane min_lane_chunks chunks of at least min_lane_points source points
// uses fewer lanes (measured: c3 -4.4 %, c4 -8 % with two lanes; c2 would split into one chunk per lane
// and lose 1.3 us to the fork and join).  A chunk is a block of whole batch items with all their
// poses, or -- when one item's poses do not fit -- a block of poses of one item.
struct ChunkPlan {
  int lanes, items_per_chunk, poses_per_chunk;
  long long chunk_jobs, nchunks;
};
void plan_chunks(size_t budget_bytes, int lanes, long long min_lane_points, int min_lane_chunks, int host_group, int n,
                 int s, int p, int h, int w, ChunkPlan* out) {
  const long long hw = (long long)h * w, J = (long long)n * p;
  const size_t job_bytes = (size_t)hw * (16 + 8 * (size_t)s);
  lanes = std::max(1, lanes);
  int items_per_chunk = 1, PC = 1;
  long long chunk_jobs = 1, nchunks_total = 1;
  for (;; --lanes) {
    long long jpc = std::max<long long>(1, (long long)(budget_bytes / lanes / job_bytes));
    jpc = std::min(jpc, (J + lanes - 1) / lanes);
    jpc = std::min<long long>(jpc, std::max(1, 65535 / s));
    jpc = std::min<long long>(jpc, std::max<long long>(1, ((1ll << 32) - 1) / ((long long)s * hw)));  // 32-bit element offsets in the kernels
    if (host_group) jpc = std::min<long long>(jpc, (long long)p * host_group);  // host call: a chunk is one pipeline stage of the copies
    if (jpc >= p) {
      const long long nchunks = (n + (jpc / p) - 1) / (jpc / p);
      items_per_chunk = (int)((n + nchunks - 1) / nchunks);
      PC = p;
    } else {
      const long long pchunks = (p + jpc - 1) / jpc;
      items_per_chunk = 1;
      PC = (int)((p + pchunks - 1) / pchunks);
    }
    chunk_jobs = (long long)items_per_chunk * PC;
    nchunks_total = (long long)((n + items_per_chunk - 1) / items_per_chunk) * ((p + PC - 1) / PC);
    if (lanes == 1 || (nchunks_total >= (long long)lanes * min_lane_chunks && chunk_jobs * s * hw >= min_lane_points)) break;
  }
  *out = ChunkPlan{lanes, items_per_chunk, PC, chunk_jobs, nchunks_total};
}

// Grow-only device buffer; armed buffers are (re)initialised with `pattern` on `stream`.
int grow(DevBuf& b, size_t bytes, int pattern, cudaStream_t stream) {
  if (bytes <= b.cap) return SE3DS_OK;
  if (b.p) CU(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  const size_t want = (bytes + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return fail(SE3DS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  if (pattern >= 0) CU(cudaMemsetAsync(b.p, pattern, want, stream));
  return SE3DS_OK;
}

// Completes a pending asynchronous host call (se3ds_reproject_host with SE3DS_FLAG_HOST_ASYNC): its kernels use
// the workspace's buffers on the workspace's own streams.
int host_wait(se3ds_ws* ws) {
  if (!ws->host_pending) return SE3DS_OK;
  ws->host_pending = false;
  CU(cudaStreamSynchronize(ws->d2h_stream));
  CU(cudaStreamSynchronize(ws->hstream));
  return SE3DS_OK;
}

int rearm(se3ds_ws* ws, cudaStream_t stream) {
  if (ws->zbuf.p) CU(cudaMemsetAsync(ws->zbuf.p, 0xFF, ws->zbuf.cap, stream));
  if (ws->zbuf32.p) CU(cudaMemsetAsync(ws->zbuf32.p, 0xFF, ws->zbuf32.cap, stream));
  if (ws->fbuf.p) CU(cudaMemsetAsync(ws->fbuf.p, 0, ws->fbuf.cap, stream));
  if (ws->bins.p) CU(cudaMemsetAsync(ws->bins.p, 0, ws->bins.cap, stream));
  if (ws->cbin.p) CU(cudaMemsetAsync(ws->cbin.p, 0, ws->cbin.cap, stream));
  if (ws->cf32.p) CU(cudaMemsetAsync(ws->cf32.p, 0, ws->cf32.cap, stream));
  if (ws->cflag.p) CU(cudaMemsetAsync(ws->cflag.p, 0, ws->cflag.cap, stream));
  ws->dirty = false;
  return SE3DS_OK;
}

// tf.linspace in float32 (utils/pano_utils.py:211-219): exact end points, start + delta*i inside.
void linspace_f32(float start, float stop, int n, float* out) {
  if (n == 1) { out[0] = start; return; }
  const float delta = (stop - start) / (float)(n - 1);
  out[0] = start;
  for (int i = 1; i < n - 1; ++i) {
    volatile float step = delta * (float)i;  // two roundings, never an fma
    out[i] = start + step;
  }
  out[n - 1] = stop;
}

// Angle tables of equirectangular_to_pointcloud, canonical sin/cos = float32(double libm).
// Layout: sin_e[H], cos_e[H], sin_h[W], cos_h[W].  H + W values, computed once per shape.
int get_tables(se3ds_ws* ws, int h, int w, cudaStream_t stream, const float** out) {
  for (const auto& e : ws->tables)
    if (e.h == h && e.w == w && e.margin == ws->margin_scale) { *out = e.dev; return SE3DS_OK; }
  const double pi = 3.141592653589793;
  const double hp = 0.5 * pi / (double)h;
  std::vector<float> elev(h), head(w), host((size_t)4 * h + 2 * w);
  linspace_f32((float)hp, (float)(pi - hp), h, elev.data());
  linspace_f32((float)(1.5 * pi - hp), (float)(-0.5 * pi + hp), w, head.data());
  for (int r = 0; r < h; ++r) {
    host[r] = (float)std::sin((double)elev[r]);
    host[h + r] = (float)std::cos((double)elev[r]);
  }
  for (int c = 0; c < w; ++c) {
    host[2 * h + c] = (float)std::sin((double)head[c]);
    host[2 * h + w + c] = (float)std::cos((double)head[c]);
  }
  // row certification table of the fast projection (FastProj::rowb): q = z / rad certainly belongs to row r
  // when it lies strictly between cos((r+1) pi/H) + m and cos(r pi/H) - m; m = the pixel margin dy = 2 H
  // margin_scale as an angle, bounds rounded inwards by one float32 step
  const double m = 2.0 * pi * (double)ws->margin_scale;
  for (int r = 0; r < h; ++r) {
    const float lo = (float)(std::cos((double)(r + 1) * pi / (double)h) + m), hi = (float)(std::cos((double)r * pi / (double)h) - m);
    host[(size_t)2 * h + 2 * w + 2 * r] = std::nextafterf(lo, 2.0f);
    host[(size_t)2 * h + 2 * w + 2 * r + 1] = std::nextafterf(hi, -2.0f);
  }
  float* dev = nullptr;
  CU(cudaMalloc(&dev, host.size() * sizeof(float)));
  CU(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  CU(cudaStreamSynchronize(stream));  // host vector goes out of scope
  if (ws->tables.size() >= 16) {
    cudaFree(ws->tables.front().dev);
    ws->tables.erase(ws->tables.begin());
  }
  ws->tables.push_back({h, w, ws->margin_scale, dev});
  *out = dev;
  return SE3DS_OK;
}

// constants of the certified fast projection: polynomials in pixel units (canon_math.cuh)
void fill_fast_proj(FastProj& fp, const se3ds_ws* ws, int h, int w, const float* tab) {
  const double kx = (double)w / (2.0 * 3.141592653589793), ky = (double)h / 3.141592653589793;
  const double atan_c[9] = CANON_ATAN_COEFFS, acos_c[5] = FAST_ACOS_COEFFS;
  for (int i = 0; i < 9; ++i) fp.ca[i] = (float)(kx * atan_c[i]);
  for (int i = 0; i < 5; ++i) fp.ce[i] = (float)(ky * acos_c[i]);
  fp.kx = (float)kx;
  fp.w4 = (float)(0.25 * w); fp.w34 = (float)(0.75 * w); fp.hf = (float)h;
  fp.dx = (float)w * ws->margin_scale;
  fp.rowb = reinterpret_cast<const float2*>(tab + 2 * (size_t)h + 2 * (size_t)w);
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SE3DS_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return SE3DS_OK;
}

// Launch with programmatic stream serialization (see pdl_enter in kernels.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// K2 grid: a block walks several rows (kernels.cuh); rows_per_block is sized so that the grid is about
// one resident wave of kK2BlocksPerSM blocks per SM, capped so that short panos still spread over the SMs.
constexpr int kK2BlocksPerSM = 8;
constexpr int kHostStages = 4;  // pipeline stages of a blocking se3ds_reproject_host call (measured, see reproject_host)
int k2_row_groups(const se3ds_ws* ws, int gx, int h, int job_frames) {
  const long long row_blocks = (long long)gx * h * job_frames;
  const long long resident = (long long)ws->sm_count * kK2BlocksPerSM;
  long long rpb = (row_blocks + resident - 1) / resident;
  rpb = std::max<long long>(1, std::min<long long>(rpb, 32));
  return (int)((h + rpb - 1) / rpb);
}

template <typename RGB_T, int PPT, bool KEY64>
int run_chunk_t(se3ds_ws* ws, const FusedParams& q, int nitems, cudaStream_t st) {
  const int gx = (q.W + kThreads * PPT - 1) / (kThreads * PPT);
  const int jobs = nitems * q.PC;
  const dim3 grid(gx, q.H, jobs * q.S), block(kThreads);
  const int gx2 = (q.W + kThreads * 4 - 1) / (kThreads * 4);
  const dim3 grid2(gx2, k2_row_groups(ws, gx2, q.H, jobs * q.S), jobs * q.S);
  cudaEvent_t* ev = nullptr;
  FusedParams qs = q;  // + this chunk's stamp slots in profile mode 2
  if (ws->profile == 2 && ws->stamp_used < kMaxStampChunks) qs.stamps = (unsigned long long*)ws->stamps.p + kStampWords * ws->stamp_used++;
  if (ws->profile == 1) {
    if (ws->ev_used + 4 > ws->ev_pool.size())
      for (int i = 0; i < 4; ++i) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ws->ev_pool.push_back(e);
      }
    ev = &ws->ev_pool[ws->ev_used];
    ws->ev_used += 4;
    CU(cudaEventRecord(ev[0], st));
  }
  // FAST feature mode: see splat_depth_kernel
  const bool fast = std::is_same<RGB_T, uint8_t>::value && q.pv == -1 &&
                    (!(q.flags & SE3DS_FLAG_FILTER_VOID) || q.uv == -1);
  const int proj = ws->proj_mode;
  const bool pdl = ws->pdl && ws->profile != 1;
  constexpr bool VEC = PPT == 4;
  // PLAIN: see splat_depth_kernel (no compaction, unproject_void == -1, every thread of the grid inside its row)
  const bool plain = fast && VEC && !(q.flags & SE3DS_FLAG_FILTER_VOID) && q.uv == -1 && q.W % (kThreads * 4) == 0;
#define LAUNCH_K2(F, P, PL)                                                                                        \
  do {                                                                                                            \
    if (q.tgt_rot) CU(launch_pdl(splat_depth_kernel<RGB_T, VEC, F, P, KEY64, true, PL>, grid2, block, st, pdl, qs)); \
    else CU(launch_pdl(splat_depth_kernel<RGB_T, VEC, F, P, KEY64, false, PL>, grid2, block, st, pdl, qs));          \
  } while (0)
  if constexpr (VEC && std::is_same<RGB_T, uint8_t>::value) {
    if (plain) { if (proj == 0) LAUNCH_K2(true, 0, true); else if (proj == 1) LAUNCH_K2(true, 1, true); else LAUNCH_K2(true, 2, true); }
  }
  if (plain) {}
  else if (fast) { if (proj == 0) LAUNCH_K2(true, 0, false); else if (proj == 1) LAUNCH_K2(true, 1, false); else LAUNCH_K2(true, 2, false); }
  else { if (proj == 0) LAUNCH_K2(false, 0, false); else if (proj == 1) LAUNCH_K2(false, 1, false); else LAUNCH_K2(false, 2, false); }
#undef LAUNCH_K2
  if (ev) CU(cudaEventRecord(ev[1], st));
  CU(launch_pdl(splat_feat_kernel<RGB_T, PPT, KEY64>, grid, block, st, pdl, qs));
  if (ev) CU(cudaEventRecord(ev[2], st));
  if (q.flags & SE3DS_FLAG_COMPACT_OUT) CU(launch_pdl(resolve_kernel<PPT, KEY64, true>, dim3(gx, q.H, jobs), block, st, pdl, qs));
  else CU(launch_pdl(resolve_kernel<PPT, KEY64, false>, dim3(gx, q.H, jobs), block, st, pdl, qs));
  if (ev) CU(cudaEventRecord(ev[3], st));
  ws->launches += 3;
  return launch_check("fused reprojection kernels");
}

template <typename RGB_T>
int run_chunk(se3ds_ws* ws, const FusedParams& q, int nitems, bool vec, bool key64, cudaStream_t st) {
  if (key64) return vec ? run_chunk_t<RGB_T, 4, true>(ws, q, nitems, st) : run_chunk_t<RGB_T, 1, true>(ws, q, nitems, st);
  return vec ? run_chunk_t<RGB_T, 4, false>(ws, q, nitems, st) : run_chunk_t<RGB_T, 1, false>(ws, q, nitems, st);
}

}  // namespace

extern "C" {

int se3ds_version(void) { return SE3DS_GEOM_VERSION; }

const char* se3ds_status_string(int status) {
  switch (status) {
    case SE3DS_OK: return "ok";
    case SE3DS_ERR_BAD_SHAPE: return "bad shape";
    case SE3DS_ERR_BAD_DTYPE: return "bad dtype";
    case SE3DS_ERR_BAD_ARG: return "bad argument";
    case SE3DS_ERR_CUDA: return "CUDA error";
    case SE3DS_ERR_NOMEM: return "out of device memory";
    default: return "unknown status";
  }
}

const char* se3ds_last_error(void) { return g_err; }

int se3ds_ws_create(int device, size_t max_bytes, size_t l2_chunk_bytes, se3ds_ws** out) {
  if (!out) return fail(SE3DS_ERR_BAD_ARG, "out is NULL");
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(SE3DS_ERR_BAD_ARG, "device %d of %d", device, count);
  GUARD(device);
  se3ds_ws* ws = new se3ds_ws();
  ws->device = device;
  cudaDeviceGetAttribute(&ws->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (max_bytes) ws->max_bytes = max_bytes;
  if (l2_chunk_bytes) ws->chunk_bytes = l2_chunk_bytes;
  *out = ws;
  // created up front: nothing is allocated while a caller captures a CUDA graph
  CU(cudaEventCreateWithFlags(&ws->fork_ev, cudaEventDisableTiming));
  for (int i = 0; i < se3ds_ws::kMaxLanes - 1; ++i) {
    CU(cudaStreamCreateWithFlags(&ws->lane_stream[i], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ws->join_ev[i], cudaEventDisableTiming));
  }
  return SE3DS_OK;
}

int se3ds_ws_destroy(se3ds_ws* ws) {
  if (!ws) return SE3DS_OK;
  DeviceGuard guard_(ws->device);
  cudaDeviceSynchronize();
  for (DevBuf* b : {&ws->zbuf, &ws->zbuf32, &ws->fbuf, &ws->scf, &ws->scr, &ws->bins, &ws->cbin, &ws->cf32, &ws->cflag, &ws->dbg, &ws->stamps, &ws->s_rgb, &ws->s_depth,
                    &ws->s_src, &ws->s_tgt, &ws->s_img, &ws->s_dep, &ws->s_msk, &ws->s_win})
    if (b->p) cudaFree(b->p);
  for (auto& e : ws->tables) cudaFree(e.dev);
  for (auto& e : ws->ev_pool) cudaEventDestroy(e);
  for (auto& e : ws->pipe_ev) cudaEventDestroy(e);
  for (cudaStream_t st : {ws->hstream, ws->h2d_stream, ws->d2h_stream, ws->lane_stream[0], ws->lane_stream[1], ws->lane_stream[2]})
    if (st) cudaStreamDestroy(st);
  if (ws->fork_ev) cudaEventDestroy(ws->fork_ev);
  for (auto& e : ws->join_ev)
    if (e) cudaEventDestroy(e);
  delete ws;
  return SE3DS_OK;
}

int se3ds_ws_bytes(const se3ds_ws* ws, size_t* bytes) {
  if (!ws || !bytes) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  *bytes = ws_total(ws);
  return SE3DS_OK;
}

int se3ds_ws_projection_mode(se3ds_ws* ws, int mode, float margin_scale) {
  if (!ws || mode < 0 || mode > 2) return fail(SE3DS_ERR_BAD_ARG, "mode must be 0, 1 or 2");
  ws->proj_mode = mode;
  if (margin_scale > 0.0f) ws->margin_scale = margin_scale;
  return SE3DS_OK;
}

int se3ds_ws_pdl(se3ds_ws* ws, int enable) {
  if (!ws) return fail(SE3DS_ERR_BAD_ARG, "NULL workspace");
  ws->pdl = enable != 0;
  return SE3DS_OK;
}

int se3ds_ws_lanes(se3ds_ws* ws, int lanes, long long min_points_per_lane, int min_chunks_per_lane) {
  if (!ws) return fail(SE3DS_ERR_BAD_ARG, "NULL workspace");
  if (lanes < 1 || lanes > se3ds_ws::kMaxLanes) return fail(SE3DS_ERR_BAD_ARG, "lanes must be in [1, %d]", se3ds_ws::kMaxLanes);
  if (min_points_per_lane < 0 || min_chunks_per_lane < 0) return fail(SE3DS_ERR_BAD_ARG, "lane minima must be >= 0");
  ws->lanes = lanes;
  ws->min_lane_points = min_points_per_lane ? min_points_per_lane : (1ll << 20);
  ws->min_lane_chunks = min_chunks_per_lane ? min_chunks_per_lane : 2;
  return SE3DS_OK;
}

int se3ds_plan_chunks(size_t l2_chunk_bytes, int lanes, long long min_points_per_lane, int min_chunks_per_lane, int n,
                      int s, int p, int h, int w, long long plan[5]) {
  if (!plan) return fail(SE3DS_ERR_BAD_ARG, "plan is NULL");
  if (n <= 0 || s <= 0 || p <= 0 || h <= 0 || w != 2 * h) return fail(SE3DS_ERR_BAD_SHAPE, "bad shape");
  if (lanes < 1 || lanes > se3ds_ws::kMaxLanes) return fail(SE3DS_ERR_BAD_ARG, "lanes must be in [1, %d]", se3ds_ws::kMaxLanes);
  ChunkPlan c;
  plan_chunks(l2_chunk_bytes ? l2_chunk_bytes : kDefaultChunkBytes, lanes, min_points_per_lane ? min_points_per_lane : (1ll << 20),
              min_chunks_per_lane ? min_chunks_per_lane : 2, 0, n, s, p, h, w, &c);
  plan[0] = c.lanes; plan[1] = c.items_per_chunk; plan[2] = c.poses_per_chunk; plan[3] = c.chunk_jobs; plan[4] = c.nchunks;
  return SE3DS_OK;
}

int se3ds_ws_verify_read(se3ds_ws* ws, unsigned long long counts[3], float max_dev[2]) {
  if (!ws || !counts || !max_dev) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  counts[0] = counts[1] = counts[2] = 0;
  max_dev[0] = max_dev[1] = 0.f;
  if (!ws->dbg.p) return SE3DS_OK;
  GUARD(ws->device);
  CU(cudaDeviceSynchronize());
  unsigned long long h[4];
  CU(cudaMemcpy(h, ws->dbg.p, sizeof(h), cudaMemcpyDeviceToHost));
  CU(cudaMemset(ws->dbg.p, 0, sizeof(h)));
  counts[0] = h[0]; counts[1] = h[1]; counts[2] = h[2];
  const unsigned hi = (unsigned)(h[3] >> 32), lo = (unsigned)h[3];
  memcpy(&max_dev[0], &hi, 4);
  memcpy(&max_dev[1], &lo, 4);
  return SE3DS_OK;
}

int se3ds_ws_profile(se3ds_ws* ws, int mode) {
  if (!ws || mode < 0 || mode > 2) return fail(SE3DS_ERR_BAD_ARG, "mode must be 0, 1 or 2");
  ws->profile = mode;
  ws->ev_used = 0;
  ws->stamp_used = 0;
  if (mode == 2) {
    GUARD(ws->device);
    CU(cudaDeviceSynchronize());
    if (int rc = grow(ws->stamps, kMaxStampChunks * kStampWords * sizeof(unsigned long long), -1, nullptr)) return rc;
    CU(cudaMemset(ws->stamps.p, 0, kMaxStampChunks * kStampWords * sizeof(unsigned long long)));
  }
  return SE3DS_OK;
}

int se3ds_ws_profile_read_stamps(se3ds_ws* ws, double ms[3], long long* chunks) {
  if (!ws || !ms || !chunks) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  ms[0] = ms[1] = ms[2] = 0.0;
  *chunks = 0;
  if (!ws->stamps.p || ws->stamp_used < 2) return SE3DS_OK;
  GUARD(ws->device);
  CU(cudaDeviceSynchronize());
  std::vector<unsigned long long> raw(kStampWords * ws->stamp_used), h(3 * ws->stamp_used);
  CU(cudaMemcpy(raw.data(), ws->stamps.p, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CU(cudaMemset(ws->stamps.p, 0, kMaxStampChunks * kStampWords * sizeof(unsigned long long)));
  for (size_t i = 0; i < h.size(); ++i)  // end of a kernel = the latest of its blocks' stamps
    h[i] = *std::max_element(raw.begin() + i * kStampSlots, raw.begin() + (i + 1) * kStampSlots);
  // chunk i: K2 ends at h[3i], K3 at h[3i+1], K4 at h[3i+2]; a kernel's share = its end - the previous end.
  // The first chunk has no predecessor: it is skipped.
  for (size_t i = 1; i < ws->stamp_used; ++i) {
    const unsigned long long prev = h[3 * i - 1];
    if (!prev || !h[3 * i] || !h[3 * i + 1] || !h[3 * i + 2]) continue;
    ms[0] += (double)((long long)(h[3 * i] - prev)) * 1e-6;
    ms[1] += (double)((long long)(h[3 * i + 1] - h[3 * i])) * 1e-6;
    ms[2] += (double)((long long)(h[3 * i + 2] - h[3 * i + 1])) * 1e-6;
    ++*chunks;
  }
  ws->stamp_used = 0;
  return SE3DS_OK;
}

int se3ds_ws_profile_read(se3ds_ws* ws, float ms[3], unsigned long long* launches) {
  if (!ws || !ms) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(ws->device);
  ms[0] = ms[1] = ms[2] = 0.f;
  for (size_t i = 0; i + 4 <= ws->ev_used; i += 4) {
    CU(cudaEventSynchronize(ws->ev_pool[i + 3]));
    for (int k = 0; k < 3; ++k) {
      float t = 0.f;
      CU(cudaEventElapsedTime(&t, ws->ev_pool[i + k], ws->ev_pool[i + k + 1]));
      ms[k] += t;
    }
  }
  ws->ev_used = 0;
  if (launches) *launches = ws->launches;
  return SE3DS_OK;
}

int se3ds_mask_pano(const void* pano, int dtype, int n, int h, int w, int c, double proportion,
                    double masked_region_value, void* out, void* stream) {
  if (!pano || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL tensor");
  GUARD(device_of(out));
  if (n < 0 || h <= 0 || w <= 0 || c <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "pano must be (N,H,W,C)");
  const long long total = (long long)n * h * w * c;
  if (total == 0) return SE3DS_OK;
  const int mh = (int)(h * proportion);
  const long long row_elems = (long long)w * c;
  const int blocks = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 16);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SE3DS_U8:
      mask_pano_kernel<uint8_t><<<blocks, kThreads, 0, st>>>((const uint8_t*)pano, (uint8_t*)out, total, h, row_elems, mh, (uint8_t)masked_region_value);
      break;
    case SE3DS_I32:
      mask_pano_kernel<int><<<blocks, kThreads, 0, st>>>((const int*)pano, (int*)out, total, h, row_elems, mh, (int)masked_region_value);
      break;
    case SE3DS_F32:
      mask_pano_kernel<float><<<blocks, kThreads, 0, st>>>((const float*)pano, (float*)out, total, h, row_elems, mh, (float)masked_region_value);
      break;
    default: return fail(SE3DS_ERR_BAD_DTYPE, "dtype %d", dtype);
  }
  return launch_check("mask_pano_kernel");
}

int se3ds_unproject_equirect(se3ds_ws* ws, const void* feats, int in_dtype, const float* depth, int n,
                             int h, int w, int c, double void_class, float depth_scale,
                             float* xyz1_out, void* feats_out, int out_dtype, void* stream) {
  if (!ws || !feats || !depth || !xyz1_out || !feats_out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  if (n < 0 || h <= 0 || w <= 0 || c <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "feats should have shape (N, H, W) or (N, H, W, C)");
  // W == 2H is asserted by the caller on the *input* image (pano_utils.py:202); after a size_mult
  // resize the scaled width can be odd, and the angle tables only need (h, w).
  if (void_class < 0.0 && in_dtype == SE3DS_U8)
    return fail(SE3DS_ERR_BAD_DTYPE, "feats datatype must be signed if the void class is negative");
  if (out_dtype != in_dtype && out_dtype != SE3DS_F32) return fail(SE3DS_ERR_BAD_DTYPE, "out dtype must be the input dtype or f32");
  if (n == 0) return SE3DS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GUARD(ws->device);
  const float* tab = nullptr;
  if (int rc = get_tables(ws, h, w, st, &tab)) return rc;
  const long long total = (long long)n * h * w;
  const int blocks = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 16);
#define UNPROJ(TI, TO)                                                                              \
  unproject_kernel<TI, TO><<<blocks, kThreads, 0, st>>>((const TI*)feats, depth, tab, n, h, w, c, \
                                                        depth_scale, (TO)void_class, xyz1_out, (TO*)feats_out)
  if (in_dtype == SE3DS_U8 && out_dtype == SE3DS_U8) UNPROJ(uint8_t, uint8_t);
  else if (in_dtype == SE3DS_U8 && out_dtype == SE3DS_F32) UNPROJ(uint8_t, float);
  else if (in_dtype == SE3DS_I32 && out_dtype == SE3DS_I32) UNPROJ(int, int);
  else if (in_dtype == SE3DS_I32 && out_dtype == SE3DS_F32) UNPROJ(int, float);
  else if (in_dtype == SE3DS_F32 && out_dtype == SE3DS_F32) UNPROJ(float, float);
  else return fail(SE3DS_ERR_BAD_DTYPE, "dtype combination %d -> %d", in_dtype, out_dtype);
#undef UNPROJ
  return launch_check("unproject_kernel");
}

int se3ds_project_cloud(se3ds_ws* ws, const float* coords, const void* feats, int feat_dtype, int n,
                        long long m, int c, int h, int w, int mode, float input_void_class,
                        float output_void_class, float depth_scale, float* depth_out,
                        float* feats_out, int32_t* winner_out, void* stream) {
  if (!ws || !depth_out || !feats_out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  if (n < 0 || m < 0 || c <= 0 || h <= 0 || w <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "feats should have shape (N, M) or (N, M, C)");
  if (m > 0 && (!coords || !feats)) return fail(SE3DS_ERR_BAD_ARG, "NULL cloud");
  if (m >= (1ll << 31) || (long long)h * w > (long long)kScPixMask) return fail(SE3DS_ERR_BAD_SHAPE, "cloud or image too large");
  if (mode != 0 && mode != 1) return fail(SE3DS_ERR_BAD_ARG, "mode %d", mode);
  if (feat_dtype < SE3DS_U8 || feat_dtype > SE3DS_F32) return fail(SE3DS_ERR_BAD_DTYPE, "dtype %d", feat_dtype);
  if (n == 0) return SE3DS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GUARD(ws->device);
  if (int rc = host_wait(ws)) return rc;
  const long long npix = (long long)n * h * w;
  if (ws->dirty)
    if (int rc = rearm(ws, st)) return rc;
  const bool key64 = winner_out != nullptr;
  // F16 mode: three channels of integer features with output void class 0 (every RGB caller of the reference)
  const bool f16 = c == 3 && feat_dtype != SE3DS_F32 && output_void_class == 0.0f;
  if (key64) { if (int rc = grow(ws->zbuf, (size_t)npix * 8, 0xFF, st)) return rc; }
  else { if (int rc = grow(ws->zbuf32, (size_t)npix * 4, 0xFF, st)) return rc; }
  if (f16) {
    if (int rc = grow(ws->fbuf, (size_t)npix * 8, 0, st)) return rc;
    if (int rc = grow(ws->cf32, (size_t)npix * 12, 0, st)) return rc;
    if (int rc = grow(ws->cflag, 256, 0, st)) return rc;
  }
  if (int rc = grow(ws->scf, (size_t)std::max<long long>(n * m, 1) * 4, -1, st)) return rc;
  if (int rc = grow(ws->scr, (size_t)std::max<long long>(n * m, 1) * 4, -1, st)) return rc;
  if (int rc = grow(ws->cbin, (size_t)(1 + c) * 4, 0, st)) return rc;
  CloudParams q{};
  q.coords = coords; q.feats = feats;
  q.zbuf = (unsigned long long*)ws->zbuf.p; q.zbuf32 = (uint32_t*)ws->zbuf32.p;
  q.fbuf = (uint2*)ws->fbuf.p; q.fbuf32 = (float*)ws->cf32.p; q.used32 = (uint32_t*)ws->cflag.p;
  q.sc_flat = (uint32_t*)ws->scf.p; q.sc_rad = (float*)ws->scr.p;
  q.bin = (uint32_t*)ws->cbin.p;
  q.depth_out = depth_out; q.feats_out = feats_out; q.winner_out = winner_out;
  q.M = m; q.N = n; q.C = c; q.H = h; q.W = w; q.HW = h * w; q.mode = mode;
  q.void_in = input_void_class; q.void_out = output_void_class; q.depth_scale = depth_scale;
  if (mode == 0) {
    const float* tab = nullptr;
    if (int rc = get_tables(ws, h, w, st, &tab)) return rc;
    fill_fast_proj(q.fast, ws, h, w, tab);
  }
  ws->dirty = true;
  const bool pdl = ws->pdl;
  const dim3 block(kThreads);
  if (!f16) {  // the float32 atomics accumulate in the caller's tensor: it starts as the output void class
    fill_f32_kernel<<<(int)std::min<long long>((npix * c + kThreads - 1) / kThreads, 148 * 16), kThreads, 0, st>>>(feats_out, npix * c, output_void_class);
    ws->launches += 1;
  }
  if (m > 0) {
    // persistent grid-stride kernels: about one resident wave in x, the batch in y
    const long long want = (m + kThreads - 1) / kThreads, cap = std::max<long long>(1, (long long)ws->sm_count * 16 / std::max(n, 1));
    const dim3 grid((unsigned)std::min(want, cap), n);
#define CLOUD_SPLAT(T)                                                                                       \
  do {                                                                                                      \
    if (key64) {                                                                                            \
      CU(launch_pdl(cloud_depth_kernel<T, true>, grid, block, st, pdl, q));                                 \
      if (f16) CU(launch_pdl(cloud_feat_kernel<T, true, true>, grid, block, st, pdl, q));                   \
      else CU(launch_pdl(cloud_feat_kernel<T, true, false>, grid, block, st, pdl, q));                      \
    } else {                                                                                                \
      CU(launch_pdl(cloud_depth_kernel<T, false>, grid, block, st, pdl, q));                                \
      if (f16) CU(launch_pdl(cloud_feat_kernel<T, false, true>, grid, block, st, pdl, q));                  \
      else CU(launch_pdl(cloud_feat_kernel<T, false, false>, grid, block, st, pdl, q));                     \
    }                                                                                                       \
  } while (0)
    switch (feat_dtype) {
      case SE3DS_U8: CLOUD_SPLAT(uint8_t); break;
      case SE3DS_I32: CLOUD_SPLAT(int); break;
      default: CLOUD_SPLAT(float); break;
    }
#undef CLOUD_SPLAT
    ws->launches += 2;
  }
  const dim3 rgrid((unsigned)((npix + kThreads - 1) / kThreads));
  if (key64) {
    if (f16) CU(launch_pdl(cloud_resolve_kernel<true, true>, rgrid, block, st, pdl, q));
    else CU(launch_pdl(cloud_resolve_kernel<true, false>, rgrid, block, st, pdl, q));
  } else {
    if (f16) CU(launch_pdl(cloud_resolve_kernel<false, true>, rgrid, block, st, pdl, q));
    else CU(launch_pdl(cloud_resolve_kernel<false, false>, rgrid, block, st, pdl, q));
  }
  ws->launches += 1;
  if (f16) {
    CU(launch_pdl(cloud_clear_flag_kernel, dim3(1), dim3(32), st, pdl, q.used32));
    ws->launches += 1;
  }
  if (int rc = launch_check("cloud projection kernels")) return rc;
  ws->dirty = false;
  return SE3DS_OK;
}

}  // extern "C"

namespace {

// Host pipeline of se3ds_reproject_host: chunk c (one batch item) waits for its inputs on the
// compute stream and hands its outputs to the device->host stream as soon as its resolve is done.
struct HostPipe {
  int group;  // batch items per pipeline stage (upload -> kernels -> download)
  cudaStream_t d2h;
  cudaEvent_t* in_ready;   // [n]
  cudaEvent_t* out_ready;  // [n]
  float *image_host, *depth_host, *mask_host;
  int32_t* winner_host;
};

int reproject_core(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                   const float* src_pos, const float* tgt_pos, const float* tgt_rot, int n, int s, int s_capacity, int p, int h, int w,
                   float depth_scale, double mask_proportion, int mask_frames, int unproject_void,
                   int project_void, unsigned flags, float* proj_image, float* proj_depth,
                   float* proj_mask, int32_t* winner_out, float* bin_out, void* stream, const HostPipe* pipe) {
  const bool compact = flags & SE3DS_FLAG_COMPACT_OUT;
  if (!ws || !rgb || !depth || !src_pos || !tgt_pos || !proj_image || !proj_depth || (!proj_mask && !compact))
    return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  if (compact && (flags & SE3DS_FLAG_RAW_FEATURES)) return fail(SE3DS_ERR_BAD_ARG, "COMPACT_OUT and RAW_FEATURES exclude each other");
  if (n < 0 || s <= 0 || p <= 0 || h <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "rgb must be (N,S,H,W,3), tgt_pos (N,P,3)");
  if (s_capacity == 0) s_capacity = s;
  if (s_capacity < s) return fail(SE3DS_ERR_BAD_SHAPE, "frame capacity %d < frames %d", s_capacity, s);
  if (w != 2 * h) return fail(SE3DS_ERR_BAD_SHAPE, "Expected equirectangular input images");
  if (rgb_dtype != SE3DS_U8 && rgb_dtype != SE3DS_I32) return fail(SE3DS_ERR_BAD_DTYPE, "rgb must be uint8 or int32");
  if (unproject_void < -1 || unproject_void > 255 || project_void < -1 || project_void > 255)
    return fail(SE3DS_ERR_BAD_ARG, "void classes must be in [-1, 255]");
  const long long hw = (long long)h * w;
  if (hw > (long long)kScPixMask || (long long)s * hw >= (1ll << 31)) return fail(SE3DS_ERR_BAD_SHAPE, "S*H*W too large");
  if ((long long)n * (s_capacity ? s_capacity : s) >= (1ll << 31) / 3) return fail(SE3DS_ERR_BAD_SHAPE, "N*S too large");
  if (s > 32767 || h > 65535) return fail(SE3DS_ERR_BAD_SHAPE, "S or H too large");
  if (n == 0) return SE3DS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GUARD(ws->device);
  if (!pipe)
    if (int rc = host_wait(ws)) return rc;
  if (pipe && !pipe->in_ready) pipe = nullptr;  // host call that moves the whole batch itself: nothing per item here

  // job chunking and lanes (plan_chunks below)
  const long long J = (long long)n * p;
  ChunkPlan plan;
  plan_chunks(std::min(ws->chunk_bytes, ws->max_bytes), (pipe || ws->profile == 1) ? 1 : ws->lanes, ws->min_lane_points,
              ws->min_lane_chunks, pipe ? pipe->group : 0, n, s, p, h, w, &plan);
  const int lanes = plan.lanes, items_per_chunk = plan.items_per_chunk, PC = plan.poses_per_chunk;
  const long long chunk_jobs = plan.chunk_jobs, nchunks_total = plan.nchunks;
  const bool per_job = flags & SE3DS_FLAG_BIN_PER_JOB;
  if (bin_out && per_job) return fail(SE3DS_ERR_BAD_ARG, "bin_out needs the per-call bin mode");

  if (ws->dirty)
    if (int rc = rearm(ws, st)) return rc;
  // 64-bit packed (depth | index) keys when winner indices are wanted (or forced), else depth-only keys
  const bool key64 = winner_out != nullptr || (flags & SE3DS_FLAG_KEY64);
  const size_t lane_px = (size_t)chunk_jobs * hw;  // every lane owns one chunk's slice of each buffer
  if (key64) { if (int rc = grow(ws->zbuf, lanes * lane_px * 8, 0xFF, st)) return rc; }
  else { if (int rc = grow(ws->zbuf32, lanes * lane_px * 4, 0xFF, st)) return rc; }
  if (int rc = grow(ws->fbuf, lanes * lane_px * 8, 0, st)) return rc;
  if (int rc = grow(ws->scf, lanes * lane_px * s * 4, -1, st)) return rc;
  if (int rc = grow(ws->scr, lanes * lane_px * s * 4, -1, st)) return rc;
  if (int rc = grow(ws->bins, (size_t)(per_job ? J : 1) * kBinReplicas * sizeof(Bin), 0, st)) return rc;
  const float* tab = nullptr;
  if (int rc = get_tables(ws, h, w, st, &tab)) return rc;

  FusedParams q{};
  q.rgb = rgb; q.depth = depth; q.src_pos = src_pos; q.tgt_pos = tgt_pos; q.tgt_rot = tgt_rot; q.tab = tab;
  q.zbuf = (unsigned long long*)ws->zbuf.p; q.zbuf32 = (uint32_t*)ws->zbuf32.p; q.fbuf = (uint2*)ws->fbuf.p;
  q.sc_flat = (uint32_t*)ws->scf.p; q.sc_rad = (float*)ws->scr.p; q.bins = (Bin*)ws->bins.p;
  q.out_image = proj_image; q.out_depth = proj_depth; q.out_mask = proj_mask; q.out_winner = winner_out;
  q.N = n; q.S = s; q.SC = s_capacity; q.P = p; q.H = h; q.W = w; q.HW = (int)hw;
  q.mh = (int)(h * mask_proportion);
  q.mask_frames = mask_frames;
  q.uv = unproject_void; q.pv = project_void; q.flags = flags; q.depth_scale = depth_scale;
  q.finalize_bins = (per_job || nchunks_total == 1) ? 1 : 0;
  q.bin_out = bin_out;
  // measured: c3 (4 frames, 17 MB feature buffer per chunk) K3 -25 %, K2 -6 %; c5 (8 frames, 67 MB
  // feature buffer, 34 MB z-buffer per chunk) K3 +19 % but K2 -12 %
  q.prefilter_z = (s > 1 && lanes * lane_px * (key64 ? 8 : 4) <= ((size_t)64 << 20)) ? 1 : 0;
  q.prefilter_f = (s > 1 && lanes * lane_px * 8 <= ((size_t)48 << 20)) ? 1 : 0;
  if (int rc = grow(ws->dbg, 4 * sizeof(unsigned long long), 0, st)) return rc;
  q.dbg = (unsigned long long*)ws->dbg.p;
  fill_fast_proj(q.fast, ws, h, w, tab);
  {
    volatile float one = 1.0f, ds = depth_scale;
    q.inv_depth_scale = one / ds;  // IEEE single division on the host: RN(1 / depth_scale)
  }
  // the 4-pixels-per-thread kernels use 128 / 256-bit accesses on the inputs AND on every output plane
  const bool vec = (w % 4 == 0) && aligned(depth, 16) && aligned(rgb, rgb_dtype == SE3DS_U8 ? 4 : 16) &&
                   aligned(proj_image, compact ? 4 : 16) && aligned(proj_depth, 16) && (compact || aligned(proj_mask, 16)) &&
                   (!winner_out || aligned(winner_out, 16));

  ws->dirty = true;
  // fork: everything enqueued so far on the caller's stream (its inputs, re-arming, tables) comes first
  cudaStream_t lane_st[se3ds_ws::kMaxLanes] = {st, ws->lane_stream[0], ws->lane_stream[1], ws->lane_stream[2]};
  if (lanes > 1) {
    CU(cudaEventRecord(ws->fork_ev, st));
    for (int l = 1; l < lanes; ++l) CU(cudaStreamWaitEvent(lane_st[l], ws->fork_ev, 0));
  }
  auto join = [&]() -> cudaError_t {
    for (int l = 1; l < lanes; ++l) {
      if (cudaError_t e = cudaEventRecord(ws->join_ev[l - 1], lane_st[l])) return e;
      if (cudaError_t e = cudaStreamWaitEvent(st, ws->join_ev[l - 1], 0)) return e;
    }
    return cudaSuccess;
  };
  long long chunk_no = 0;
  for (int n0 = 0; n0 < n; n0 += items_per_chunk) {
    const int nitems = std::min(items_per_chunk, n - n0);
    if (pipe) CU(cudaStreamWaitEvent(st, pipe->in_ready[n0], 0));  // recorded after the upload of items n0 .. n0 + nitems - 1
    for (int p0 = 0; p0 < p; p0 += PC, ++chunk_no) {
      const int lane = (int)(chunk_no % lanes);
      q.n0 = n0; q.p0 = p0; q.PC = std::min(PC, p - p0);
      q.zbuf = (unsigned long long*)ws->zbuf.p + (key64 ? lane * lane_px : 0);
      q.zbuf32 = (uint32_t*)ws->zbuf32.p + (key64 ? 0 : lane * lane_px);
      q.fbuf = (uint2*)ws->fbuf.p + lane * lane_px;
      q.sc_flat = (uint32_t*)ws->scf.p + lane * lane_px * s;
      q.sc_rad = (float*)ws->scr.p + lane * lane_px * s;
      const int rc = rgb_dtype == SE3DS_U8 ? run_chunk<uint8_t>(ws, q, nitems, vec, key64, lane_st[lane])
                                           : run_chunk<int>(ws, q, nitems, vec, key64, lane_st[lane]);
      if (rc) { join(); return rc; }
    }
    if (pipe) {  // ship the guidance tensors of this stage's items
      CU(cudaEventRecord(pipe->out_ready[n0], st));
      CU(cudaStreamWaitEvent(pipe->d2h, pipe->out_ready[n0], 0));
      const size_t o = (size_t)n0 * p * hw, cnt = (size_t)nitems * p * hw;
      if (compact) {
        CU(cudaMemcpyAsync((uint8_t*)pipe->image_host + o * 3, (const uint8_t*)proj_image + o * 3, cnt * 3, cudaMemcpyDeviceToHost, pipe->d2h));
      } else {
        CU(cudaMemcpyAsync(pipe->image_host + o * 3, proj_image + o * 3, cnt * 12, cudaMemcpyDeviceToHost, pipe->d2h));
        CU(cudaMemcpyAsync(pipe->mask_host + o, proj_mask + o, cnt * 4, cudaMemcpyDeviceToHost, pipe->d2h));
      }
      CU(cudaMemcpyAsync(pipe->depth_host + o, proj_depth + o, cnt * 4, cudaMemcpyDeviceToHost, pipe->d2h));
      if (pipe->winner_host)
        CU(cudaMemcpyAsync(pipe->winner_host + o, winner_out + o, cnt * 4, cudaMemcpyDeviceToHost, pipe->d2h));
    }
  }
  CU(join());
  if (bin_out) {
    export_bin_kernel<<<1, 32, 0, st>>>(q);
    ws->launches += 1;
    if (int rc = launch_check("export_bin_kernel")) return rc;
  } else if (!q.finalize_bins) {
    patch_owner_kernel<<<1, 32, 0, st>>>(q);
    ws->launches += 1;
    if (int rc = launch_check("patch_owner_kernel")) return rc;
    if (pipe) {  // job 0's pixel (0,0) changed after it was shipped: send its 20 bytes again
      CU(cudaEventRecord(pipe->out_ready[0], st));
      CU(cudaStreamWaitEvent(pipe->d2h, pipe->out_ready[0], 0));
      CU(cudaMemcpyAsync(pipe->image_host, proj_image, compact ? 3 : 12, cudaMemcpyDeviceToHost, pipe->d2h));
      CU(cudaMemcpyAsync(pipe->depth_host, proj_depth, 4, cudaMemcpyDeviceToHost, pipe->d2h));
      if (!compact) CU(cudaMemcpyAsync(pipe->mask_host, proj_mask, 4, cudaMemcpyDeviceToHost, pipe->d2h));
      if (pipe->winner_host) CU(cudaMemcpyAsync(pipe->winner_host, winner_out, 4, cudaMemcpyDeviceToHost, pipe->d2h));
    }
  }
  ws->dirty = false;
  return SE3DS_OK;
}

}  // namespace

extern "C" {

int se3ds_reproject(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                    const float* src_pos, const float* tgt_pos, int n, int s, int p, int h, int w,
                    float depth_scale, double mask_proportion, int mask_frames, int unproject_void,
                    int project_void, unsigned flags, float* proj_image, float* proj_depth,
                    float* proj_mask, int32_t* winner_out, float* bin_out, void* stream) {
  return reproject_core(ws, rgb, rgb_dtype, depth, src_pos, tgt_pos, nullptr, n, s, s, p, h, w, depth_scale,
                        mask_proportion, mask_frames, unproject_void, project_void, flags, proj_image, proj_depth,
                        proj_mask, winner_out, bin_out, stream, nullptr);
}

int se3ds_reproject_se3(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                        const float* src_pos, const float* tgt_pos, const float* tgt_rot, int n, int s,
                        int p, int h, int w, float depth_scale, double mask_proportion, int mask_frames,
                        int unproject_void, int project_void, unsigned flags, float* proj_image,
                        float* proj_depth, float* proj_mask, int32_t* winner_out, float* bin_out,
                        void* stream) {
  return reproject_core(ws, rgb, rgb_dtype, depth, src_pos, tgt_pos, tgt_rot, n, s, s, p, h, w, depth_scale,
                        mask_proportion, mask_frames, unproject_void, project_void, flags, proj_image, proj_depth,
                        proj_mask, winner_out, bin_out, stream, nullptr);
}

int se3ds_reproject_ring(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                         const float* src_pos, const float* tgt_pos, const float* tgt_rot, int n, int s,
                         int s_capacity, int p, int h, int w, float depth_scale, double mask_proportion,
                         int mask_frames, int unproject_void, int project_void, unsigned flags,
                         float* proj_image, float* proj_depth, float* proj_mask, int32_t* winner_out,
                         float* bin_out, void* stream) {
  return reproject_core(ws, rgb, rgb_dtype, depth, src_pos, tgt_pos, tgt_rot, n, s, s_capacity, p, h, w, depth_scale,
                        mask_proportion, mask_frames, unproject_void, project_void, flags, proj_image, proj_depth,
                        proj_mask, winner_out, bin_out, stream, nullptr);
}

int se3ds_reproject_host(se3ds_ws* ws, const void* rgb_host, int rgb_dtype, const float* depth_host,
                         const float* src_pos_host, const float* tgt_pos_host, int n, int s, int p,
                         int h, int w, float depth_scale, double mask_proportion, int mask_frames,
                         int unproject_void, int project_void, unsigned flags,
                         float* proj_image_host, float* proj_depth_host, float* proj_mask_host,
                         int32_t* winner_out_host) {
  const bool compact = flags & SE3DS_FLAG_COMPACT_OUT;
  if (!ws || !rgb_host || !depth_host || !src_pos_host || !tgt_pos_host || !proj_image_host ||
      !proj_depth_host || (!proj_mask_host && !compact))
    return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  if (n <= 0 || s <= 0 || p <= 0 || h <= 0 || w != 2 * h) return fail(SE3DS_ERR_BAD_SHAPE, "bad shape");
  if (rgb_dtype != SE3DS_U8 && rgb_dtype != SE3DS_I32) return fail(SE3DS_ERR_BAD_DTYPE, "rgb must be uint8 or int32");
  GUARD(ws->device);
  if (int rc = host_wait(ws)) return rc;  // the staging buffers and streams of a pending call are still in use
  for (cudaStream_t* sp : {&ws->hstream, &ws->h2d_stream, &ws->d2h_stream})
    if (!*sp) CU(cudaStreamCreateWithFlags(sp, cudaStreamNonBlocking));
  while (ws->pipe_ev.size() < (size_t)2 * n) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ws->pipe_ev.push_back(e);
  }
  cudaStream_t st = ws->hstream, up = ws->h2d_stream;
  const size_t hw = (size_t)h * w, npts = (size_t)n * s * hw, npix = (size_t)n * p * hw;
  const size_t px = rgb_dtype == SE3DS_U8 ? 3 : 12;  // rgb bytes per point
  if (int rc = grow(ws->s_rgb, npts * px, -1, st)) return rc;
  if (int rc = grow(ws->s_depth, npts * 4, -1, st)) return rc;
  if (int rc = grow(ws->s_src, (size_t)n * s * 12, -1, st)) return rc;
  if (int rc = grow(ws->s_tgt, (size_t)n * p * 12, -1, st)) return rc;
  if (int rc = grow(ws->s_img, npix * (compact ? 3 : 12), -1, st)) return rc;
  if (int rc = grow(ws->s_dep, npix * 4, -1, st)) return rc;
  if (!compact)
    if (int rc = grow(ws->s_msk, npix * 4, -1, st)) return rc;
  if (winner_out_host)
    if (int rc = grow(ws->s_win, npix * 4, -1, st)) return rc;
  CU(cudaMemcpyAsync(ws->s_src.p, src_pos_host, (size_t)n * s * 12, cudaMemcpyHostToDevice, up));
  CU(cudaMemcpyAsync(ws->s_tgt.p, tgt_pos_host, (size_t)n * p * 12, cudaMemcpyHostToDevice, up));
  if (flags & SE3DS_FLAG_HOST_ASYNC) {
    // Asynchronous call: the overlap comes from the caller's other workspace (its batch travels while this one
    // computes / travels back), so the whole batch moves in as few DMA transfers as possible -- measured on
    // B200 / PCIe 5: 16 + 16 interleaved item-sized copies cost 0.76 ms where one copy each way costs 0.59 ms.
    CU(cudaMemcpyAsync(ws->s_rgb.p, rgb_host, npts * px, cudaMemcpyHostToDevice, up));
    CU(cudaMemcpyAsync(ws->s_depth.p, depth_host, npts * 4, cudaMemcpyHostToDevice, up));
    CU(cudaEventRecord(ws->pipe_ev[0], up));
    CU(cudaStreamWaitEvent(st, ws->pipe_ev[0], 0));
    const HostPipe whole_batch{};  // no per-item events: reproject_core runs the batch as one device call
    const int rc = reproject_core(ws, ws->s_rgb.p, rgb_dtype, (const float*)ws->s_depth.p, (const float*)ws->s_src.p,
                                  (const float*)ws->s_tgt.p, nullptr, n, s, s, p, h, w, depth_scale, mask_proportion, mask_frames,
                                  unproject_void, project_void, flags, (float*)ws->s_img.p, (float*)ws->s_dep.p,
                                  (float*)ws->s_msk.p, winner_out_host ? (int32_t*)ws->s_win.p : nullptr, nullptr, st, &whole_batch);
    ws->host_pending = true;  // the caller's buffers are in use until host_wait
    if (rc) {
      host_wait(ws);
      return rc;
    }
    cudaStream_t down = ws->d2h_stream;
    CU(cudaEventRecord(ws->pipe_ev[1], st));
    CU(cudaStreamWaitEvent(down, ws->pipe_ev[1], 0));
    CU(cudaMemcpyAsync(proj_image_host, ws->s_img.p, npix * (compact ? 3 : 12), cudaMemcpyDeviceToHost, down));
    CU(cudaMemcpyAsync(proj_depth_host, ws->s_dep.p, npix * 4, cudaMemcpyDeviceToHost, down));
    if (!compact) CU(cudaMemcpyAsync(proj_mask_host, ws->s_msk.p, npix * 4, cudaMemcpyDeviceToHost, down));
    if (winner_out_host) CU(cudaMemcpyAsync(winner_out_host, ws->s_win.p, npix * 4, cudaMemcpyDeviceToHost, down));
    return SE3DS_OK;
  }
  // Blocking call: a pipeline over groups of batch items -- group g computes while group g+1 is still on the wire
  // and group g-1 is already travelling back.  Few stages: every DMA transfer costs ~5 us while both directions
  // of the link are busy, and the uploads of later groups share the link with the downloads of earlier ones.
  // Measured (512x1024 panos, compact / float32 outputs, ms per call): batch 8: 1 stage 1.16 / 2.13, 2: 0.98 / 1.90,
  // 4: 0.95 / 1.83, 8: 1.02 / 1.87; batch 16: 1: 1.84 / 3.70, 4: 1.70 / 3.47, 8: 1.74 / 3.43, 16: 2.03 / 3.61.
  const int group = (n + kHostStages - 1) / kHostStages;
  ChunkPlan hplan;  // the chunks reproject_core will form: one upload event per chunk of items
  plan_chunks(std::min(ws->chunk_bytes, ws->max_bytes), 1, ws->min_lane_points, ws->min_lane_chunks, group, n, s, p, h, w, &hplan);
  const size_t item_pts = (size_t)s * hw;
  for (int i = 0; i < n; i += hplan.items_per_chunk) {
    const size_t cnt = (size_t)std::min(hplan.items_per_chunk, n - i) * item_pts;
    CU(cudaMemcpyAsync((char*)ws->s_rgb.p + i * item_pts * px, (const char*)rgb_host + i * item_pts * px, cnt * px,
                       cudaMemcpyHostToDevice, up));
    CU(cudaMemcpyAsync((float*)ws->s_depth.p + i * item_pts, depth_host + i * item_pts, cnt * 4,
                       cudaMemcpyHostToDevice, up));
    CU(cudaEventRecord(ws->pipe_ev[i], up));
  }
  HostPipe pipe{group, ws->d2h_stream, ws->pipe_ev.data(), ws->pipe_ev.data() + n, proj_image_host, proj_depth_host,
                proj_mask_host, winner_out_host};
  const int rc = reproject_core(ws, ws->s_rgb.p, rgb_dtype, (const float*)ws->s_depth.p, (const float*)ws->s_src.p,
                                (const float*)ws->s_tgt.p, nullptr, n, s, s, p, h, w, depth_scale, mask_proportion, mask_frames,
                                unproject_void, project_void, flags, (float*)ws->s_img.p, (float*)ws->s_dep.p,
                                (float*)ws->s_msk.p, winner_out_host ? (int32_t*)ws->s_win.p : nullptr, nullptr, st, &pipe);
  ws->host_pending = true;
  if (rc) {  // whatever was enqueued before the failure still uses the caller's buffers
    host_wait(ws);
    return rc;
  }
  return host_wait(ws);
}

int se3ds_ws_host_wait(se3ds_ws* ws) {
  if (!ws) return fail(SE3DS_ERR_BAD_ARG, "NULL workspace");
  GUARD(ws->device);
  return host_wait(ws);
}

int se3ds_apply_bin(const float* bin, float depth_scale, unsigned flags, float* proj_image, float* proj_depth,
                    float* proj_mask, int32_t* winner, void* stream) {
  if (!bin || !proj_image || !proj_depth || (!proj_mask && !(flags & SE3DS_FLAG_COMPACT_OUT))) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(proj_image));
  apply_bin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(bin, depth_scale, (flags & SE3DS_FLAG_RAW_FEATURES) != 0, (flags & SE3DS_FLAG_COMPACT_OUT) != 0, proj_image, proj_depth, proj_mask, winner);
  return launch_check("apply_bin_kernel");
}

int se3ds_filtered_coords_and_feats(const void* feats, int dtype, const float* depth, int n, int h, int w, int c,
                                    float depth_scale, float kinv_x, float kinv_y, float* xyz_out, float* feats_out,
                                    void* stream) {
  if (!feats || !depth || !xyz_out || !feats_out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(xyz_out));
  if (n < 0 || h <= 0 || w <= 0 || c <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "feats should have shape (N, H, W) or (N, H, W, C)");
  const long long total = (long long)n * h * w;
  if (total == 0) return SE3DS_OK;
  const int blocks = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SE3DS_U8) filtered_coords_kernel<uint8_t><<<blocks, kThreads, 0, st>>>((const uint8_t*)feats, depth, n, h, w, c, depth_scale, kinv_x, kinv_y, xyz_out, feats_out);
  else if (dtype == SE3DS_I32) filtered_coords_kernel<int><<<blocks, kThreads, 0, st>>>((const int*)feats, depth, n, h, w, c, depth_scale, kinv_x, kinv_y, xyz_out, feats_out);
  else if (dtype == SE3DS_F32) filtered_coords_kernel<float><<<blocks, kThreads, 0, st>>>((const float*)feats, depth, n, h, w, c, depth_scale, kinv_x, kinv_y, xyz_out, feats_out);
  else return fail(SE3DS_ERR_BAD_DTYPE, "dtype %d", dtype);
  return launch_check("filtered_coords_kernel");
}

int se3ds_pixel_rays(int output_height, float* out, void* stream) {
  if (!out || output_height <= 0) return fail(SE3DS_ERR_BAD_ARG, "bad argument");
  GUARD(device_of(out));
  const long long total = 2ll * output_height * output_height;
  pixel_rays_kernel<<<(int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32), kThreads, 0, (cudaStream_t)stream>>>(output_height, out);
  return launch_check("pixel_rays_kernel");
}

int se3ds_rotate_pano(const float* pano, const float* matrix, int n, int h, int w, int c, int output_height,
                      float* out, void* stream) {
  if (!pano || !matrix || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(out));
  if (n < 0 || h < 2 || c <= 0 || output_height <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "pano must be (N,H,W,C)");
  if (w != 2 * h) return fail(SE3DS_ERR_BAD_SHAPE, "Pano width must be twice height.");
  const long long total = (long long)n * output_height * 2 * output_height;
  if (total == 0) return SE3DS_OK;
  rotate_pano_kernel<<<(int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32), kThreads, 0, (cudaStream_t)stream>>>(
      pano, matrix, n, h, w, c, output_height, out);
  return launch_check("rotate_pano_kernel");
}

int se3ds_project_perspective_image(const float* image, int h, int w, int c, const float world_to_image[9],
                                    int output_height, int pad, float pad_value, int round_to_nearest, float* out,
                                    void* stream) {
  if (!image || !world_to_image || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(out));
  if (h < 1 || w < 1 || c <= 0 || output_height <= 0 || (!pad && (h < 2 || w < 2))) return fail(SE3DS_ERR_BAD_SHAPE, "image must be (H,W,C)");
  Mat3 m;
  memcpy(m.m, world_to_image, sizeof(m.m));
  const long long total = 2ll * output_height * output_height;
  persp_to_equirect_kernel<<<(int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32), kThreads, 0, (cudaStream_t)stream>>>(
      image, m, h, w, c, output_height, pad ? 1 : 0, pad_value, round_to_nearest, out);
  return launch_check("persp_to_equirect_kernel");
}

int se3ds_perspective_from_equirect(const float* image, int eq_h, int eq_w, int c, const float kinv_t[9],
                                    const float rotation[9], int height, int width, float* out, void* stream) {
  if (!image || !kinv_t || !rotation || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(out));
  if (eq_h < 2 || eq_w < 2 || c <= 0 || height <= 0 || width <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "image must be (H,W,C)");
  Mat3 a, r;
  memcpy(a.m, kinv_t, sizeof(a.m));
  memcpy(r.m, rotation, sizeof(r.m));
  const long long total = (long long)height * width;
  equirect_to_persp_kernel<<<(int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32), kThreads, 0, (cudaStream_t)stream>>>(
      image, a, r, eq_h, eq_w, c, height, width, out);
  return launch_check("equirect_to_persp_kernel");
}

int se3ds_resize(const void* in, int dtype, int n, int h, int w, int c, int out_h, int out_w, int bilinear,
                 void* out, void* stream) {
  if (!in || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(out));
  if (n < 0 || h <= 0 || w <= 0 || c <= 0 || out_h <= 0 || out_w <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "images must be (N,H,W,C)");
  const long long total = (long long)n * out_h * out_w;
  if (total == 0) return SE3DS_OK;
  const int blocks = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define RESIZE(TI, TO, B) resize_kernel<TI, TO, B><<<blocks, kThreads, 0, st>>>((const TI*)in, n, h, w, c, out_h, out_w, (TO*)out)
  if (bilinear) {
    if (dtype == SE3DS_U8) RESIZE(uint8_t, float, true);
    else if (dtype == SE3DS_I32) RESIZE(int, float, true);
    else if (dtype == SE3DS_F32) RESIZE(float, float, true);
    else return fail(SE3DS_ERR_BAD_DTYPE, "dtype %d", dtype);
  } else {
    if (dtype == SE3DS_U8) RESIZE(uint8_t, uint8_t, false);
    else if (dtype == SE3DS_I32) RESIZE(int, int, false);
    else if (dtype == SE3DS_F32) RESIZE(float, float, false);
    else return fail(SE3DS_ERR_BAD_DTYPE, "dtype %d", dtype);
  }
#undef RESIZE
  return launch_check("resize_kernel");
}

int se3ds_interpolate_bilinear(const float* grid, const float* query_points, int b, int h, int w, int c,
                               long long num_queries, int indexing_xy, float* out, void* stream) {
  if (!grid || !query_points || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(out));
  if (b < 0 || h < 2 || w < 2 || c <= 0 || num_queries < 0) return fail(SE3DS_ERR_BAD_SHAPE, "Grid must be at least 2x2 with shape (B,H,W,C)");
  const long long total = (long long)b * num_queries;
  if (total == 0) return SE3DS_OK;
  const int blocks = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32);
  interpolate_bilinear_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(grid, query_points, b, h, w, c, num_queries,
                                                                            indexing_xy, out);
  return launch_check("interpolate_bilinear_kernel");
}

int se3ds_expand_guidance(const uint8_t* rgb_u8, const float* proj_depth, long long njobs, long long px_per_job,
                          const int32_t* job_map, float* proj_image, float* depth_out, float* proj_mask, void* stream) {
  if (!rgb_u8 || !proj_depth || !proj_image || !proj_mask) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  if (njobs < 0 || px_per_job < 0) return fail(SE3DS_ERR_BAD_SHAPE, "negative size");
  if (njobs == 0 || px_per_job == 0) return SE3DS_OK;
  GUARD(device_of(proj_image));
  const int vec = px_per_job % 4 == 0 && aligned(rgb_u8, 4) && aligned(proj_depth, 16) && aligned(proj_image, 16) &&
                  aligned(proj_mask, 16) && (!depth_out || aligned(depth_out, 16));
  const long long work = njobs * (vec ? px_per_job / 4 : px_per_job);
  expand_guidance_kernel<<<(int)std::min<long long>((work + kThreads - 1) / kThreads, 148 * 16), kThreads, 0, (cudaStream_t)stream>>>(
      rgb_u8, proj_depth, njobs, px_per_job, vec, job_map, proj_image, depth_out, proj_mask);
  return launch_check("expand_guidance_kernel");
}

int se3ds_quantize_rgb(const float* image, int n, long long elems_per_item, int32_t* out, long long out_item_stride,
                       void* stream) {
  if (!image || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  if (n < 0 || elems_per_item < 0 || out_item_stride < elems_per_item) return fail(SE3DS_ERR_BAD_SHAPE, "bad shape");
  const long long total = (long long)n * elems_per_item;
  if (total == 0) return SE3DS_OK;
  GUARD(device_of(out));
  quantize_rgb_kernel<<<(int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32), kThreads, 0, (cudaStream_t)stream>>>(
      image, elems_per_item, n, out, out_item_stride);
  return launch_check("quantize_rgb_kernel");
}

int se3ds_proportion_invalid(const float* offsets, int p, const float* depth, int h, int w,
                             float distance_padding, float depth_scale, float* out, void* stream) {
  if (!offsets || !depth || !out) return fail(SE3DS_ERR_BAD_ARG, "NULL argument");
  GUARD(device_of(out));
  if (p < 0 || h <= 0 || w <= 0) return fail(SE3DS_ERR_BAD_SHAPE, "depth_image must be (H, W)");
  if (p == 0) return SE3DS_OK;
  proportion_invalid_kernel<<<p, kThreads, 0, (cudaStream_t)stream>>>(offsets, depth, h, w, distance_padding, depth_scale, out);
  return launch_check("proportion_invalid_kernel");
}

}  // extern "C"
