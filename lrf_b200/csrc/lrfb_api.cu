// C ABI of liblrfb.so (see include/lrfb.h).  Host-side orchestration of the sm_100a kernels.
// The same file compiles with g++ -DLRFB_SIM for the test-only CPU SIMT shim (tests/cpu_sim/).
#include "../../include/lrfb.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include <zlib.h>

#include "bcd.cuh"
#include "bcd_resident.cuh"
#include "bcd_tc.cuh"
#include "decode.cuh"
#include "deflate9.cuh"
#include "inflate9.cuh"
#include "eig.cuh"
#include "frontend.cuh"
#include "frontgram.cuh"
#include "gram.cuh"
#include "gram_i8.cuh"
#include "gram_u8.cuh"
#include "lrfb_common.cuh"
#include "svdcodec.cuh"

using namespace lrfb;

#define LRFB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local char g_err[512] = "";
long long g_launches = 0;  // kernels launched by this library (bench.py reports it as gpu_launches)

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#ifdef LRFB_SIM
int dev_copy(void* dst, const void* src, size_t n, cudaStream_t) {
  memcpy(dst, src, n);
  return 0;
}
int check_launch(const char*) {
  ++g_launches;
  return 0;
}
int num_sms() { return 4; }
#else
int dev_copy(void* dst, const void* src, size_t n, cudaStream_t st) {
  return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st);
}
int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
int num_sms() {
  static thread_local int cached = 0;
  if (!cached) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}
#endif

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Environment knobs exist only in development builds (nvcc -DLRFB_DEV): the shipped library reads no environment.
inline const char* dev_getenv(const char* name) {
#ifdef LRFB_DEV
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}
std::atomic<int> g_decode_v1{0};  // lrfb_debug_set("decode_v1", 1): per-row float decoder instead of the DP4A one

struct Geometry {
  FrontParams fp;
  lrfb_qmf_layout lay;
};

int make_geometry(const lrfb_qmf_config* c, int batch, Geometry* g) {
  if (!c) return fail(LRFB_E_ARG, "config is null");
  if (c->height <= 0 || c->width <= 0 || c->patch_h <= 0 || c->patch_w <= 0)
    return fail(LRFB_E_ARG, "non-positive image or patch size");
  if (c->color_space != LRFB_RGB && c->color_space != LRFB_YCBCR)
    return fail(LRFB_E_ARG, "color_space must be LRFB_RGB or LRFB_YCBCR");
  if (c->input_dtype != LRFB_U8 && c->input_dtype != LRFB_F32) return fail(LRFB_E_ARG, "bad input_dtype");
  const int ycbcr = c->color_space == LRFB_YCBCR;
  memset(g, 0, sizeof(*g));
  FrontParams& fp = g->fp;
  fp.H = c->height, fp.W = c->width, fp.p = c->patch_h, fp.q = c->patch_w, fp.ycbcr = ycbcr, fp.n_img = batch;
  lrfb_qmf_layout& L = g->lay;
  L.n_planes = ycbcr ? 3 : 1;
  L.cols = (ycbcr ? 1 : 3) * c->patch_h * c->patch_w;
  int64_t off = 0;
  for (int pl = 0; pl < L.n_planes; ++pl) {
    int h = c->height, w = c->width;
    if (ycbcr && pl > 0) {
      if (!(c->scale_h > 0.0) || !(c->scale_w > 0.0)) return fail(LRFB_E_ARG, "scale_factor must be positive");
      h = (int)floor((double)c->height * c->scale_h);  // F.interpolate(scale_factor): floor(in*scale)
      w = (int)floor((double)c->width * c->scale_w);
      if (h <= 0 || w <= 0) return fail(LRFB_E_UNSUPPORTED, "chroma plane would be empty");
    }
    int padh = (c->patch_h - h % c->patch_h) % c->patch_h;  // compression/utils.py:125-130
    int padw = (c->patch_w - w % c->patch_w) % c->patch_w;
    PlaneGeom& pg = fp.g[pl];
    pg.h = h, pg.w = w, pg.hp = h + padh, pg.wp = w + padw, pg.top = padh / 2, pg.left = padw / 2;
    if (padh - pg.top >= h || padw - pg.left >= w)
      return fail(LRFB_E_UNSUPPORTED, "reflect padding needs pad < dimension (plane %d is %dx%d)", pl, h, w);
    pg.nbw = pg.wp / c->patch_w;
    pg.rows = (pg.hp / c->patch_h) * pg.nbw;
    L.orig_h[pl] = h, L.orig_w[pl] = w, L.pad_h[pl] = pg.hp, L.pad_w[pl] = pg.wp, L.rows[pl] = pg.rows;
    L.rank[pl] = c->rank[pl];
    if (c->rank[pl] <= 0) return fail(LRFB_E_ARG, "rank[%d] must be positive", pl);
    if (c->rank[pl] > kGenMaxR) return fail(LRFB_E_UNSUPPORTED, "rank %d > %d", c->rank[pl], kGenMaxR);
    L.u_offset[pl] = off;
    off += (int64_t)pg.rows * c->rank[pl];
    L.v_offset[pl] = off;
    off += (int64_t)L.cols * c->rank[pl];
    L.x_floats += (int64_t)pg.rows * L.cols;
  }
  L.record_bytes = off;
  if (L.cols > 1024) return fail(LRFB_E_UNSUPPORTED, "patch too large (N=%d)", L.cols);
  return 0;
}

int check_bounds(float lo, float hi) {
  if (!(ceilf(lo) <= floorf(hi))) return fail(LRFB_E_ARG, "empty bounds");
  if (ceilf(lo) < -128.0f || floorf(hi) > 127.0f)
    return fail(LRFB_E_UNSUPPORTED, "bounds outside int8 are not implemented");
  return 0;
}

int make_map(const Geometry& g, int batch, lrfb_qmf_workspace_map* m) {
  memset(m, 0, sizeof(*m));
  int64_t off = 0;
  const lrfb_qmf_layout& L = g.lay;
  for (int pl = 0; pl < L.n_planes; ++pl) {
    m->x[pl] = off;
    off = align_up(off + (int64_t)batch * L.rows[pl] * L.cols * 4, 256);
  }
  for (int pl = 0; pl < L.n_planes; ++pl) {
    m->u[pl] = off;
    off = align_up(off + (int64_t)batch * L.rows[pl] * L.rank[pl] * 4, 256);
    m->v[pl] = off;
    off = align_up(off + (int64_t)batch * L.cols * L.rank[pl] * 4, 256);
    m->gram[pl] = off;
    off = align_up(off + (int64_t)batch * L.cols * L.cols * 8, 256);
    m->evec[pl] = off;
    off = align_up(off + (int64_t)batch * L.cols * L.rank[pl] * 8, 256);
    m->sigma[pl] = off;
    off = align_up(off + (int64_t)batch * L.rank[pl] * 12 + 16, 256);  // f64 + f32 singular values, work counter
  }
  m->total_bytes = off;
  return 0;
}

// ---- factorisation of one batch of equally shaped matrices ------------------------------------------
struct FactorWs {  // scratch beyond x/u/v/gram/evec/sigma
  static int gram_split(int n_mat, int M) {
    int tiles = (M + kDmmaTileRows - 1) / kDmmaTileRows;
    int want = (4 * num_sms() + n_mat - 1) / n_mat;
    return std::max(1, std::min(std::min(want, tiles), 32));
  }
  static int64_t bytes(int n_mat, int M, int N, int R) {
    int split = gram_split(n_mat, M);
    int64_t b = 0;
    if (split > 1) b += align_up((int64_t)n_mat * split * N * N * 8, 256);
    b += align_up((int64_t)n_mat * (int64_t)EigScratch::doubles(N, R) * 8, 256);
    b += align_up((int64_t)kGenGrid * gen_scratch_floats(N, R) * 4, 256);  // bcd_generic scratch
    return b;
  }
};

template <int BPT>
void launch_gram(const float* x, long long xs, int M, int N, double* out, int split, int n_mat, int threads,
                 int nblocks, cudaStream_t st) {
  size_t smem = (size_t)kGramTileRows * ((N + 3) & ~3) * 8;
#ifndef LRFB_SIM
  if (smem > 48 * 1024) cudaFuncSetAttribute(gram_kernel<BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  int gz = (nblocks + threads * BPT - 1) / (threads * BPT);
  LRFB_LAUNCH(gram_kernel<BPT>, dim3(split, n_mat, gz), dim3(threads), smem, st, x, xs, M, N, out, split);
}

template <int R, int TM, int NT>
int launch_bcd_cfg(const BcdBatch& b, cudaStream_t st) {
  constexpr int N = 64;
  auto kern = bcd_kernel<N, R, TM, NT>;
  size_t smem = sizeof(BcdSmem<N, R, TM, NT>);
  int per_sm = 2;
#ifndef LRFB_SIM
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail((int)e, "bcd smem attribute: %s", cudaGetErrorString(e));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
  if (per_sm < 1) per_sm = 1;
#endif
  int grid = std::min(b.n_mat, num_sms() * per_sm);
  LRFB_LAUNCH(kern, dim3(grid), dim3(NT), smem, st, b);
  return check_launch("bcd_kernel");
}

template <int R>
int launch_bcd_fast(const BcdBatch& b, cudaStream_t st) {
  static int variant = -1;  // dev knob: LRFB_BCD_VARIANT=0 (128 rows, 64 thr) | 1 (128, 128) | 2 (64, 64)
  if (variant < 0) {
    const char* e = dev_getenv("LRFB_BCD_VARIANT");
    variant = e ? atoi(e) : 1;
  }
#ifdef LRFB_DEV  // alternative shapes are instantiated in development builds only
  if (variant == 0) return launch_bcd_cfg<R, 128, 64>(b, st);
  if (variant == 2) return launch_bcd_cfg<R, 64, 64>(b, st);
#endif
  return launch_bcd_cfg<R, 128, 128>(b, st);
}

// shared-memory-resident cluster kernel: N = 64, R <= 4.  Two shapes: 384 rows x 192 threads per CTA with
// 2 CTAs per SM and clusters of up to 16 (default: the serial per-sweep tail of one cluster overlaps the
// compute of the other), or 768 rows x 384 threads with 1 CTA per SM and clusters of up to 8.
template <int R, int ROWS, int NT, int MAXC>
int launch_bcd_resident_cfg(const BcdBatch& b, cudaStream_t st) {
  const int need = (b.M + ROWS - 1) / ROWS;
  int csize = 1;
  while (csize < need) csize *= 2;
  const int rows_per_cta = (b.M + csize - 1) / csize;
  auto kern = bcd_resident_kernel<R, ROWS, NT>;
  const size_t smem = sizeof(ResSmem<R, ROWS, NT>);
#ifdef LRFB_SIM
  (void)MAXC;
  LRFB_LAUNCH(kern, dim3(std::min(b.n_mat, 2)), dim3(NT), smem, st, b, 1, rows_per_cta);
  return check_launch("bcd_resident_kernel");
#else
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail((int)e, "bcd_resident smem attribute: %s", cudaGetErrorString(e));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (csize > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fail((int)e, "non-portable cluster size: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(NT), cfg.dynamicSmemBytes = smem, cfg.stream = st, cfg.attrs = attr, cfg.numAttrs = 1;
  cfg.gridDim = dim3(csize);
  int max_clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (e != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    max_clusters = std::max(1, num_sms() * (ROWS <= 384 ? 2 : 1) / csize);
  }
  cfg.gridDim = dim3((unsigned)(std::min(b.n_mat, max_clusters) * csize));
  if (dev_getenv("LRFB_DEBUG"))
    fprintf(stderr, "[lrfb] bcd_resident R=%d rows/cta=%d threads=%d cluster=%d max_active_clusters=%d grid=%u smem=%zu\n",
            R, ROWS, NT, csize, max_clusters, cfg.gridDim.x, smem);
  e = cudaLaunchKernelEx(&cfg, kern, b, csize, rows_per_cta);
  if (e != cudaSuccess) return fail((int)e, "bcd_resident launch: %s", cudaGetErrorString(e));
  return check_launch("bcd_resident_kernel");
#endif
}

#ifndef LRFB_SIM
// resident sweeps with the V-phase on tcgen05 (int8, A operand in TMEM): N = 64, R <= 4, X in [0, 256)
// 3-D tensor map over X [n_mat][M][64] f32 with a 32-column x 256-row box and the 128-byte swizzle (see x_off in
// bcd_tc.cuh).  The encoder lives in the driver library: fetched through the runtime, no link-time dependency.
bool make_x_tensor_map(const BcdBatch& b, CUtensorMap* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (!dev_getenv("LRFB_NO_TMA") &&
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    cudaGetLastError();
  }
  if (!encode || (reinterpret_cast<uintptr_t>(b.X) & 15) || ((b.x_stride * 4) & 15)) return false;
  const cuuint64_t dims[3] = {64, (cuuint64_t)b.M, (cuuint64_t)b.n_mat};
  const cuuint64_t strides[2] = {256, (cuuint64_t)b.x_stride * 4};
  const cuuint32_t box[3] = {32, 256, 1}, estr[3] = {1, 1, 1};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(b.X), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int R, int ROWS, int NT, int MAXC = (ROWS > 384 ? 8 : 16)>
int launch_bcd_tc_cfg(const BcdBatch& b, cudaStream_t st) {
  const int need = (b.M + ROWS - 1) / ROWS;
  int csize = 1;
  while (csize < need) csize *= 2;
  if (csize > MAXC) return fail(LRFB_E_UNSUPPORTED, "bcd_tc: %d rows need a cluster of %d > %d", b.M, csize, MAXC);
  const int rows_per_cta = (b.M + csize - 1) / csize;
  auto kern = bcd_tc_kernel<R, ROWS, NT, MAXC>;
  const size_t smem = sizeof(TcSmem<R, ROWS, NT>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail((int)e, "bcd_tc smem attribute: %s", cudaGetErrorString(e));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (csize > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fail((int)e, "non-portable cluster size: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  // 2 CTAs per SM only pay off when they belong to DIFFERENT clusters (CTAs of one cluster move in lock-step)
  attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
  attr[1].val.clusterSchedulingPolicyPreference = cudaClusterSchedulingPolicySpread;
  static const int policy = dev_getenv("LRFB_TC_POLICY") ? atoi(dev_getenv("LRFB_TC_POLICY")) : 1;
  attr[1].val.clusterSchedulingPolicyPreference = (cudaClusterSchedulingPolicy)policy;
  cfg.blockDim = dim3(NT), cfg.dynamicSmemBytes = smem, cfg.stream = st, cfg.attrs = attr;
  cfg.numAttrs = (ROWS <= 384 && csize > 1) ? 2 : 1;
  cfg.gridDim = dim3(csize);
  int max_clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (e != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    max_clusters = std::max(1, num_sms() / csize);
  }
  // the occupancy calculation assumes one CTA per SM for any kernel that allocates tensor memory; two 384-row CTAs do
  // share an SM (tools/probes/occ_probe.cu).  Clusters that find no SM pair start late and find no work left.
  if (ROWS <= 384) max_clusters *= 2;
  if (const char* ov = dev_getenv("LRFB_TC_CLUSTERS")) max_clusters = std::max(1, atoi(ov) * 8 / csize);  // dev knob (per 8-CTA unit)
  cfg.gridDim = dim3((unsigned)(std::min(b.n_mat, max_clusters) * csize));
  if (dev_getenv("LRFB_DEBUG")) {
    int per_sm = -1;
    cudaFuncAttributes fa;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    cudaFuncGetAttributes(&fa, kern);
    fprintf(stderr, "[lrfb] bcd_tc blocks/SM=%d regs=%d static_smem=%zu max_dyn=%d\n", per_sm, fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
  }
  if (dev_getenv("LRFB_DEBUG"))
    fprintf(stderr, "[lrfb] bcd_tc R=%d rows/cta=%d threads=%d cluster=%d max_active_clusters=%d grid=%u smem=%zu\n",
            R, ROWS, NT, csize, max_clusters, cfg.gridDim.x, smem);
  CUtensorMap x_map;
  memset(&x_map, 0, sizeof(x_map));
  const int use_tma = (ROWS == 768 && make_x_tensor_map(b, &x_map)) ? 1 : 0;  // box rows = ROWS / 3 = 256
  e = cudaLaunchKernelEx(&cfg, kern, b, csize, rows_per_cta, x_map, use_tma);
  if (e != cudaSuccess) return fail((int)e, "bcd_tc launch: %s", cudaGetErrorString(e));
  return check_launch("bcd_tc_kernel");
}
int tc_variant() {
  static int v = -1;  // dev knob: LRFB_TC_VARIANT=0 (768 rows x 384 threads, 1 CTA/SM) | 1 (384 x 192, 2 CTAs/SM)
  if (v < 0) {
    const char* e = dev_getenv("LRFB_TC_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}
template <int R>
int launch_bcd_tc(const BcdBatch& b, cudaStream_t st) {
#ifdef LRFB_DEV  // alternative shapes are instantiated in development builds only
  if (tc_variant() == 1 || (tc_variant() == 2 && R <= 2)) return launch_bcd_tc_cfg<R, 384, 192>(b, st);
#endif
  if (b.M > 8 * kTcRows) return launch_bcd_tc_cfg<R, 768, 384, 16>(b, st);  // clusters of 16: up to 12 288 rows resident
  // 256 threads x 3 rows with two of a thread's rows in registers: the FMA chains are bound by the shared-memory reads
  // of X, which this shape cuts from 1/2 to 1/3 of the rows (luma 11.93 -> 11.68 ms, step 19.4 -> 19.05 ms)
  if (tc_variant() != 6) return launch_bcd_tc_cfg<R, 768, 256>(b, st);
  return launch_bcd_tc_cfg<R, 768, 384>(b, st);
}
bool tc_enabled() {
  static int v = -1;  // dev knob: LRFB_BCD_TC=0 keeps the FFMA V-phase
  if (v < 0) {
    const char* e = dev_getenv("LRFB_BCD_TC");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}
#endif

int resident_variant() {
  static int v = -1;  // dev knob: LRFB_RES_VARIANT=0 (768 rows, 1 CTA/SM) | 1 (384 rows, 2 CTAs/SM)
  if (v < 0) {
    const char* e = dev_getenv("LRFB_RES_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int R>
int launch_bcd_resident(const BcdBatch& b, cudaStream_t st) {
#ifdef LRFB_DEV  // alternative shapes are instantiated in development builds only
  if (resident_variant() == 2) return launch_bcd_resident_cfg<R, 768, 256, 8>(b, st);
  if (resident_variant() == 1) return launch_bcd_resident_cfg<R, 384, 192, 16>(b, st);
#endif
  return launch_bcd_resident_cfg<R, 768, 384, 8>(b, st);
}

bool resident_ok(int N, int R, int M) {
  static int enabled = -1;  // dev knob: LRFB_BCD_RESIDENT=0 forces the streaming kernel
  if (enabled < 0) {
    const char* e = dev_getenv("LRFB_BCD_RESIDENT");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled || N != 64 || R > 4 || bmm_native(N, M, R)) return false;
#ifdef LRFB_SIM
  return M <= (resident_variant() == 1 ? 384 : 768);  // the CPU shim has no clusters
#else
  return M <= (resident_variant() == 1 ? 16 * 384 : 8 * 768);
#endif
}

// the tensor-core sweeps kernel additionally runs 768-row CTAs in (non-portable) clusters of 16
bool tc_resident_ok(int N, int R, int M, bool u8_range) {
#ifdef LRFB_SIM
  (void)N, (void)R, (void)M, (void)u8_range;
  return false;
#else
  return u8_range && tc_enabled() && N == 64 && R <= 4 && !bmm_native(N, M, R) && M <= 16 * kTcRows;
#endif
}

// `counter`: 4 bytes of device memory private to this call (the tensor-core kernel hands out matrices dynamically)
int run_bcd(const BcdBatch& b0, int N, int R, float* bwork, cudaStream_t st, int* counter) {
  BcdBatch b = b0;
  b.work_counter = counter;
#ifndef LRFB_SIM
  if (tc_resident_ok(N, R, b.M, b.x_u8_range != 0)) {
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return fail((int)e, "work counter reset: %s", cudaGetErrorString(e));
    switch (R) {
      case 1: return launch_bcd_tc<1>(b, st);
      case 2: return launch_bcd_tc<2>(b, st);
      case 3: return launch_bcd_tc<3>(b, st);
      default: return launch_bcd_tc<4>(b, st);
    }
  }
#endif
  if (resident_ok(N, R, b.M)) {
    switch (R) {
      case 1: return launch_bcd_resident<1>(b, st);
      case 2: return launch_bcd_resident<2>(b, st);
      case 3: return launch_bcd_resident<3>(b, st);
      default: return launch_bcd_resident<4>(b, st);
    }
  }
  const bool fast = (N == 64 && R <= 4 && !bmm_native(N, b.M, R));
  if (fast) {
    switch (R) {
      case 1: return launch_bcd_fast<1>(b, st);
      case 2: return launch_bcd_fast<2>(b, st);
      case 3: return launch_bcd_fast<3>(b, st);
      default: return launch_bcd_fast<4>(b, st);
    }
  }
  int grid = std::min(b.n_mat, kGenGrid);
  LRFB_LAUNCH(bcd_generic_kernel, dim3(grid), dim3(256), 0, st, b, N, R, bwork);
  return check_launch("bcd_generic_kernel");
}

// SVD init (unless injected) + sweeps.  u, v: f32 working/output buffers.
int factorize_batch(const float* x, int n_mat, int M, int N, int R, float lo, float hi, int iters, float* u,
                    float* v, int8_t* uq, int8_t* vq, long long q_stride, const float* init_u,
                    const float* init_v, const int32_t* sign_flip, double* gram, double* evec, double* sigma,
                    unsigned char* scratch, int stop_after_init, cudaStream_t st, int phase = 0,
                    bool x_in_u8_range = false, bool gram_done = false, bool x_integer_u8 = false) {
  // phase 0: init + sweeps, 1: init only, 2: sweeps only (after a phase-1 call with the same arguments)
  int rc;
  const bool injected = init_u && init_v;
  float* s0 = injected ? nullptr : reinterpret_cast<float*>(sigma + (size_t)n_mat * R);
  const int split = FactorWs::gram_split(n_mat, M);
  double* gram_part = nullptr;
  if (split > 1) {
    gram_part = reinterpret_cast<double*>(scratch);
    scratch += align_up((int64_t)n_mat * split * N * N * 8, 256);
  }
  double* eig_scratch = reinterpret_cast<double*>(scratch);
  scratch += align_up((int64_t)n_mat * (int64_t)EigScratch::doubles(N, R) * 8, 256);
  float* bwork = reinterpret_cast<float*>(scratch);

  if (phase == 2) {
    // nothing to initialise
  } else if (injected) {
    if ((rc = dev_copy(u, init_u, (size_t)n_mat * M * R * 4, st))) return fail(rc, "copy init u");
    if ((rc = dev_copy(v, init_v, (size_t)n_mat * N * R * 4, st))) return fail(rc, "copy init v");
  } else {
    // G = X^T X in f64 (unless the fused front end has produced it already)
    const int nb = ((N + 3) / 4);
    const int nblocks = nb * (nb + 1) / 2;
    double* gout = split > 1 ? gram_part : gram;
    for (int m0 = 0; m0 < n_mat && !gram_done; m0 += 65535) {
      int cnt = std::min(65535, n_mat - m0);
      const float* xx = x + (size_t)m0 * M * N;
      double* go = gout + (size_t)m0 * split * N * N;
#ifndef LRFB_SIM
      static int use_i8 = -1;  // dev knob: LRFB_GRAM_I8=0 keeps the FP64 (DMMA) Gram for uint8-range planes too
      if (use_i8 < 0) {
        const char* ev = dev_getenv("LRFB_GRAM_I8");
        use_i8 = ev ? atoi(ev) : 1;
      }
      if (N == kU8N && x_integer_u8 && use_i8 && (M + split - 1) / split <= 60000) {
        // RGB planes of uint8 images: the bytes are the operands (one slice)
        const size_t usmem = sizeof(GramU8Smem) + 1024;
        cudaError_t e = cudaFuncSetAttribute(gram192_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem);
        if (e != cudaSuccess) return fail((int)e, "gram_u8 smem attribute: %s", cudaGetErrorString(e));
        gram192_u8_kernel<<<dim3(split, cnt), dim3(kI8Threads), usmem, st>>>(xx, (long long)M * N, M, go, split);
      } else if (N == 64 && x_in_u8_range && use_i8 && (M + split - 1) / split <= 60000) {
        // exact int8 tensor-core Gram (tcgen05): entries of X are in [0, 256)
        const size_t ismem = sizeof(GramI8Smem) + 1024;
        cudaError_t e = cudaFuncSetAttribute(gram64_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem);
        if (e != cudaSuccess) return fail((int)e, "gram_i8 smem attribute: %s", cudaGetErrorString(e));
        gram64_i8_kernel<<<dim3(split, cnt), dim3(kI8Threads), ismem, st>>>(xx, (long long)M * N, M, go, split);
      } else
#endif
      if (N == 64) {
        LRFB_LAUNCH(gram64_dmma_kernel, dim3(split, cnt), dim3(128), 0, st, xx, (long long)M * N, M, go, split);
      } else if (nblocks <= 160) {
        launch_gram<1>(xx, (long long)M * N, M, N, go, split, cnt, ((nblocks + 31) / 32) * 32, nblocks, st);
      } else {
        int threads = std::min(256, (((nblocks + 4) / 5 + 31) / 32) * 32);
        launch_gram<5>(xx, (long long)M * N, M, N, go, split, cnt, threads, nblocks, st);
      }
      if ((rc = check_launch("gram_kernel"))) return rc;
      if (split > 1) {
        LRFB_LAUNCH(gram_reduce_kernel, dim3((N * N + 255) / 256, cnt), dim3(256), 0, st, go,
                    gram + (size_t)m0 * N * N, N * N, split);
        if ((rc = check_launch("gram_reduce_kernel"))) return rc;
      }
    }
    // top-R eigenpairs: one warp per matrix
    const int use_shared = EigScratch::fits_shared(N, R);
    const size_t eig_smem = use_shared ? ((size_t)N * N + EigScratch::doubles(N, R)) * 8 : 0;
#ifndef LRFB_SIM
    if (eig_smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eig_topr_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem);
      if (e != cudaSuccess) return fail((int)e, "eig smem attribute: %s", cudaGetErrorString(e));
    }
#endif
    // s0: f32 singular values stored behind the f64 ones
    static int eig_v3 = -1;  // dev knob: LRFB_EIG_V3=0 selects the one-warp shared-memory kernel
    if (eig_v3 < 0) {
      const char* ev = dev_getenv("LRFB_EIG_V3");
#ifdef LRFB_SIM
      eig_v3 = ev ? atoi(ev) : 0;  // the 2-warp kernel is barrier-heavy: minutes per matrix on the CPU shim
#else
      eig_v3 = ev ? atoi(ev) : 1;
#endif
    }
    if (use_shared && eig_v3)
      LRFB_LAUNCH(eig64_topr_kernel, dim3(n_mat), dim3(64), 0, st, gram, R, evec, sigma, sign_flip, M, v, s0, x,
                  (long long)M * N);
    else if (use_shared)
      LRFB_LAUNCH(eig_topr_kernel<64>, dim3(n_mat), dim3(32), eig_smem, st, gram, N, R, eig_scratch, evec, sigma,
                  sign_flip, use_shared, M, v, s0, x, (long long)M * N);
    else
      LRFB_LAUNCH(eig_topr_kernel<0>, dim3(n_mat), dim3(32), 0, st, gram, N, R, eig_scratch, evec, sigma,
                  sign_flip, use_shared, M, v, s0, x, (long long)M * N);
    if ((rc = check_launch("eig_topr_kernel"))) return rc;
  }
  if (stop_after_init || phase == 1) return 0;
  BcdBatch b;
  b.X = x, b.x_stride = (long long)M * N, b.U = u, b.V = v, b.Uq = uq, b.Vq = vq;
  b.uq_stride = b.vq_stride = q_stride;
  b.M = M, b.n_mat = n_mat, b.num_iters = iters, b.lo = ceilf(lo), b.hi = floorf(hi);
  b.s0 = s0;
  b.x_u8_range = x_in_u8_range ? 1 : 0;
  if (iters <= 0) return fail(LRFB_E_UNSUPPORTED, "num_iters must be >= 1");
  return run_bcd(b, N, R, bwork, st, reinterpret_cast<int*>(reinterpret_cast<char*>(sigma) + (size_t)n_mat * R * 12));
}

#ifndef LRFB_SIM
// Helper stream per device so the chroma sweeps can fill the SMs the luma clusters leave idle.
// One set per device, shared by every caller: `mu` is held from the first record to the last wait of a call (a call only
// ENQUEUES work, so the critical section is short), which keeps two host threads from interleaving their
// cudaEventRecord / cudaStreamWaitEvent pairs on the shared events.  A partially initialised set is never handed out.
struct SideStream {
  std::mutex mu;
  bool ready = false, failed = false;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, init2 = nullptr;
};
SideStream* side_stream(int which = 0) {  // returns with s->mu LOCKED (or nullptr); set 0: encode, set 1: lossless stage
  static SideStream table[2][64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = table[which][dev];
  s.mu.lock();
  if (!s.ready && !s.failed) {
    // lowest priority: luma-chain blocks (caller's stream) are scheduled first, chroma fills what is left
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    bool ok = cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, least) == cudaSuccess &&
              cudaStreamCreateWithPriority(&s.stream2, cudaStreamNonBlocking, least) == cudaSuccess &&
              cudaEventCreateWithFlags(&s.init2, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) cudaGetLastError();
    s.ready = ok, s.failed = !ok;
  }
  if (!s.ready) {
    s.mu.unlock();
    return nullptr;
  }
  return &s;
}
struct SideUnlock {  // releases the set on every exit path of lrfb_qmf_encode
  SideStream* s;
  ~SideUnlock() {
    if (s) s->mu.unlock();
  }
};
#endif

int64_t encode_scratch_bytes(const Geometry& g, int batch) {
  int64_t mx = align_up((int64_t)batch * 16, 256);  // svd codec: per-image (min, max) of u and v
  int64_t fw = 0;
  for (int pl = 0; pl < g.lay.n_planes; ++pl)
    fw = std::max(fw, FactorWs::bytes(batch, g.lay.rows[pl], g.lay.cols, g.lay.rank[pl]));
  return mx + fw;
}

// vectorised kernels apply: 8x8 patches, YCbCr, W % 16 == 0, chroma exactly half width, aligned base
bool fast8_geometry(const Geometry& g, const void* base) {
  const FrontParams& f = g.fp;
  return f.ycbcr && f.p == 8 && f.q == 8 && f.W % 16 == 0 && f.g[1].w * 2 == f.W &&
         ((uintptr_t)base % 16) == 0;
}

// one pass over the image for all three planes: additionally H % 16 == 0 and chroma exactly half height (no padding)
bool fused8_geometry(const lrfb_qmf_config* cfg, const Geometry& g, const void* base) {
  const FrontParams& f = g.fp;
  return cfg->input_dtype == LRFB_U8 && fast8_geometry(g, base) && f.H % 16 == 0 && f.g[1].h * 2 == f.H &&
         f.g[0].hp == f.H && f.g[1].hp * 2 == f.H && g.lay.n_planes == 3;
}

// front end fused with the luma Gram (frontgram.cuh): additionally W % 256 == 0, N = 64, no Gram row split (large batch)
bool frontgram_ok(const lrfb_qmf_config* cfg, const Geometry& g, int batch) {
#ifdef LRFB_SIM
  (void)cfg, (void)g, (void)batch;
  return false;
#else
  static const bool off = dev_getenv("LRFB_NO_FRONTGRAM") != nullptr;
  return !off && cfg->width % 256 == 0 && g.lay.cols == 64 && g.lay.rows[0] <= 60000 &&
         FactorWs::gram_split(batch, g.lay.rows[0]) == 1;
#endif
}

// planes: bit 0 = plane 0 (luma / RGB), bit 1 = planes 1 and 2 (chroma)
int run_frontend(const lrfb_qmf_config* cfg, const Geometry& g, int batch, const void* d_images,
                 float* const* xs, cudaStream_t st, int planes = 3) {
  if (planes == 3 && fused8_geometry(cfg, g, d_images)) {
    dim3 grid((unsigned)((g.fp.g[0].nbw + 31) / 32), (unsigned)(g.fp.H / 16), (unsigned)std::min(batch, 65535));
    LRFB_LAUNCH(frontend8_fused_kernel, grid, dim3(256), 0, st, (const unsigned char*)d_images, xs[0], xs[1], xs[2], g.fp);
    return check_launch("frontend8_fused_kernel");
  }
  if (cfg->input_dtype == LRFB_U8 && fast8_geometry(g, d_images)) {
    const unsigned char* img = (const unsigned char*)d_images;
    if (planes & 1) {
      long long items = (long long)g.fp.g[0].hp * g.fp.g[0].nbw;
      dim3 grid((unsigned)std::min<long long>((items + 255) / 256, 4096), std::min(batch, 65535));
      LRFB_LAUNCH(frontend8_luma_kernel, grid, dim3(256), 0, st, img, xs[0], g.fp);
      int rc = check_launch("frontend8_luma_kernel");
      if (rc) return rc;
    }
    if (planes & 2) {
      long long items = (long long)g.fp.g[1].hp * g.fp.g[1].nbw;
      dim3 grid2((unsigned)std::min<long long>((items + 255) / 256, 4096), std::min(batch, 65535));
      LRFB_LAUNCH(frontend8_chroma_kernel, grid2, dim3(256), 0, st, img, xs[1], xs[2], g.fp);
      return check_launch("frontend8_chroma_kernel");
    }
    return 0;
  }
  if (!g.fp.ycbcr && cfg->input_dtype == LRFB_U8 && g.fp.p == 8 && g.fp.q == 8 && g.fp.W % 8 == 0 &&
      g.fp.g[0].wp == g.fp.W && ((uintptr_t)d_images % 8) == 0 && (planes & 1)) {
    long long items = 3LL * g.fp.g[0].hp * g.fp.g[0].nbw;
    dim3 grid((unsigned)std::min<long long>((items + 255) / 256, 8192), std::min(batch, 65535));
    LRFB_LAUNCH(frontend8_rgb_kernel, grid, dim3(256), 0, st, (const unsigned char*)d_images, xs[0], g.fp);
    return check_launch("frontend8_rgb_kernel");
  }
  for (int pl = 0; pl < g.lay.n_planes; ++pl) {
    if (!((pl == 0 ? 1 : 2) & planes)) continue;
    long long per_img = (long long)g.lay.rows[pl] * g.lay.cols;
    int gx = (int)std::min<long long>((per_img + 255) / 256, 8192);
    dim3 grid(gx, std::min(batch, 65535));
    if (cfg->input_dtype == LRFB_U8)
      LRFB_LAUNCH(frontend_kernel<unsigned char>, grid, dim3(256), 0, st, (const unsigned char*)d_images,
                  xs[pl], g.fp, pl);
    else
      LRFB_LAUNCH(frontend_kernel<float>, grid, dim3(256), 0, st, (const float*)d_images, xs[pl], g.fp, pl);
    int rc = check_launch("frontend_kernel");
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

// =====================================================================================================

LRFB_EXPORT int32_t lrfb_abi_version(void) { return LRFB_ABI_VERSION; }
LRFB_EXPORT const char* lrfb_last_error(void) { return g_err; }

LRFB_EXPORT int32_t lrfb_device_count(void) {
#ifdef LRFB_SIM
  return 0;
#else
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail((int)e, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
#endif
}

LRFB_EXPORT int32_t lrfb_qmf_layout_query(const lrfb_qmf_config* cfg, lrfb_qmf_layout* out) {
  if (!out) return fail(LRFB_E_ARG, "out is null");
  Geometry g;
  int rc = make_geometry(cfg, 1, &g);
  if (rc) return rc;
  *out = g.lay;
  return 0;
}

LRFB_EXPORT int32_t lrfb_qmf_workspace_query(const lrfb_qmf_config* cfg, int32_t batch,
                                             lrfb_qmf_workspace_map* out) {
  if (!out || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  Geometry g;
  int rc = make_geometry(cfg, batch, &g);
  if (rc) return rc;
  make_map(g, batch, out);
  out->total_bytes += encode_scratch_bytes(g, batch);
  return 0;
}

LRFB_EXPORT int32_t lrfb_qmf_frontend(const lrfb_qmf_config* cfg, int32_t batch, const void* d_images,
                                      float* d_x, void* stream) {
  if (!d_images || !d_x || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  Geometry g;
  int rc = make_geometry(cfg, batch, &g);
  if (rc) return rc;
  lrfb_qmf_workspace_map m;
  make_map(g, batch, &m);
  float* xs[3] = {nullptr, nullptr, nullptr};
  for (int pl = 0; pl < g.lay.n_planes; ++pl)
    xs[pl] = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(d_x) + (m.x[pl] - m.x[0]));
  return run_frontend(cfg, g, batch, d_images, xs, (cudaStream_t)(uintptr_t)stream);
}

LRFB_EXPORT int32_t lrfb_qmf_encode(const lrfb_qmf_config* cfg, int32_t batch, const void* d_images,
                                    int8_t* d_factors, void* d_workspace, int64_t workspace_bytes,
                                    const lrfb_qmf_debug* dbg, void* stream) {
  if (!d_images || !d_factors || !d_workspace || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  Geometry g;
  int rc = make_geometry(cfg, batch, &g);
  if (rc) return rc;
  if ((rc = check_bounds(cfg->bound_lo, cfg->bound_hi))) return rc;
  lrfb_qmf_workspace_map m;
  make_map(g, batch, &m);
  const int64_t need = m.total_bytes + encode_scratch_bytes(g, batch);
  if (workspace_bytes < need)
    return fail(LRFB_E_WORKSPACE, "workspace %lld < required %lld bytes", (long long)workspace_bytes, (long long)need);
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  float* xs[3];
  for (int pl = 0; pl < 3; ++pl) xs[pl] = reinterpret_cast<float*>(ws + m.x[pl]);
  const lrfb_qmf_layout& L = g.lay;
  bool luma_gram_done = false;
  auto run_plane = [&](int pl, int phase, cudaStream_t s) {
    return factorize_batch(xs[pl], batch, L.rows[pl], L.cols, L.rank[pl], cfg->bound_lo, cfg->bound_hi,
                           cfg->num_iters, reinterpret_cast<float*>(ws + m.u[pl]),
                           reinterpret_cast<float*>(ws + m.v[pl]), d_factors + L.u_offset[pl],
                           d_factors + L.v_offset[pl], L.record_bytes, dbg ? dbg->d_init_u[pl] : nullptr,
                           dbg ? dbg->d_init_v[pl] : nullptr, dbg ? dbg->d_sign_flip[pl] : nullptr,
                           reinterpret_cast<double*>(ws + m.gram[pl]), reinterpret_cast<double*>(ws + m.evec[pl]),
                           reinterpret_cast<double*>(ws + m.sigma[pl]), ws + m.total_bytes,
                           dbg && dbg->stop_after == 2, s, phase, cfg->input_dtype == LRFB_U8,
                           pl == 0 && luma_gram_done,
                           cfg->input_dtype == LRFB_U8 && cfg->color_space == LRFB_RGB);
  };
  // The planes are independent.  When every plane runs the shared-memory-resident sweeps (no shared scratch),
  // the chroma work goes to a low-priority helper stream: the luma sweeps occupy 15 clusters x 8 SMs, the
  // chroma clusters fill the remaining SMs; with large batches (no Gram row split) the whole chroma chain
  // (front end, Gram, eigen-solver, sweeps) runs there so the latency-bound eigen-solver of one plane
  // overlaps the DMMA-bound Gram of another and the luma chain is the only critical path.
  bool overlap = false;
#ifndef LRFB_SIM
  overlap = L.n_planes == 3 && !(dbg && dbg->stop_after) && cfg->num_iters > 0 && !dev_getenv("LRFB_NO_OVERLAP");
  for (int pl = 0; pl < L.n_planes && overlap; ++pl)
    overlap = resident_ok(L.cols, L.rank[pl], L.rows[pl]) ||
              tc_resident_ok(L.cols, L.rank[pl], L.rows[pl], cfg->input_dtype == LRFB_U8);
  SideStream* side = overlap ? side_stream() : nullptr;
  SideUnlock side_guard{side};
  overlap = overlap && side;
  if (overlap) {
    bool chains = !dev_getenv("LRFB_NO_CHAIN_OVERLAP");
    for (int pl = 0; pl < 3 && chains; ++pl) chains = FactorWs::gram_split(batch, L.rows[pl]) == 1;
    if (chains) {
      // Three initialisation chains (Gram + eigen-solver: the latter latency-bound) run side by side; then the luma
      // sweeps take their 15 clusters x 8 SMs and the chroma sweeps (Cb, then Cr) the SMs that leaves free.
      if (fused8_geometry(cfg, g, d_images)) {  // the image is read once; the chroma chains fork after it
        if (frontgram_ok(cfg, g, batch) && !(dbg && dbg->d_init_u[0] && dbg->d_init_v[0])) {
          // front end + luma Gram in one pass: X_y is written but never read back for its Gram
          const size_t fsmem = sizeof(FrontGramSmem) + 1024;
          cudaError_t e = cudaFuncSetAttribute(frontgram8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
          if (e != cudaSuccess) return fail((int)e, "frontgram smem attribute: %s", cudaGetErrorString(e));
          frontgram8_kernel<<<dim3((unsigned)batch), dim3(kFgThreads), fsmem, st>>>(
              (const unsigned char*)d_images, xs[0], xs[1], xs[2], reinterpret_cast<double*>(ws + m.gram[0]), g.fp);
          if ((rc = check_launch("frontgram8_kernel"))) return rc;
          luma_gram_done = true;
        } else if ((rc = run_frontend(cfg, g, batch, d_images, xs, st, 3))) {
          return rc;
        }
        cudaEventRecord(side->fork, st);
        cudaStreamWaitEvent(side->stream, side->fork, 0);
      } else {
        cudaEventRecord(side->fork, st);  // orders the helper streams after whatever produced d_images
        cudaStreamWaitEvent(side->stream, side->fork, 0);
        if ((rc = run_frontend(cfg, g, batch, d_images, xs, st, 1))) return rc;
        if ((rc = run_frontend(cfg, g, batch, d_images, xs, side->stream, 2))) return rc;
        cudaEventRecord(side->fork, side->stream);
      }
      cudaStreamWaitEvent(side->stream2, side->fork, 0);
      static const bool timeline = dev_getenv("LRFB_TIMELINE") != nullptr;  // dev aid: when does each chain finish?
      cudaEvent_t ev[8] = {};
      if (timeline) {
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
      }
      if ((rc = run_plane(0, 1, st))) return rc;
      if (timeline) cudaEventRecord(ev[1], st);
      if ((rc = run_plane(1, 1, side->stream))) return rc;
      if (timeline) cudaEventRecord(ev[2], side->stream);
      if ((rc = run_plane(2, 1, side->stream2))) return rc;
      cudaEventRecord(side->init2, side->stream2);
      if (timeline) cudaEventRecord(ev[3], side->stream2);
      // the chroma sweeps become runnable together with the luma sweeps, not before: the (higher-priority) luma clusters
      // are placed first and the chroma clusters take the SMs they leave; later chroma clusters start as SMs free up
      cudaEventRecord(side->join, st);
      cudaStreamWaitEvent(side->stream, side->join, 0);
      if ((rc = run_plane(0, 2, st))) return rc;
      if (timeline) cudaEventRecord(ev[4], st);
      if ((rc = run_plane(1, 2, side->stream))) return rc;
      if (timeline) cudaEventRecord(ev[5], side->stream);
      cudaStreamWaitEvent(side->stream, side->init2, 0);
      if ((rc = run_plane(2, 2, side->stream))) return rc;
      if (timeline) {
        cudaEventRecord(ev[6], side->stream);
        cudaStreamSynchronize(side->stream);
        cudaStreamSynchronize(st);
        float t[7] = {};
        for (int i = 1; i < 7; ++i) cudaEventElapsedTime(&t[i], ev[0], ev[i]);
        fprintf(stderr, "[lrfb timeline, ms after the front end] init: luma %.2f cb %.2f cr %.2f | sweeps: luma %.2f cb %.2f cr %.2f\n",
                t[1], t[2], t[3], t[4], t[5], t[6]);
        for (auto& e : ev) cudaEventDestroy(e);
      }
    } else {
      if ((rc = run_frontend(cfg, g, batch, d_images, xs, st))) return rc;
      for (int pl = 0; pl < 3; ++pl)
        if ((rc = run_plane(pl, 1, st))) return rc;
      cudaEventRecord(side->fork, st);
      cudaStreamWaitEvent(side->stream, side->fork, 0);
      if ((rc = run_plane(0, 2, st))) return rc;
      if ((rc = run_plane(1, 2, side->stream))) return rc;
      if ((rc = run_plane(2, 2, side->stream))) return rc;
    }
    cudaEventRecord(side->join, side->stream);
    cudaStreamWaitEvent(st, side->join, 0);
    return 0;
  }
#endif
  if ((rc = run_frontend(cfg, g, batch, d_images, xs, st))) return rc;
  if (dbg && dbg->stop_after == 1) return 0;
  for (int pl = 0; pl < L.n_planes; ++pl)
    if ((rc = run_plane(pl, 0, st))) return rc;
  return 0;
}

LRFB_EXPORT int64_t lrfb_factorize_workspace_bytes(int32_t n_mat, int32_t M, int32_t N, int32_t R) {
  if (n_mat <= 0 || M <= 0 || N <= 0 || R <= 0) return 0;
  return align_up((int64_t)n_mat * N * N * 8, 256) + align_up((int64_t)n_mat * N * R * 8, 256) +
         align_up((int64_t)n_mat * R * 12 + 16, 256) + FactorWs::bytes(n_mat, M, N, R);
}

LRFB_EXPORT int32_t lrfb_factorize(const float* d_x, int32_t n_mat, int32_t M, int32_t N, int32_t R,
                                   float bound_lo, float bound_hi, int32_t num_iters, float* d_u, float* d_v,
                                   const float* d_init_u, const float* d_init_v, const int32_t* d_sign_flip,
                                   void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (!d_x || !d_u || !d_v || !d_workspace || n_mat <= 0 || M <= 0 || N <= 0 || R <= 0)
    return fail(LRFB_E_ARG, "bad arguments");
  if (R > kGenMaxR || N > 1024) return fail(LRFB_E_UNSUPPORTED, "N=%d R=%d not implemented", N, R);
  int rc = check_bounds(bound_lo, bound_hi);
  if (rc) return rc;
  if (workspace_bytes < lrfb_factorize_workspace_bytes(n_mat, M, N, R))
    return fail(LRFB_E_WORKSPACE, "workspace too small");
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  double* gram = reinterpret_cast<double*>(ws);
  ws += align_up((int64_t)n_mat * N * N * 8, 256);
  double* evec = reinterpret_cast<double*>(ws);
  ws += align_up((int64_t)n_mat * N * R * 8, 256);
  double* sigma = reinterpret_cast<double*>(ws);
  ws += align_up((int64_t)n_mat * R * 12 + 16, 256);
  return factorize_batch(d_x, n_mat, M, N, R, bound_lo, bound_hi, num_iters, d_u, d_v, nullptr, nullptr, 0,
                         d_init_u, d_init_v, d_sign_flip, gram, evec, sigma, ws, 0,
                         (cudaStream_t)(uintptr_t)stream);
}

LRFB_EXPORT int64_t lrfb_launch_count(void) { return g_launches; }

LRFB_EXPORT int32_t lrfb_bcd(const float* d_x, int32_t n_mat, int32_t M, int32_t N, int32_t R, float bound_lo,
                             float bound_hi, int32_t num_iters, float* d_u, float* d_v, const float* d_s0,
                             uint32_t flags, void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (!d_x || !d_u || !d_v || n_mat <= 0 || M <= 0 || N <= 0 || R <= 0 || num_iters <= 0)
    return fail(LRFB_E_ARG, "bad arguments");
  if (R > kGenMaxR || N > 1024) return fail(LRFB_E_UNSUPPORTED, "N=%d R=%d not implemented", N, R);
  int rc = check_bounds(bound_lo, bound_hi);
  if (rc) return rc;
  if (workspace_bytes < (int64_t)kGenGrid * gen_scratch_floats(N, R) * 4 || !d_workspace)
    return fail(LRFB_E_WORKSPACE, "workspace too small");
  BcdBatch b;
  b.X = d_x, b.x_stride = (long long)M * N, b.U = d_u, b.V = d_v, b.Uq = nullptr, b.Vq = nullptr;
  b.uq_stride = b.vq_stride = 0;
  b.s0 = d_s0;
  b.x_u8_range = (flags & 1u) ? 1 : 0;
  b.M = M, b.n_mat = n_mat, b.num_iters = num_iters, b.lo = ceilf(bound_lo), b.hi = floorf(bound_hi);
  // the generic kernel's scratch is unused whenever the counter is (tensor-core path): they may share the buffer
  return run_bcd(b, N, R, reinterpret_cast<float*>(d_workspace), (cudaStream_t)(uintptr_t)stream,
                 reinterpret_cast<int*>(d_workspace));
}

LRFB_EXPORT int32_t lrfb_ffma_probe(float* d_out, int32_t iters, void* stream) {
  if (!d_out || iters <= 0) return fail(LRFB_E_ARG, "bad arguments");
  LRFB_LAUNCH(ffma_probe_kernel, dim3(num_sms() * 8), dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_out, iters);
  return check_launch("ffma_probe_kernel");
}

LRFB_EXPORT int32_t lrfb_qmf_decode(const lrfb_qmf_config* cfg, int32_t batch, const int8_t* d_factors,
                                    uint8_t* d_images, void* stream) {
  if (!d_factors || !d_images || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  Geometry g;
  int rc = make_geometry(cfg, batch, &g);
  if (rc) return rc;
  DecodeParams P;
  memset(&P, 0, sizeof(P));
  P.H = cfg->height, P.W = cfg->width, P.p = cfg->patch_h, P.q = cfg->patch_w;
  P.ycbcr = cfg->color_space == LRFB_YCBCR, P.n_img = batch, P.record_bytes = g.lay.record_bytes;
  for (int pl = 0; pl < 3; ++pl) {
    P.g[pl] = g.fp.g[pl], P.rank[pl] = g.lay.rank[pl];
    P.u_off[pl] = g.lay.u_offset[pl], P.v_off[pl] = g.lay.v_offset[pl];
  }
  long long hw = (long long)cfg->height * cfg->width;
  dim3 grid((unsigned)std::min<long long>((hw + 255) / 256, 8192), std::min(batch, 65535));
  if (fast8_geometry(g, d_images) && g.lay.record_bytes % 8 == 0 && ((uintptr_t)d_factors % 8) == 0 &&
      g.lay.v_offset[0] % 8 == 0 && g.lay.v_offset[1] % 4 == 0 && g.lay.v_offset[2] % 4 == 0) {
    long long items = hw / 8;
#ifndef LRFB_SIM
    if (fused8_geometry(cfg, g, d_images) && g.lay.rank[0] <= 4 && g.lay.rank[1] <= 4 && g.lay.rank[2] <= 4 &&
        !g_decode_v1.load(std::memory_order_relaxed)) {
      dim3 grid2((unsigned)std::min<long long>((items / 8 + 255) / 256, 4096), std::min(batch, 65535));  // one thread per luma patch
      LRFB_LAUNCH(qmf_decode8x2_kernel, grid2, dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_factors, d_images, P);
      return check_launch("qmf_decode8x2_kernel");
    }
#endif
    dim3 grid8((unsigned)std::min<long long>((items + 255) / 256, 4096), std::min(batch, 65535));
    LRFB_LAUNCH(qmf_decode8_kernel, grid8, dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_factors, d_images, P);
    return check_launch("qmf_decode8_kernel");
  }
  LRFB_LAUNCH(qmf_decode_kernel, grid, dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_factors, d_images, P);
  return check_launch("qmf_decode_kernel");
}

LRFB_EXPORT int32_t lrfb_qmf_decode_planes(int32_t color_space, int32_t height, int32_t width, int32_t chroma_h,
                                           int32_t chroma_w, const int32_t* rank, int32_t batch,
                                           const int8_t* const* d_u, const int8_t* const* d_v, uint8_t* d_images,
                                           void* stream) {
  if (!rank || !d_u || !d_v || !d_images || batch <= 0 || height <= 0 || width <= 0)
    return fail(LRFB_E_ARG, "bad arguments");
  if (color_space != LRFB_RGB && color_space != LRFB_YCBCR) return fail(LRFB_E_ARG, "bad color_space");
  PlanesParams P;
  memset(&P, 0, sizeof(P));
  P.H = height, P.W = width, P.ch = chroma_h, P.cw = chroma_w, P.ycbcr = color_space == LRFB_YCBCR, P.n_img = batch;
  const int n_pl = P.ycbcr ? 3 : 1;
  for (int pl = 0; pl < n_pl; ++pl) {
    if (!d_u[pl] || !d_v[pl] || rank[pl] <= 0) return fail(LRFB_E_ARG, "plane %d: null factor or bad rank", pl);
    P.u[pl] = d_u[pl], P.v[pl] = d_v[pl], P.rank[pl] = rank[pl];
  }
  if (P.ycbcr && (chroma_h <= 0 || chroma_w <= 0)) return fail(LRFB_E_ARG, "bad chroma size");
  const long long hw = (long long)height * width;
  dim3 grid((unsigned)std::min<long long>((hw + 255) / 256, 8192), std::min(batch, 65535));
  LRFB_LAUNCH(qmf_decode_planes_kernel, grid, dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_images, P);
  return check_launch("qmf_decode_planes_kernel");
}

LRFB_EXPORT int32_t lrfb_svd_encode(const lrfb_qmf_config* cfg, int32_t batch, const void* d_images,
                                    uint8_t* d_codes, float* d_qparams, void* d_workspace, int64_t workspace_bytes,
                                    const lrfb_qmf_debug* dbg, void* stream) {
  if (!d_images || !d_codes || !d_qparams || !d_workspace || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  if (!cfg || cfg->color_space != LRFB_RGB)
    return fail(LRFB_E_UNSUPPORTED, "svd codec: only color_space RGB is implemented (the reference's YCbCr branch is broken)");
  Geometry g;
  int rc = make_geometry(cfg, batch, &g);
  if (rc) return rc;
  lrfb_qmf_workspace_map m;
  make_map(g, batch, &m);
  const int64_t need = m.total_bytes + encode_scratch_bytes(g, batch);
  if (workspace_bytes < need) return fail(LRFB_E_WORKSPACE, "workspace %lld < required %lld bytes", (long long)workspace_bytes, (long long)need);
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  float* xs[3] = {reinterpret_cast<float*>(ws + m.x[0]), nullptr, nullptr};
  if ((rc = run_frontend(cfg, g, batch, d_images, xs, st))) return rc;
  const lrfb_qmf_layout& L = g.lay;
  const int M = L.rows[0], N = L.cols, R = L.rank[0];
  float* u = reinterpret_cast<float*>(ws + m.u[0]);
  float* v = reinterpret_cast<float*>(ws + m.v[0]);
  double* evec = reinterpret_cast<double*>(ws + m.evec[0]);
  double* sigma = reinterpret_cast<double*>(ws + m.sigma[0]);
  float* mm = reinterpret_cast<float*>(ws + m.total_bytes);  // [2][batch][2]
  unsigned char* scratch = ws + m.total_bytes + align_up((int64_t)batch * 16, 256);
  rc = factorize_batch(xs[0], batch, M, N, R, -1.f, 1.f, 1, u, v, nullptr, nullptr, 0, nullptr, nullptr,
                       dbg ? dbg->d_sign_flip[0] : nullptr, reinterpret_cast<double*>(ws + m.gram[0]), evec, sigma,
                       scratch, 0, st, 1, false, false, cfg->input_dtype == LRFB_U8);
  if (rc) return rc;
  for (int m0 = 0; m0 < batch; m0 += 65535) {
    const int cnt = std::min(65535, batch - m0);
    const int gx = std::max(1, std::min((M + kProjRows - 1) / kProjRows, 256));
    const size_t psmem = (size_t)N * kProjCols * 8;
    const bool vec = N % 4 == 0;
#ifndef LRFB_SIM
    if (psmem > 48 * 1024) {
      cudaError_t e = vec ? cudaFuncSetAttribute(svd_project_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem)
                          : cudaFuncSetAttribute(svd_project_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
      if (e != cudaSuccess) return fail((int)e, "project smem attribute: %s", cudaGetErrorString(e));
    }
#endif
    if (vec)
      LRFB_LAUNCH(svd_project_kernel<true>, dim3(gx, cnt), dim3(kProjRows), psmem, st, xs[0] + (size_t)m0 * M * N,
                  (long long)M * N, M, N, R, evec + (size_t)m0 * N * R, sigma + (size_t)m0 * R, u + (size_t)m0 * M * R);
    else
      LRFB_LAUNCH(svd_project_kernel<false>, dim3(gx, cnt), dim3(kProjRows), psmem, st, xs[0] + (size_t)m0 * M * N,
                  (long long)M * N, M, N, R, evec + (size_t)m0 * N * R, sigma + (size_t)m0 * R, u + (size_t)m0 * M * R);
    if ((rc = check_launch("svd_project_kernel"))) return rc;
  }
  LRFB_LAUNCH(minmax_kernel, dim3(batch), dim3(256), 0, st, u, (long long)M * R, mm);
  if ((rc = check_launch("minmax_kernel"))) return rc;
  LRFB_LAUNCH(minmax_kernel, dim3(batch), dim3(256), 0, st, v, (long long)N * R, mm + 2 * (size_t)batch);
  if ((rc = check_launch("minmax_kernel"))) return rc;
  for (int m0 = 0; m0 < batch; m0 += 65535) {
    const int cnt = std::min(65535, batch - m0);
    LRFB_LAUNCH(quantize_u8_kernel, dim3(std::min((M * R + 255) / 256, 256), cnt), dim3(256), 0, st,
                u + (size_t)m0 * M * R, M, R, mm + 2 * (size_t)m0, d_codes + (size_t)m0 * L.record_bytes + L.u_offset[0],
                (long long)L.record_bytes, d_qparams + 4 * (size_t)m0, 4LL);
    if ((rc = check_launch("quantize_u8_kernel"))) return rc;
    LRFB_LAUNCH(quantize_u8_kernel, dim3(std::min((N * R + 255) / 256, 256), cnt), dim3(256), 0, st,
                v + (size_t)m0 * N * R, N, R, mm + 2 * (size_t)batch + 2 * (size_t)m0,
                d_codes + (size_t)m0 * L.record_bytes + L.v_offset[0], (long long)L.record_bytes,
                d_qparams + 4 * (size_t)m0 + 2, 4LL);
    if ((rc = check_launch("quantize_u8_kernel"))) return rc;
  }
  return 0;
}

LRFB_EXPORT int32_t lrfb_svd_decode(const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* d_codes,
                                    const float* d_qparams6, uint8_t* d_images, void* stream) {
  if (!d_codes || !d_qparams6 || !d_images || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  if (!cfg || cfg->color_space != LRFB_RGB) return fail(LRFB_E_UNSUPPORTED, "svd codec: only color_space RGB");
  Geometry g;
  int rc = make_geometry(cfg, batch, &g);
  if (rc) return rc;
  SvdDecodeParams P;
  memset(&P, 0, sizeof(P));
  P.H = cfg->height, P.W = cfg->width, P.p = cfg->patch_h, P.q = cfg->patch_w, P.n_img = batch, P.R = g.lay.rank[0];
  P.g = g.fp.g[0], P.record_bytes = g.lay.record_bytes, P.u_off = g.lay.u_offset[0], P.v_off = g.lay.v_offset[0];
  long long hw = (long long)cfg->height * cfg->width;
  dim3 grid((unsigned)std::min<long long>((hw + 255) / 256, 8192), std::min(batch, 65535));
  LRFB_LAUNCH(svd_decode_kernel, grid, dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_codes, d_qparams6, d_images, P);
  return check_launch("svd_decode_kernel");
}

LRFB_EXPORT int32_t lrfb_sse_u8(const uint8_t* d_a, const uint8_t* d_b, int64_t elems_per_image, int32_t batch,
                                uint64_t* d_sse, void* stream) {
  if (!d_a || !d_b || !d_sse || batch <= 0 || elems_per_image <= 0) return fail(LRFB_E_ARG, "bad arguments");
  dim3 grid((unsigned)std::min<long long>((elems_per_image + 256 * 16 - 1) / (256 * 16), 1024),
            std::min(batch, 65535));
  LRFB_LAUNCH(sse_u8_kernel, grid, dim3(256), 0, (cudaStream_t)(uintptr_t)stream, d_a, d_b,
              (long long)elems_per_image, batch, reinterpret_cast<unsigned long long*>(d_sse));
  return check_launch("sse_u8_kernel");
}

// ---- host-buffer path ----------------------------------------------------------------------------

constexpr int kHostSlots = 3;  // input chunks in flight: one being consumed by the kernels, two on the wire
struct lrfb_ctx {
  int device;
  cudaStream_t stream;   // compute + D2H
  cudaStream_t copy;     // H2D
  void* d_in;            // kHostSlots chunk-sized input buffers back to back
  void* d_out;           // two chunk-sized record buffers (the D2H of chunk i overlaps the kernels of chunk i+1)
  void* d_ws;
  size_t in_cap, out_cap, ws_cap;
  size_t chunk_bytes;    // input bytes per pipeline chunk
  // lrfb_qmf_encode_bytes_host: two blob slots + their offset tables, the packer's workspace, pinned offset staging
  void* d_blob;
  void* d_offs;
  void* d_pws;
  long long* h_offs;
  size_t blob_cap, offs_cap, pws_cap, hoffs_cap;
#ifndef LRFB_SIM
  cudaStream_t back;     // D2H
  cudaStream_t pack[2];  // lossless stage, one stream per output slot: two chunks' deflate launches share the GPU with the next encode
  cudaEvent_t landed[kHostSlots], consumed[kHostSlots], encoded[2], drained[2], packed[2], sized[2];
#endif
};

#ifdef LRFB_SIM
namespace {
int grow(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return 0;
  free(*p);
  *p = malloc(need);
  *cap = need;
  return *p ? 0 : fail(2, "out of memory");
}
int h2d(void* d, const void* h, size_t n, cudaStream_t) { memcpy(d, h, n); return 0; }
int d2h(void* h, const void* d, size_t n, cudaStream_t) { memcpy(h, d, n); return 0; }
int sync_stream(cudaStream_t) { return 0; }
}  // namespace
LRFB_EXPORT int32_t lrfb_ctx_create(int32_t device, lrfb_ctx** out) {
  if (!out) return fail(LRFB_E_ARG, "out is null");
  *out = new lrfb_ctx();
  memset(*out, 0, sizeof(lrfb_ctx));
  (*out)->device = device;
  (*out)->chunk_bytes = (size_t)256 << 20;
  return 0;
}
LRFB_EXPORT void lrfb_ctx_destroy(lrfb_ctx* c) {
  if (!c) return;
  free(c->d_in), free(c->d_out), free(c->d_ws), free(c->d_blob), free(c->d_offs), free(c->d_pws), free(c->h_offs);
  delete c;
}
#else
namespace {
int grow(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr, *cap = 0;
  cudaError_t e = cudaMalloc(p, need);
  if (e != cudaSuccess) return fail((int)e, "cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
  *cap = need;
  return 0;
}
int h2d(void* d, const void* h, size_t n, cudaStream_t st) {
  cudaError_t e = cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st);
  return e == cudaSuccess ? 0 : fail((int)e, "H2D: %s", cudaGetErrorString(e));
}
int d2h(void* h, const void* d, size_t n, cudaStream_t st) {
  cudaError_t e = cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st);
  return e == cudaSuccess ? 0 : fail((int)e, "D2H: %s", cudaGetErrorString(e));
}
int sync_stream(cudaStream_t st) {
  cudaError_t e = cudaStreamSynchronize(st);
  return e == cudaSuccess ? 0 : fail((int)e, "stream sync: %s", cudaGetErrorString(e));
}
}  // namespace
LRFB_EXPORT int32_t lrfb_ctx_create(int32_t device, lrfb_ctx** out) {
  if (!out) return fail(LRFB_E_ARG, "out is null");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail((int)e, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  lrfb_ctx* c = new lrfb_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->chunk_bytes = (size_t)256 << 20;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->back, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->pack[0], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->pack[1], cudaStreamNonBlocking);
  for (int i = 0; i < kHostSlots && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&c->landed[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->consumed[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&c->encoded[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->drained[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->packed[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->sized[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    delete c;
    return fail((int)e, "lrfb_ctx_create: %s", cudaGetErrorString(e));
  }
  *out = c;
  return 0;
}
LRFB_EXPORT void lrfb_ctx_destroy(lrfb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->d_in) cudaFree(c->d_in);
  if (c->d_out) cudaFree(c->d_out);
  if (c->d_ws) cudaFree(c->d_ws);
  if (c->d_blob) cudaFree(c->d_blob);
  if (c->d_offs) cudaFree(c->d_offs);
  if (c->d_pws) cudaFree(c->d_pws);
  if (c->h_offs) cudaFreeHost(c->h_offs);
  for (int i = 0; i < kHostSlots; ++i) cudaEventDestroy(c->landed[i]), cudaEventDestroy(c->consumed[i]);
  for (int i = 0; i < 2; ++i)
    cudaEventDestroy(c->encoded[i]), cudaEventDestroy(c->drained[i]), cudaEventDestroy(c->packed[i]), cudaEventDestroy(c->sized[i]);
  cudaStreamDestroy(c->pack[0]);
  cudaStreamDestroy(c->pack[1]);
  cudaStreamDestroy(c->back);
  cudaStreamDestroy(c->copy);
  cudaStreamDestroy(c->stream);
  delete c;
}
#endif

LRFB_EXPORT int32_t lrfb_ctx_set_chunk_bytes(lrfb_ctx* c, int64_t bytes) {
  if (!c || bytes <= 0) return fail(LRFB_E_ARG, "bad arguments");
  c->chunk_bytes = (size_t)bytes;
  return 0;
}

// Chunked pipeline over three streams: H2D of chunks i+1, i+2 (copy stream) | kernels of chunk i (compute stream) |
// D2H of the int8 records of chunk i-1 (back stream).  The workspace is sized for one chunk, so host batches larger
// than device memory would allow still encode.
LRFB_EXPORT int32_t lrfb_qmf_encode_host(lrfb_ctx* c, const lrfb_qmf_config* cfg, int32_t batch,
                                         const void* h_images, int8_t* h_factors) {
  if (!c || !h_images || !h_factors || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  lrfb_qmf_workspace_map m, mt;
  lrfb_qmf_layout L;
  int rc;
  if ((rc = lrfb_qmf_layout_query(cfg, &L))) return rc;
  const size_t img_bytes = (size_t)3 * cfg->height * cfg->width * (cfg->input_dtype == LRFB_U8 ? 1 : 4);
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)batch, c->chunk_bytes / std::max<size_t>(img_bytes, 1)));
  // the scratch need is not monotonic in the batch (the Gram row split grows as the batch shrinks): size the workspace
  // for the full chunk AND for the ragged tail
  if ((rc = lrfb_qmf_workspace_query(cfg, chunk, &m))) return rc;
  int64_t ws_need = m.total_bytes;
  if (batch % chunk) {
    if ((rc = lrfb_qmf_workspace_query(cfg, batch % chunk, &mt))) return rc;
    ws_need = std::max(ws_need, mt.total_bytes);
  }
#ifndef LRFB_SIM
  cudaSetDevice(c->device);
#endif
  if ((rc = grow(&c->d_in, &c->in_cap, (size_t)kHostSlots * chunk * img_bytes))) return rc;
  if ((rc = grow(&c->d_out, &c->out_cap, (size_t)2 * chunk * L.record_bytes))) return rc;
  if ((rc = grow(&c->d_ws, &c->ws_cap, (size_t)ws_need))) return rc;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(h_images);
  int idx = 0;
  for (int i0 = 0; i0 < batch; i0 += chunk, ++idx) {
    const int n = std::min(chunk, batch - i0);
    const int slot = idx % kHostSlots, oslot = idx & 1;
    unsigned char* d_in = reinterpret_cast<unsigned char*>(c->d_in) + (size_t)slot * chunk * img_bytes;
    int8_t* d_out = reinterpret_cast<int8_t*>(c->d_out) + (size_t)oslot * chunk * L.record_bytes;
#ifndef LRFB_SIM
    if (idx >= kHostSlots) cudaStreamWaitEvent(c->copy, c->consumed[slot], 0);  // kernels of chunk idx-3 are done with it
    if ((rc = h2d(d_in, src + (size_t)i0 * img_bytes, (size_t)n * img_bytes, c->copy))) return rc;
    cudaEventRecord(c->landed[slot], c->copy);
    cudaStreamWaitEvent(c->stream, c->landed[slot], 0);
    if (idx >= 2) cudaStreamWaitEvent(c->stream, c->drained[oslot], 0);  // records of chunk idx-2 have left d_out
#else
    if ((rc = h2d(d_in, src + (size_t)i0 * img_bytes, (size_t)n * img_bytes, c->stream))) return rc;
#endif
    if ((rc = lrfb_qmf_encode(cfg, n, d_in, d_out, c->d_ws, (int64_t)c->ws_cap, nullptr, (void*)(uintptr_t)c->stream)))
      return rc;
#ifndef LRFB_SIM
    cudaEventRecord(c->consumed[slot], c->stream);
    cudaEventRecord(c->encoded[oslot], c->stream);
    cudaStreamWaitEvent(c->back, c->encoded[oslot], 0);
    if ((rc = d2h(h_factors + (size_t)i0 * L.record_bytes, d_out, (size_t)n * L.record_bytes, c->back))) return rc;
    cudaEventRecord(c->drained[oslot], c->back);
#else
    if ((rc = d2h(h_factors + (size_t)i0 * L.record_bytes, d_out, (size_t)n * L.record_bytes, c->stream))) return rc;
#endif
  }
#ifndef LRFB_SIM
  if ((rc = sync_stream(c->back))) return rc;
#endif
  return sync_stream(c->stream);
}

LRFB_EXPORT int32_t lrfb_qmf_decode_host(lrfb_ctx* c, const lrfb_qmf_config* cfg, int32_t batch,
                                         const int8_t* h_factors, uint8_t* h_images) {
  if (!c || !h_images || !h_factors || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  lrfb_qmf_layout L;
  int rc;
  if ((rc = lrfb_qmf_layout_query(cfg, &L))) return rc;
#ifndef LRFB_SIM
  cudaSetDevice(c->device);
#endif
  const size_t img_bytes = (size_t)3 * cfg->height * cfg->width;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)batch, c->chunk_bytes / std::max<size_t>(img_bytes, 1)));
  // d_in doubles as the record buffer (whole batch: records are ~40x smaller than images), d_out holds two image chunks
  if ((rc = grow(&c->d_in, &c->in_cap, (size_t)batch * L.record_bytes))) return rc;
  if ((rc = grow(&c->d_out, &c->out_cap, (size_t)2 * chunk * img_bytes))) return rc;
  int idx = 0;
  for (int i0 = 0; i0 < batch; i0 += chunk, ++idx) {
    const int n = std::min(chunk, batch - i0);
    const int oslot = idx & 1;
    uint8_t* d_img = reinterpret_cast<uint8_t*>(c->d_out) + (size_t)oslot * chunk * img_bytes;
    // the records go up chunk by chunk too: the copy-back, the bound of this call, starts after one chunk's worth
    if ((rc = h2d(reinterpret_cast<int8_t*>(c->d_in) + (size_t)i0 * L.record_bytes, h_factors + (size_t)i0 * L.record_bytes,
                  (size_t)n * L.record_bytes, c->stream)))
      return rc;
#ifndef LRFB_SIM
    if (idx >= 2) cudaStreamWaitEvent(c->stream, c->drained[oslot], 0);
#endif
    if ((rc = lrfb_qmf_decode(cfg, n, reinterpret_cast<const int8_t*>(c->d_in) + (size_t)i0 * L.record_bytes, d_img,
                              (void*)(uintptr_t)c->stream)))
      return rc;
#ifndef LRFB_SIM
    cudaEventRecord(c->encoded[oslot], c->stream);
    cudaStreamWaitEvent(c->back, c->encoded[oslot], 0);
    if ((rc = d2h(h_images + (size_t)i0 * img_bytes, d_img, (size_t)n * img_bytes, c->back))) return rc;
    cudaEventRecord(c->drained[oslot], c->back);
#else
    if ((rc = d2h(h_images + (size_t)i0 * img_bytes, d_img, (size_t)n * img_bytes, c->stream))) return rc;
#endif
  }
#ifndef LRFB_SIM
  if ((rc = sync_stream(c->back))) return rc;
#endif
  return sync_stream(c->stream);
}

// ---- lossless packing on host threads (the reference's encode_tensor / combine_bytes, byte for byte) ----------------

namespace {
typedef std::vector<unsigned char> Bytes;
void put_be32(Bytes& o, size_t v) {
  o.push_back((unsigned char)(v >> 24)), o.push_back((unsigned char)(v >> 16));
  o.push_back((unsigned char)(v >> 8)), o.push_back((unsigned char)v);
}
// combine_bytes(parts) = left fold of combine(a, b) = BE32(len a) | a | b  (lrf/compression/utils.py:246-300): the
// result is the k-1 nested length prefixes (outermost first) followed by the parts themselves
void combine_into(Bytes& out, const std::vector<const Bytes*>& parts) {
  std::vector<size_t> acc(parts.size());
  size_t run = parts[0]->size();
  acc[0] = run;
  for (size_t i = 1; i < parts.size(); ++i) acc[i] = run = 4 + run + parts[i]->size();
  for (size_t i = parts.size() - 1; i >= 1; --i) put_be32(out, acc[i - 1]);
  for (const Bytes* p : parts) out.insert(out.end(), p->begin(), p->end());
}
// One deflate state per worker thread, reset per column: what zlib.compress(data, 9) / compress2 produce (deflateInit,
// one deflate(Z_FINISH), deflateEnd) without re-allocating and re-zeroing ~270 KB of state for every 64-byte V column.
struct Deflater {
  z_stream zs;
  bool ok;
  Deflater() : ok(false) {
    memset(&zs, 0, sizeof(zs));
    ok = deflateInit(&zs, 9) == Z_OK;
  }
  ~Deflater() {
    if (ok) deflateEnd(&zs);
  }
  int run(Bytes& out, const unsigned char* src, size_t n) {
    if (!ok) return Z_MEM_ERROR;
    if (deflateReset(&zs) != Z_OK) return Z_STREAM_ERROR;
    out.resize(compressBound((uLong)n));
    zs.next_in = const_cast<Bytef*>(src), zs.avail_in = (uInt)n;
    zs.next_out = out.data(), zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) return rc == Z_OK ? Z_BUF_ERROR : rc;
    out.resize(zs.total_out);
    return Z_OK;
  }
};

// encode_matrix (lrf/compression/utils.py:354-390) of one fiber-major factor: R columns of `rows` int8 each
int encode_fibers(Bytes& out, const int8_t* fibers, int R, int rows, const char* dtype_name, std::vector<Bytes>& cols,
                  Deflater& z) {
  cols.resize(R);
  for (int r = 0; r < R; ++r) {
    const int zr = z.run(cols[r], reinterpret_cast<const unsigned char*>(fibers) + (size_t)r * rows, (size_t)rows);
    if (zr != Z_OK) return zr;
  }
  char hdr[96];
  int n = snprintf(hdr, sizeof(hdr), "{\"num_fibers\": %d, \"mode\": \"col\", \"dtype\": \"%s\"}", R, dtype_name);
  Bytes meta(hdr, hdr + n), body;
  std::vector<const Bytes*> cp;
  for (auto& cb : cols) cp.push_back(&cb);
  combine_into(body, cp);
  combine_into(out, {&meta, &body});
  return 0;
}
}  // namespace

LRFB_EXPORT int64_t lrfb_qmf_pack_bound(const lrfb_qmf_config* cfg, int64_t metadata_len) {
  lrfb_qmf_layout L;
  if (lrfb_qmf_layout_query(cfg, &L) || metadata_len < 0) return -1;
  int64_t b = 8 + metadata_len;
  for (int pl = 0; pl < L.n_planes; ++pl)
    b += 2 * (128 + 8) + (int64_t)L.rank[pl] * (8 + (int64_t)compressBound((uLong)L.rows[pl]) + (int64_t)compressBound((uLong)L.cols));
  return b;
}

LRFB_EXPORT int32_t lrfb_qmf_pack_host(const lrfb_qmf_config* cfg, int32_t batch, const int8_t* h_records,
                                       const char* metadata_json, int64_t metadata_len, uint8_t* h_out,
                                       int64_t out_stride, int64_t* out_sizes, int32_t threads) {
  if (!h_records || !metadata_json || !h_out || !out_sizes || batch <= 0 || metadata_len <= 0)
    return fail(LRFB_E_ARG, "bad arguments");
  lrfb_qmf_layout L;
  int rc = lrfb_qmf_layout_query(cfg, &L);
  if (rc) return rc;
  if (out_stride < lrfb_qmf_pack_bound(cfg, metadata_len)) return fail(LRFB_E_WORKSPACE, "out_stride below lrfb_qmf_pack_bound");
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, (int)batch));
  std::atomic<int> next{0}, err{0};
  auto work = [&]() {
    std::vector<Bytes> cols;
    Deflater z;
    Bytes meta(metadata_json, metadata_json + metadata_len), body, img;
    std::vector<Bytes> enc(2 * L.n_planes);
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= batch || err.load()) break;
      const int8_t* rec = h_records + (size_t)i * L.record_bytes;
      for (int pl = 0; pl < L.n_planes; ++pl) {
        enc[2 * pl].clear(), enc[2 * pl + 1].clear();
        int zr = encode_fibers(enc[2 * pl], rec + L.u_offset[pl], L.rank[pl], L.rows[pl], "int8", cols, z);
        if (!zr) zr = encode_fibers(enc[2 * pl + 1], rec + L.v_offset[pl], L.rank[pl], L.cols, "int8", cols, z);
        if (zr) err.store(zr);
      }
      body.clear(), img.clear();
      std::vector<const Bytes*> ep;
      for (auto& e : enc) ep.push_back(&e);
      combine_into(body, ep);
      combine_into(img, {&meta, &body});
      if ((int64_t)img.size() > out_stride) {
        err.store(-100);
        break;
      }
      memcpy(h_out + (size_t)i * out_stride, img.data(), img.size());
      out_sizes[i] = (int64_t)img.size();
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  if (err.load()) return fail(LRFB_E_UNSUPPORTED, "zlib / packing failed (%d)", err.load());
  return 0;
}

// ---- lossless stage on the device: every factor column deflated by one warp (deflate9.cuh), then framed ----------------

namespace {
struct PackPlan {
  lrfb_qmf_layout L;
  int n_mat, cols_total;
  int len[6], ncols[6], col0[6], slot[6], slot_off[6];
  long long rec_off[6];
  long long img_stride;     // column-buffer bytes per image
  int n_groups;
  int group_len[6];         // distinct column lengths, longest first
  int grid[6];
  int nw[6];               // warps per column: 1, kWarps (few long columns) or kWarpsHuge (shared memory holds < 4 per SM)
  long long off_cbuf, off_csize, off_sizes, off_counter, off_scratch[6], total;  // one scratch region per launch: they overlap in time
};
int ctas_per_sm(int len, int nw) {
  const int per = d9::smem_bytes(len) + 1024;  // 1 KB of system-reserved shared memory per CTA
  // the kernel's __launch_bounds__: 32 single-warp CTAs, 8 of kWarps warps, 2 of kWarpsHuge
  return std::max(1, std::min(nw == 1 ? 32 : nw <= d9::kWarps ? 8 : 2, (227 * 1024) / per));
}
// One warp per column gives the most columns in flight and the best throughput (measured on 4096 images of 768x512:
// 39.7 ms against 51.8 ms with four warps per luma column).  Four warps per column shorten a column's critical path
// (the long chain walks are shared), which is what counts when there are fewer long columns than single-warp slots:
// one image, small batches.  Columns so long that shared memory holds fewer than four per SM (CLIC-sized luma columns:
// one) get eight warps: 256 CLIC-sized records 196 ms with one warp, 108 ms with four.
int warps_per_column(int len, long long streams) {
  if (len <= d9::kLongColumn) return 1;
  if (ctas_per_sm(len, 1) < 4) return d9::kWarpsHuge;
  return streams <= (long long)num_sms() * ctas_per_sm(len, 1) ? d9::kWarps : 1;
}
int make_pack_plan(const lrfb_qmf_config* cfg, int batch, PackPlan& P) {
  int rc = lrfb_qmf_layout_query(cfg, &P.L);
  if (rc) return rc;
  const lrfb_qmf_layout& L = P.L;
  P.n_mat = 2 * L.n_planes, P.cols_total = 0;
  long long so = 0;
  P.n_groups = 0;
  for (int mtx = 0; mtx < P.n_mat; ++mtx) {
    const int pl = mtx >> 1;
    P.len[mtx] = (mtx & 1) ? L.cols : L.rows[pl];
    P.rec_off[mtx] = (mtx & 1) ? L.v_offset[pl] : L.u_offset[pl];
    P.ncols[mtx] = L.rank[pl];
    if (P.len[mtx] > d9::kMaxLen) return fail(LRFB_E_UNSUPPORTED, "column of %d bytes: the device deflate takes at most %d (use lrfb_qmf_pack_host)", P.len[mtx], d9::kMaxLen);
    if (P.ncols[mtx] > 64) return fail(LRFB_E_UNSUPPORTED, "rank above 64");
    P.col0[mtx] = P.cols_total, P.cols_total += P.ncols[mtx];
    P.slot[mtx] = d9::slot_bytes(P.len[mtx]);
    P.slot_off[mtx] = (int)so, so += (long long)P.ncols[mtx] * P.slot[mtx];
    bool seen = false;
    for (int g = 0; g < P.n_groups; ++g) seen |= P.group_len[g] == P.len[mtx];
    if (!seen) P.group_len[P.n_groups++] = P.len[mtx];
  }
  if (so > 0x7fffffffll) return fail(LRFB_E_UNSUPPORTED, "image record too large");
  std::sort(P.group_len, P.group_len + P.n_groups, [](int a, int b) { return a > b; });
  P.img_stride = so;
  long long scratch[6] = {0};
  for (int g = 0; g < P.n_groups; ++g) {
    long long streams = 0;
    for (int mtx = 0; mtx < P.n_mat; ++mtx)
      if (P.len[mtx] == P.group_len[g]) streams += (long long)batch * P.ncols[mtx];
    P.nw[g] = warps_per_column(P.group_len[g], streams);
    P.grid[g] = (int)std::min<long long>(streams, (long long)num_sms() * ctas_per_sm(P.group_len[g], P.nw[g]));
    scratch[g] = P.grid[g] * d9::scratch_per_cta(P.group_len[g]);
  }
  auto up = [](long long v) { return (v + 255) & ~255ll; };
  long long o = 0;
  P.off_cbuf = o, o = up(o + (long long)batch * P.img_stride);
  P.off_csize = o, o = up(o + 4ll * batch * P.cols_total);
  P.off_sizes = o, o = up(o + 8ll * batch);
  P.off_counter = o, o = up(o + 64);
  for (int g = 0; g < P.n_groups; ++g) P.off_scratch[g] = o, o = up(o + scratch[g]);
  P.total = o;
  return 0;
}
}  // namespace

LRFB_EXPORT int64_t lrfb_qmf_pack_device_workspace(const lrfb_qmf_config* cfg, int32_t batch) {
  PackPlan P;
  if (batch <= 0 || make_pack_plan(cfg, batch, P)) return -1;
  return P.total;
}

LRFB_EXPORT int32_t lrfb_qmf_pack_device(const lrfb_qmf_config* cfg, int32_t batch, const int8_t* d_records,
                                         const char* metadata_json, int64_t metadata_len, uint8_t* d_blob,
                                         int64_t blob_capacity, int64_t* d_offsets, void* d_workspace,
                                         int64_t workspace_bytes, void* stream) {
  if (!d_records || !metadata_json || !d_blob || !d_offsets || !d_workspace || batch <= 0 || metadata_len <= 0)
    return fail(LRFB_E_ARG, "bad arguments");
  if (metadata_len > 1024) return fail(LRFB_E_UNSUPPORTED, "metadata json above 1024 bytes");
  PackPlan P;
  int rc = make_pack_plan(cfg, batch, P);
  if (rc) return rc;
  if (workspace_bytes < P.total) return fail(LRFB_E_WORKSPACE, "workspace %lld < %lld", (long long)workspace_bytes, P.total);
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  if (cudaMemsetAsync(ws + P.off_counter, 0, 64, st) != cudaSuccess) return fail(LRFB_E_ARG, "memset failed");
  // The launch of the longest columns (luma) goes to the caller's stream, the others to a low-priority helper stream:
  // their CTAs move in as the long launch drains (its last wave is mostly empty), instead of waiting for its end.
#ifndef LRFB_SIM
  SideStream* side = P.n_groups > 1 ? side_stream(1) : nullptr;
  SideUnlock side_guard{side};
  if (side) {
    cudaEventRecord(side->fork, st);
    cudaStreamWaitEvent(side->stream, side->fork, 0);
  }
#endif
  for (int g = 0; g < P.n_groups; ++g) {
#ifndef LRFB_SIM
    cudaStream_t gs = (g > 0 && side) ? side->stream : st;
#else
    cudaStream_t gs = st;
#endif
    d9::Params K;
    memset(&K, 0, sizeof(K));
    K.rec = reinterpret_cast<const unsigned char*>(d_records), K.rec_stride = P.L.record_bytes;
    K.len = P.group_len[g], K.slot = d9::slot_bytes(K.len);
    for (int mtx = 0; mtx < P.n_mat; ++mtx)
      if (P.len[mtx] == K.len) {
        d9::ColSeg& sg = K.seg[K.n_seg++];
        sg.rec_off = (int)P.rec_off[mtx], sg.ncols = P.ncols[mtx], sg.col0 = P.col0[mtx], sg.out_off = P.slot_off[mtx];
        K.cols_per_image += P.ncols[mtx];
      }
    K.cols_total = P.cols_total, K.batch = batch;
    K.cbuf = ws + P.off_cbuf, K.img_stride = P.img_stride;
    K.csize = reinterpret_cast<unsigned*>(ws + P.off_csize);
    K.scratch = ws + P.off_scratch[g];
    K.counter = reinterpret_cast<int*>(ws + P.off_counter) + g;
    const int smem = d9::smem_bytes(K.len);
    const int nw = P.nw[g];
    const bool multi = K.len > d9::kOneBlock;
    typedef void (*kernel_t)(d9::Params);
    const kernel_t fn = nw == d9::kWarpsHuge ? (multi ? d9::deflate9_kernel<d9::kWarpsHuge, true> : d9::deflate9_kernel<d9::kWarpsHuge, false>)
                        : nw == d9::kWarps   ? (multi ? d9::deflate9_kernel<d9::kWarps, true> : d9::deflate9_kernel<d9::kWarps, false>)
                                             : (multi ? d9::deflate9_kernel<1, true> : d9::deflate9_kernel<1, false>);
#ifndef LRFB_SIM
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return fail((int)e, "deflate9 shared memory %d: %s", smem, cudaGetErrorString(e));
    }
    cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#endif
    LRFB_LAUNCH(fn, dim3(P.grid[g]), dim3(32 * nw), smem, gs, K);
    if ((rc = check_launch("deflate9_kernel"))) return rc;
  }
#ifndef LRFB_SIM
  if (side) {
    cudaEventRecord(side->join, side->stream);
    cudaStreamWaitEvent(st, side->join, 0);
  }
#endif
  d9::FrameParams F;
  memset(&F, 0, sizeof(F));
  F.n_mat = P.n_mat;
  for (int mtx = 0; mtx < P.n_mat; ++mtx) {
    F.ncols[mtx] = P.ncols[mtx], F.col0[mtx] = P.col0[mtx], F.slot_off[mtx] = P.slot_off[mtx], F.slot[mtx] = P.slot[mtx];
    F.hdr_len[mtx] = snprintf(F.hdr[mtx], sizeof(F.hdr[mtx]), "{\"num_fibers\": %d, \"mode\": \"col\", \"dtype\": \"int8\"}", P.ncols[mtx]);
  }
  F.meta_len = (int)metadata_len;
  memcpy(F.meta, metadata_json, (size_t)metadata_len);
  F.cols_total = P.cols_total, F.batch = batch;
  F.cbuf = ws + P.off_cbuf, F.img_stride = P.img_stride;
  F.csize = reinterpret_cast<const unsigned*>(ws + P.off_csize);
  F.sizes = reinterpret_cast<long long*>(ws + P.off_sizes);
  F.offsets = reinterpret_cast<long long*>(d_offsets);
  F.blob = d_blob, F.capacity = blob_capacity;
  LRFB_LAUNCH(d9::frame_sizes_kernel, dim3((batch + 127) / 128), dim3(128), 0, st, F);
  if ((rc = check_launch("frame_sizes_kernel"))) return rc;
  LRFB_LAUNCH(d9::frame_scan_kernel, dim3(1), dim3(1024), 0, st, (const long long*)F.sizes, F.offsets, (int)batch);
  if ((rc = check_launch("frame_scan_kernel"))) return rc;
  LRFB_LAUNCH(d9::frame_write_kernel, dim3(batch), dim3(d9::kFrameThreads), 0, st, F);
  return check_launch("frame_write_kernel");
}

// Host images -> finished byte streams in host memory, the whole of lrf.qmf_encode for a batch: the chunked pipeline of
// lrfb_qmf_encode_host with the lossless stage on the device behind every chunk (its own stream, so it overlaps the next
// chunk's encode kernels), and only the compressed streams on the way back.  The size of a chunk's blob is known when its
// offsets arrive, so the host trails the device by one chunk: A(i) = enqueue H2D, kernels, packer, offsets D2H;
// B(i) = wait for the offsets, enqueue the blob D2H.  Order A0 A1 B0 A2 B1 ...
LRFB_EXPORT int32_t lrfb_qmf_encode_bytes_host(lrfb_ctx* c, const lrfb_qmf_config* cfg, int32_t batch,
                                               const void* h_images, const char* metadata_json, int64_t metadata_len,
                                               uint8_t* h_blob, int64_t blob_capacity, int64_t* h_offsets) {
  if (!c || !h_images || !metadata_json || !h_blob || !h_offsets || batch <= 0 || metadata_len <= 0)
    return fail(LRFB_E_ARG, "bad arguments");
  lrfb_qmf_workspace_map m, mt;
  lrfb_qmf_layout L;
  int rc;
  if ((rc = lrfb_qmf_layout_query(cfg, &L))) return rc;
  const size_t img_bytes = (size_t)3 * cfg->height * cfg->width * (cfg->input_dtype == LRFB_U8 ? 1 : 4);
  // four times the chunk of the records pipeline: the deflate launches want a few thousand columns to fill the GPU
  // (measured on 4096 images of 768x512, host images -> streams: 16.1 Gpixel/s at 512 MiB, 15.9 at 1 GiB, 14.2 at 2 GiB)
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)batch, 4 * c->chunk_bytes / std::max<size_t>(img_bytes, 1)));
  if ((rc = lrfb_qmf_workspace_query(cfg, chunk, &m))) return rc;
  int64_t ws_need = m.total_bytes;
  if (batch % chunk) {
    if ((rc = lrfb_qmf_workspace_query(cfg, batch % chunk, &mt))) return rc;
    ws_need = std::max(ws_need, mt.total_bytes);
  }
  const int64_t pws_need = lrfb_qmf_pack_device_workspace(cfg, chunk);
  if (pws_need <= 0) return LRFB_E_UNSUPPORTED;  // message set by the plan
  const int64_t bound = lrfb_qmf_pack_bound(cfg, metadata_len);
  const size_t slot_blob = ((size_t)chunk * bound + 255) & ~(size_t)255, slot_offs = ((size_t)(chunk + 1) * 8 + 255) & ~(size_t)255;
#ifndef LRFB_SIM
  cudaSetDevice(c->device);
#endif
  if ((rc = grow(&c->d_in, &c->in_cap, (size_t)kHostSlots * chunk * img_bytes))) return rc;
  if ((rc = grow(&c->d_out, &c->out_cap, (size_t)2 * chunk * L.record_bytes))) return rc;
  if ((rc = grow(&c->d_ws, &c->ws_cap, (size_t)ws_need))) return rc;
  if ((rc = grow(&c->d_blob, &c->blob_cap, 2 * slot_blob))) return rc;
  if ((rc = grow(&c->d_offs, &c->offs_cap, 2 * slot_offs))) return rc;
  const size_t slot_pws = ((size_t)pws_need + 255) & ~(size_t)255;
  if ((rc = grow(&c->d_pws, &c->pws_cap, 2 * slot_pws))) return rc;
  if (c->hoffs_cap < 2 * slot_offs) {
#ifndef LRFB_SIM
    if (c->h_offs) cudaFreeHost(c->h_offs);
    c->h_offs = nullptr, c->hoffs_cap = 0;
    cudaError_t e = cudaHostAlloc((void**)&c->h_offs, 2 * slot_offs, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail((int)e, "cudaHostAlloc: %s", cudaGetErrorString(e));
#else
    free(c->h_offs);
    c->h_offs = (long long*)malloc(2 * slot_offs);
#endif
    c->hoffs_cap = 2 * slot_offs;
  }
  const unsigned char* src = reinterpret_cast<const unsigned char*>(h_images);
  const int n_chunks = (batch + chunk - 1) / chunk;
  int64_t written = 0;
  h_offsets[0] = 0;
  auto stage_b = [&](int idx) -> int {  // sizes of chunk idx are on the host: bring its blob back
    const int i0 = idx * chunk, n = std::min(chunk, batch - i0), oslot = idx & 1;
    const long long* ho = c->h_offs + (size_t)oslot * slot_offs / 8;
#ifndef LRFB_SIM
    cudaError_t e = cudaEventSynchronize(c->sized[oslot]);
    if (e != cudaSuccess) return fail((int)e, "pipeline: %s", cudaGetErrorString(e));
#endif
    const int64_t total = ho[n];
    if ((int64_t)slot_blob < total || written + total > blob_capacity)
      return fail(LRFB_E_WORKSPACE, "blob_capacity %lld too small (chunk %d needs %lld more)", (long long)blob_capacity, idx, (long long)total);
    for (int i = 0; i < n; ++i) h_offsets[i0 + i + 1] = written + ho[i + 1];
    const unsigned char* d_blob = reinterpret_cast<const unsigned char*>(c->d_blob) + (size_t)oslot * slot_blob;
#ifndef LRFB_SIM
    int r = d2h(h_blob + written, d_blob, (size_t)total, c->back);
    cudaEventRecord(c->drained[oslot], c->back);
#else
    int r = d2h(h_blob + written, d_blob, (size_t)total, c->stream);
#endif
    written += total;
    return r;
  };
  for (int idx = 0; idx < n_chunks; ++idx) {
    const int i0 = idx * chunk, n = std::min(chunk, batch - i0);
    const int slot = idx % kHostSlots, oslot = idx & 1;
    unsigned char* d_in = reinterpret_cast<unsigned char*>(c->d_in) + (size_t)slot * chunk * img_bytes;
    int8_t* d_out = reinterpret_cast<int8_t*>(c->d_out) + (size_t)oslot * chunk * L.record_bytes;
    uint8_t* d_blob = reinterpret_cast<uint8_t*>(c->d_blob) + (size_t)oslot * slot_blob;
    int64_t* d_offs = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(c->d_offs) + (size_t)oslot * slot_offs);
    long long* ho = c->h_offs + (size_t)oslot * slot_offs / 8;
#ifndef LRFB_SIM
    if (idx >= kHostSlots) cudaStreamWaitEvent(c->copy, c->consumed[slot], 0);
    if ((rc = h2d(d_in, src + (size_t)i0 * img_bytes, (size_t)n * img_bytes, c->copy))) return rc;
    cudaEventRecord(c->landed[slot], c->copy);
    cudaStreamWaitEvent(c->stream, c->landed[slot], 0);
    if (idx >= 2) cudaStreamWaitEvent(c->stream, c->packed[oslot], 0);  // the packer of chunk idx-2 is done with d_out[oslot]
    if ((rc = lrfb_qmf_encode(cfg, n, d_in, d_out, c->d_ws, (int64_t)c->ws_cap, nullptr, (void*)(uintptr_t)c->stream)))
      return rc;
    cudaEventRecord(c->consumed[slot], c->stream);
    cudaEventRecord(c->encoded[oslot], c->stream);
    cudaStream_t ps = c->pack[oslot];
    cudaStreamWaitEvent(ps, c->encoded[oslot], 0);
    if (idx >= 2) cudaStreamWaitEvent(ps, c->drained[oslot], 0);  // blob of chunk idx-2 has left d_blob[oslot] (B(idx-2) ran)
    if ((rc = lrfb_qmf_pack_device(cfg, n, d_out, metadata_json, metadata_len, d_blob, (int64_t)slot_blob, d_offs,
                                   reinterpret_cast<unsigned char*>(c->d_pws) + (size_t)oslot * slot_pws, (int64_t)slot_pws,
                                   (void*)(uintptr_t)ps)))
      return rc;
    cudaEventRecord(c->packed[oslot], ps);
    if ((rc = d2h(ho, d_offs, (size_t)(n + 1) * 8, ps))) return rc;
    cudaEventRecord(c->sized[oslot], ps);
#else
    if ((rc = h2d(d_in, src + (size_t)i0 * img_bytes, (size_t)n * img_bytes, c->stream))) return rc;
    if ((rc = lrfb_qmf_encode(cfg, n, d_in, d_out, c->d_ws, (int64_t)c->ws_cap, nullptr, nullptr))) return rc;
    if ((rc = lrfb_qmf_pack_device(cfg, n, d_out, metadata_json, metadata_len, d_blob, (int64_t)slot_blob, d_offs, c->d_pws,
                                   (int64_t)slot_pws, nullptr)))
      return rc;
    if ((rc = d2h(ho, d_offs, (size_t)(n + 1) * 8, c->stream))) return rc;
#endif
    if (idx >= 1 && (rc = stage_b(idx - 1))) return rc;
  }
  if ((rc = stage_b(n_chunks - 1))) return rc;
#ifndef LRFB_SIM
  if ((rc = sync_stream(c->back))) return rc;
  if ((rc = sync_stream(c->pack[0]))) return rc;
  if ((rc = sync_stream(c->pack[1]))) return rc;
#endif
  return sync_stream(c->stream);
}

// ---- the inverse on the device: un-frame + inflate a batch of encoded images into int8 records ---------------------------
LRFB_EXPORT int64_t lrfb_qmf_unpack_device_workspace(const lrfb_qmf_config* cfg, int32_t batch) {
  lrfb_qmf_layout L;
  if (batch <= 0 || lrfb_qmf_layout_query(cfg, &L)) return -1;
  int cols_total = 0;
  for (int pl = 0; pl < L.n_planes; ++pl) cols_total += 2 * L.rank[pl];
  return 256 + 8ll * batch * cols_total;
}

namespace {
int unpack_device_impl(const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* d_blob, const int64_t* d_offsets,
                       int8_t* d_records, void* d_workspace, int64_t workspace_bytes, void* stream, int image_base);
}
LRFB_EXPORT int32_t lrfb_qmf_unpack_device(const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* d_blob,
                                           const int64_t* d_offsets, int8_t* d_records, void* d_workspace,
                                           int64_t workspace_bytes, void* stream) {
  return unpack_device_impl(cfg, batch, d_blob, d_offsets, d_records, d_workspace, workspace_bytes, stream, 0);
}
namespace {
// `d_offsets` may point into a longer table (offsets are absolute positions in d_blob); `image_base` is the index of its
// first image in the caller's batch, for the error message
int unpack_device_impl(const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* d_blob, const int64_t* d_offsets,
                       int8_t* d_records, void* d_workspace, int64_t workspace_bytes, void* stream, int image_base) {
  if (!d_blob || !d_offsets || !d_records || !d_workspace || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  lrfb_qmf_layout L;
  int rc = lrfb_qmf_layout_query(cfg, &L);
  if (rc) return rc;
  d9i::Params P;
  memset(&P, 0, sizeof(P));
  P.blob = d_blob, P.offsets = reinterpret_cast<const long long*>(d_offsets), P.batch = batch;
  P.n_mat = 2 * L.n_planes;
  for (int mtx = 0; mtx < P.n_mat; ++mtx) {
    const int pl = mtx >> 1;
    P.ncols[mtx] = L.rank[pl], P.len[mtx] = (mtx & 1) ? L.cols : L.rows[pl];
    P.rec_off[mtx] = (mtx & 1) ? L.v_offset[pl] : L.u_offset[pl];
    P.col0[mtx] = P.cols_total, P.cols_total += P.ncols[mtx];
    P.max_len = std::max(P.max_len, P.len[mtx]);
  }
  const int smem = d9i::smem_bytes(P.max_len);
  if (smem > 227 * 1024) return fail(LRFB_E_UNSUPPORTED, "column of %d bytes does not fit the device inflate", P.max_len);
  if (workspace_bytes < 256 + 8ll * batch * P.cols_total) return fail(LRFB_E_WORKSPACE, "workspace too small");
  cudaStream_t st = (cudaStream_t)(uintptr_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
  P.error = reinterpret_cast<int*>(ws);
  P.col_pos = reinterpret_cast<unsigned*>(ws + 256);
  P.col_len = P.col_pos + (size_t)batch * P.cols_total;
  P.rec = d_records, P.rec_stride = L.record_bytes;
  if (cudaMemsetAsync(ws, 0, 256, st) != cudaSuccess) return fail(LRFB_E_ARG, "memset failed");
  LRFB_LAUNCH(d9i::unframe_kernel, dim3((batch + 127) / 128), dim3(128), 0, st, P);
  if ((rc = check_launch("unframe_kernel"))) return rc;
#ifndef LRFB_SIM
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(d9i::inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "inflate shared memory %d: %s", smem, cudaGetErrorString(e));
  }
#endif
  const long long streams = (long long)batch * P.cols_total;
  const int per_sm = std::max(1, std::min(32, (227 * 1024) / (smem + 1024)));
  const int grid = (int)std::min<long long>(streams, (long long)num_sms() * per_sm);
  LRFB_LAUNCH(d9i::inflate_kernel, dim3(grid), dim3(32), smem, st, P);
  if ((rc = check_launch("inflate_kernel"))) return rc;
  int err = 0;
#ifndef LRFB_SIM
  cudaError_t e = cudaMemcpyAsync(&err, P.error, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail((int)e, "lrfb_qmf_unpack_device: %s", cudaGetErrorString(e));
#else
  err = *P.error;
#endif
  if (err)
    return fail(LRFB_E_ARG, "encoded image %d is malformed (framing, deflate stream, length or adler32)", image_base + err - 1);
  return 0;
}
}  // namespace

// Encoded images in host memory -> uint8 images in host memory, the whole of lrf.qmf_decode for a batch: the streams go up
// once (a few KB per image), are un-framed and inflated on the device, and the decode kernel's output comes back chunk by
// chunk behind the next chunk's kernel.
LRFB_EXPORT int32_t lrfb_qmf_decode_bytes_host(lrfb_ctx* c, const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* h_blob,
                                               const int64_t* h_offsets, uint8_t* h_images) {
  if (!c || !h_blob || !h_offsets || !h_images || batch <= 0) return fail(LRFB_E_ARG, "bad arguments");
  lrfb_qmf_layout L;
  int rc;
  if ((rc = lrfb_qmf_layout_query(cfg, &L))) return rc;
  const int64_t total = h_offsets[batch] - h_offsets[0];
  if (h_offsets[0] != 0 || total <= 0) return fail(LRFB_E_ARG, "offsets must start at 0 and increase");
  const int64_t uws = lrfb_qmf_unpack_device_workspace(cfg, batch);
  if (uws <= 0) return fail(LRFB_E_UNSUPPORTED, "unsupported shape");
#ifndef LRFB_SIM
  cudaSetDevice(c->device);
#endif
  const size_t img_bytes = (size_t)3 * cfg->height * cfg->width;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)batch, c->chunk_bytes / std::max<size_t>(img_bytes, 1)));
  if ((rc = grow(&c->d_in, &c->in_cap, (size_t)batch * L.record_bytes))) return rc;
  if ((rc = grow(&c->d_out, &c->out_cap, (size_t)2 * chunk * img_bytes))) return rc;
  if ((rc = grow(&c->d_blob, &c->blob_cap, (size_t)total + 256))) return rc;
  if ((rc = grow(&c->d_offs, &c->offs_cap, (size_t)(batch + 1) * 8))) return rc;
  if ((rc = grow(&c->d_pws, &c->pws_cap, (size_t)uws))) return rc;
  if ((rc = h2d(c->d_blob, h_blob, (size_t)total, c->stream))) return rc;
  if ((rc = h2d(c->d_offs, h_offsets, (size_t)(batch + 1) * 8, c->stream))) return rc;
  int idx = 0;
  for (int i0 = 0; i0 < batch; i0 += chunk, ++idx) {
    const int n = std::min(chunk, batch - i0);
    const int oslot = idx & 1;
    uint8_t* d_img = reinterpret_cast<uint8_t*>(c->d_out) + (size_t)oslot * chunk * img_bytes;
    // un-framing + inflate chunk by chunk (the copy-back of the previous chunks runs meanwhile): done for the whole batch
    // up front it is 11 ms of a 4096-image call during which the copy engine, the bound of this call, has nothing to do
    if ((rc = unpack_device_impl(cfg, n, reinterpret_cast<const uint8_t*>(c->d_blob), reinterpret_cast<const int64_t*>(c->d_offs) + i0,
                                 reinterpret_cast<int8_t*>(c->d_in) + (size_t)i0 * L.record_bytes, c->d_pws, (int64_t)c->pws_cap,
                                 (void*)(uintptr_t)c->stream, i0))) {
#ifndef LRFB_SIM
      sync_stream(c->back);  // earlier chunks may still be on their way into h_images
#endif
      return rc;
    }
#ifndef LRFB_SIM
    if (idx >= 2) cudaStreamWaitEvent(c->stream, c->drained[oslot], 0);
#endif
    if ((rc = lrfb_qmf_decode(cfg, n, reinterpret_cast<const int8_t*>(c->d_in) + (size_t)i0 * L.record_bytes, d_img,
                              (void*)(uintptr_t)c->stream)))
      return rc;
#ifndef LRFB_SIM
    cudaEventRecord(c->encoded[oslot], c->stream);
    cudaStreamWaitEvent(c->back, c->encoded[oslot], 0);
    if ((rc = d2h(h_images + (size_t)i0 * img_bytes, d_img, (size_t)n * img_bytes, c->back))) return rc;
    cudaEventRecord(c->drained[oslot], c->back);
#else
    if ((rc = d2h(h_images + (size_t)i0 * img_bytes, d_img, (size_t)n * img_bytes, c->stream))) return rc;
#endif
  }
#ifndef LRFB_SIM
  if ((rc = sync_stream(c->back))) return rc;
#endif
  return sync_stream(c->stream);
}

#ifdef D9_PROF
LRFB_EXPORT int32_t lrfb_d9_prof(unsigned long long* out16, int32_t reset) {
  if (out16) cudaMemcpyFromSymbol(out16, d9::g_d9_prof, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(d9::g_d9_prof, z, sizeof(z));
  }
  return 0;
}
#endif

#ifdef LRFB_SIM
// test tooling (shim build only): the serial restatement, see deflate9.cuh
LRFB_EXPORT int64_t lrfb_sim_deflate9_serial(const uint8_t* in, int32_t n, uint8_t* out) {
  return d9::deflate9_serial(in, n, out, nullptr);
}
#endif

LRFB_EXPORT int32_t lrfb_debug_set(const char* knob, int32_t value) {
  if (!knob) return fail(LRFB_E_ARG, "knob is null");
  if (!strcmp(knob, "decode_v1")) {
    g_decode_v1.store(value);
    return 0;
  }
  return fail(LRFB_E_ARG, "unknown knob '%s'", knob);
}
