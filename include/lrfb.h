/*
 * lrfb.h — C ABI of liblrfb.so, the B200 (sm_100a) implementation of the lrf QMF codec hot path.
 *
 * The reference (pashtari/lrf) is pure Python: it has no FFI layer for this path, so the drop-in
 * boundary is the Python function API plus the encoded-bytes layout.  These entry
 * points are what a binding for that path needs; each one names the reference code it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions: plain pointers and sizes only; `d_` pointers are CUDA device pointers, `h_` pointers
 * are host pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream); no hidden
 * allocation on the device-pointer entry points (the caller provides the workspace); every call
 * returns 0 on success, a negative LRFB_E* code for a bad argument / unsupported shape, or a positive
 * cudaError_t.  lrfb_last_error() returns a thread-local message for the last failure.
 * Nothing here has a CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef LRFB_H_
#define LRFB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRFB_ABI_VERSION 1

#define LRFB_E_ARG -1         /* null pointer, non-positive size, bad enum */
#define LRFB_E_UNSUPPORTED -2 /* shape / rank / bounds outside what the kernels implement */
#define LRFB_E_WORKSPACE -3   /* workspace too small */

#define LRFB_RGB 0
#define LRFB_YCBCR 1
#define LRFB_U8 0
#define LRFB_F32 1

/* Parameters of lrf.qmf_encode (lrf/compression/qmf.py:116-127) after the host resolved the
 * rank rule (:215-225, :244-250).  patch=True branches (patch=False: lrfb_qmf_frontend with 1x1 patches + lrfb_factorize). */
typedef struct lrfb_qmf_config {
  int32_t height, width;    /* image (3, H, W) */
  int32_t patch_h, patch_w; /* patch_size */
  int32_t color_space;      /* LRFB_RGB or LRFB_YCBCR */
  int32_t input_dtype;      /* LRFB_U8 or LRFB_F32 (image.float()) */
  double scale_h, scale_w;  /* scale_factor of the chroma down-sampling (YCbCr only) */
  int32_t rank[3];          /* per plane (Y, Cb, Cr) or rank[0] for RGB */
  float bound_lo, bound_hi; /* bounds */
  int32_t num_iters;        /* QMF(num_iters=...) */
} lrfb_qmf_config;

/* Geometry derived from a config: the metadata qmf_encode stores ("original size", "padded size",
 * "rank", lrf/compression/qmf.py:252-254) and the layout of one image's int8 factor record:
 *   [U_0 | V_0 | U_1 | V_1 | U_2 | V_2], each factor fiber-major (column r of the (rows x R)
 *   matrix is contiguous: exactly the bytes encode_matrix zlib-compresses per column,
 *   lrf/compression/utils.py:368-378). */
typedef struct lrfb_qmf_layout {
  int32_t n_planes;
  int32_t cols; /* N = patch_h*patch_w (YCbCr) or 3*patch_h*patch_w (RGB) */
  int32_t orig_h[3], orig_w[3], pad_h[3], pad_w[3];
  int32_t rows[3]; /* M per plane */
  int32_t rank[3];
  int64_t u_offset[3], v_offset[3]; /* bytes from the start of the record */
  int64_t record_bytes;
  int64_t x_floats; /* f32 elements of all patch matrices of one image */
} lrfb_qmf_layout;

/* Byte offsets inside the encode workspace (for tests and stage-level callers); plane-major. */
typedef struct lrfb_qmf_workspace_map {
  int64_t x[3];     /* f32 [batch][rows][cols] */
  int64_t u[3];     /* f32 [batch][rows][rank]   init, then final factors */
  int64_t v[3];     /* f32 [batch][cols][rank] */
  int64_t gram[3];  /* f64 [batch][cols][cols] */
  int64_t evec[3];  /* f64 [batch][cols][rank] */
  int64_t sigma[3]; /* f64 [batch][rank] */
  int64_t total_bytes;
} lrfb_qmf_workspace_map;

/* Test hooks for lrfb_qmf_encode (all optional). */
typedef struct lrfb_qmf_debug {
  const float* d_init_u[3];     /* inject the SVD init (teacher forcing): [batch][rows][rank] */
  const float* d_init_v[3];     /* [batch][cols][rank] */
  const int32_t* d_sign_flip[3]; /* [batch][rank] of +1/-1 applied on top of the sign convention */
  int32_t stop_after;            /* 0: full encode, 1: after the front end, 2: after the SVD init */
} lrfb_qmf_debug;

int32_t lrfb_abi_version(void);
const char* lrfb_last_error(void);
/* number of CUDA devices visible, or a negative/positive error */
int32_t lrfb_device_count(void);

int32_t lrfb_qmf_layout_query(const lrfb_qmf_config* cfg, lrfb_qmf_layout* out);
int32_t lrfb_qmf_workspace_query(const lrfb_qmf_config* cfg, int32_t batch, lrfb_qmf_workspace_map* out);

/* Replaces the body of lrf.qmf_encode between image.float() and the int8 cast
 * (lrf/compression/qmf.py:227-262): rgb_to_ycbcr, chroma_downsampling(area), pad_image(reflect),
 * patchify, QMF.decompose (SVDInit + num_iters CoordinateDescent sweeps), .to(int8).
 * d_images: [batch][3][H][W] of input_dtype; d_factors: [batch][record_bytes] int8. */
int32_t lrfb_qmf_encode(const lrfb_qmf_config* cfg, int32_t batch, const void* d_images, int8_t* d_factors,
                        void* d_workspace, int64_t workspace_bytes, const lrfb_qmf_debug* dbg, void* stream);

/* Replaces lrf.qmf_decode after the byte un-packing (lrf/compression/qmf.py:313-351):
 * QMF.reconstruct, depatchify, unpad_image, chroma_upsampling(nearest), ycbcr_to_rgb, to_dtype(uint8).
 * d_images: [batch][3][H][W] uint8. */
int32_t lrfb_qmf_decode(const lrfb_qmf_config* cfg, int32_t batch, const int8_t* d_factors, uint8_t* d_images,
                        void* stream);

/* Replaces lrf.qmf_decode for patch=False streams after the byte un-packing (lrf/compression/qmf.py:303-311 RGB,
 * :339-351 YCbCr): QMF.reconstruct of whole-channel factors, chroma_upsampling(nearest), ycbcr_to_rgb, to_dtype(uint8).
 * Factors are row-major with their batch dimension, as the reference stores them: YCbCr: d_u[pl] [batch][h_pl][R_pl],
 * d_v[pl] [batch][w_pl][R_pl] for pl = Y, Cb, Cr (chroma planes chroma_h x chroma_w); RGB: d_u[0] [batch][3][H][R],
 * d_v[0] [batch][3][W][R].  The encode side of these branches is lrfb_qmf_frontend (1x1 "patches" give the planes)
 * followed by lrfb_factorize per plane. */
int32_t lrfb_qmf_decode_planes(int32_t color_space, int32_t height, int32_t width, int32_t chroma_h, int32_t chroma_w,
                               const int32_t* rank, int32_t batch, const int8_t* const* d_u, const int8_t* const* d_v,
                               uint8_t* d_images, void* stream);

/* Stage-level: only the front end (compression/utils.py:24-47, :76-95, :108-132, compression/qmf.py:43-56).
 * d_x: f32, plane-major [plane][batch][rows][cols] packed (offsets = workspace map x[] minus x[0]). */
int32_t lrfb_qmf_frontend(const lrfb_qmf_config* cfg, int32_t batch, const void* d_images, float* d_x,
                          void* stream);

/* Replaces lrf.QMF(rank, num_iters, bounds).decompose(x) (lrf/factorization/qmf.py:197-214) for a batch of
 * equally shaped matrices with factor=(0,1): d_x [n_mat][M][N] f32 → d_u [n_mat][M][R], d_v [n_mat][N][R]
 * (integer-valued f32).  d_init_u/d_init_v optional (skip the SVD init).  Workspace from
 * lrfb_factorize_workspace_bytes. */
int64_t lrfb_factorize_workspace_bytes(int32_t n_mat, int32_t M, int32_t N, int32_t R);
int32_t lrfb_factorize(const float* d_x, int32_t n_mat, int32_t M, int32_t N, int32_t R, float bound_lo,
                       float bound_hi, int32_t num_iters, float* d_u, float* d_v, const float* d_init_u,
                       const float* d_init_v, const int32_t* d_sign_flip, void* d_workspace,
                       int64_t workspace_bytes, void* stream);

/* Stage-level: only the block-coordinate-descent sweeps (lrf/factorization/qmf.py:207-212 with
 * CoordinateDescent.forward :149-164) in place on d_u [n_mat][M][R] / d_v [n_mat][N][R].  On entry d_v holds
 * the initialisation v0; d_u holds u0 when d_s0 is NULL, otherwise d_u is output only and the first half-sweep
 * derives u0 = (X v0) / s from d_s0 [n_mat][R] (f32 singular values) exactly as lrfb_qmf_encode does.
 * flags bit 0 (LRFB_X_IN_U8_RANGE): every entry of d_x lies in [0, 256) — true for the planes of the uint8 front
 * end; allows the exact fixed-point tensor-core path for the X^T U half-sweep.  Workspace: at least 592*(2*R*R + N*R)*4 bytes. */
#define LRFB_X_IN_U8_RANGE 1u
int32_t lrfb_bcd(const float* d_x, int32_t n_mat, int32_t M, int32_t N, int32_t R, float bound_lo,
                 float bound_hi, int32_t num_iters, float* d_u, float* d_v, const float* d_s0, uint32_t flags,
                 void* d_workspace, int64_t workspace_bytes, void* stream);

/* Measurement helpers: number of kernels this library has launched so far (process-wide), and an FP32
 * FFMA throughput probe (launches num_SMs*8 blocks of 256 threads doing iters*32 dependent-chain FMAs on
 * 16 chains: flops = num_SMs*8*256*iters*64) used as the FP32 roofline denominator. */
int64_t lrfb_launch_count(void);
int32_t lrfb_ffma_probe(float* d_out, int32_t iters, void* stream);

/* SVD baseline codec, color_space RGB + patch (replaces lrf/compression/svd.py:156-187: pad_image, patchify,
 * torch.linalg.svd, top-R, *sqrt(s), quantize(u), quantize(v)).  cfg->color_space = LRFB_RGB, cfg->rank[0] = R
 * (bounds / num_iters unused).  d_codes: [batch][record_bytes] uint8 in the lrfb_qmf_layout record layout
 * (U then V, fiber-major); d_qparams: [batch][4] f32 = {scale_u, min_u, scale_v, min_v} — the values the
 * reference stores under metadata["quantization"].  Workspace: lrfb_qmf_workspace_query. */
int32_t lrfb_svd_encode(const lrfb_qmf_config* cfg, int32_t batch, const void* d_images, uint8_t* d_codes,
                        float* d_qparams, void* d_workspace, int64_t workspace_bytes, const lrfb_qmf_debug* dbg,
                        void* stream);
/* Replaces lrf.svd_decode after un-packing (svd.py:316-359): dequantize, u @ v.T, depatchify, unpad, to uint8.
 * d_qparams6: [batch][6] f32 = {scale_u, min_u, scale_v, min_v, min code of u, min code of v}. */
int32_t lrfb_svd_decode(const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* d_codes, const float* d_qparams6,
                        uint8_t* d_images, void* stream);

/* Exact per-image sum of squared differences of two uint8 batches (lrf/utils/metrics.py:24-35 before
 * the mean); d_sse [batch] must be zeroed by the caller. psnr = 20*log10(255/sqrt(sse/n)). */
int32_t lrfb_sse_u8(const uint8_t* d_a, const uint8_t* d_b, int64_t elems_per_image, int32_t batch,
                    uint64_t* d_sse, void* stream);

/* Host-buffer path (what a CPU caller of lrf.qmf_encode / qmf_decode binds): a context owns a stream
 * and grow-only device buffers; h_ pointers should be pinned for full copy bandwidth. */
typedef struct lrfb_ctx lrfb_ctx;
int32_t lrfb_ctx_create(int32_t device, lrfb_ctx** out);
void lrfb_ctx_destroy(lrfb_ctx* ctx);
int32_t lrfb_qmf_encode_host(lrfb_ctx* ctx, const lrfb_qmf_config* cfg, int32_t batch, const void* h_images,
                             int8_t* h_factors);
int32_t lrfb_qmf_decode_host(lrfb_ctx* ctx, const lrfb_qmf_config* cfg, int32_t batch, const int8_t* h_factors,
                             uint8_t* h_images);
/* Tuning: input bytes per pipeline chunk of the two host-buffer calls (default 256 MiB; the H2D copies of two chunks
 * are in flight while the kernels consume a third). */
int32_t lrfb_ctx_set_chunk_bytes(lrfb_ctx* ctx, int64_t bytes);

/* Lossless packing on host threads — replaces the tail of lrf.qmf_encode (lrf/compression/qmf.py:288-292):
 * encode_tensor / encode_matrix (zlib level 9 per factor column, lrf/compression/utils.py:354-390, :429-455),
 * combine_bytes (:246-300) and the metadata header, byte for byte.  h_records: [batch][record_bytes] int8 records as
 * lrfb_qmf_encode(_host) writes them; metadata_json: the utf-8 JSON header (identical for every image of a batch of
 * one shape); image i's stream is written to h_out + i*out_stride (out_stride >= lrfb_qmf_pack_bound) and its length
 * to out_sizes[i].  threads <= 0: one worker per hardware thread.  Pure host code (zlib), no device work. */
int64_t lrfb_qmf_pack_bound(const lrfb_qmf_config* cfg, int64_t metadata_len);
int32_t lrfb_qmf_pack_host(const lrfb_qmf_config* cfg, int32_t batch, const int8_t* h_records,
                           const char* metadata_json, int64_t metadata_len, uint8_t* h_out, int64_t out_stride,
                           int64_t* out_sizes, int32_t threads);

/* Lossless stage on the device — the same tail of lrf.qmf_encode (lrf/compression/qmf.py:288-292; encode_matrix:
 * zlib.compress(column, level=9), lrf/compression/utils.py:354-390; combine_bytes :246-300), produced on the GPU byte
 * for byte: one warp deflates one factor column with zlib's level-9 algorithm restated (lrf_b200/csrc/deflate9.cuh),
 * then every image is framed.  d_records: [batch][record_bytes] as lrfb_qmf_encode writes them (device memory);
 * image i's stream is d_blob[d_offsets[i] .. d_offsets[i+1]) (d_offsets: batch + 1 int64 in device memory, streams packed
 * back to back).  blob_capacity >= batch * lrfb_qmf_pack_bound(cfg, metadata_len) always suffices; when
 * d_offsets[batch] > blob_capacity the images that did not fit are not written (check after the copy back).
 * Columns longer than 65 024 bytes (zlib's 64 KB window would slide) return LRFB_E_UNSUPPORTED and
 * lrfb_qmf_pack_device_workspace returns -1: use lrfb_qmf_pack_host for those shapes. */
int64_t lrfb_qmf_pack_device_workspace(const lrfb_qmf_config* cfg, int32_t batch);
int32_t lrfb_qmf_pack_device(const lrfb_qmf_config* cfg, int32_t batch, const int8_t* d_records,
                             const char* metadata_json, int64_t metadata_len, uint8_t* d_blob, int64_t blob_capacity,
                             int64_t* d_offsets, void* d_workspace, int64_t workspace_bytes, void* stream);

/* The inverse on the device — the head of lrf.qmf_decode (lrf/compression/qmf.py:313-327: separate_bytes, decode_tensor ->
 * zlib.decompress per column, lrf/compression/utils.py:393-426, :458-490): un-frames `batch` encoded images of one shape
 * (d_blob[d_offsets[i] .. d_offsets[i+1]), device memory) and inflates every factor column into the int8 records
 * lrfb_qmf_decode reads.  Accepts any valid zlib stream (reference, host packer, device packer).  Synchronises the
 * stream and returns LRFB_E_ARG naming the first malformed image (framing, deflate data, column length, adler32). */
int64_t lrfb_qmf_unpack_device_workspace(const lrfb_qmf_config* cfg, int32_t batch);
int32_t lrfb_qmf_unpack_device(const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* d_blob, const int64_t* d_offsets,
                               int8_t* d_records, void* d_workspace, int64_t workspace_bytes, void* stream);

/* The whole of lrf.qmf_encode for a batch with HOST buffers (lrf/compression/qmf.py:116-292): images in, finished byte
 * streams out.  The chunked pipeline of lrfb_qmf_encode_host with lrfb_qmf_pack_device behind every chunk; only the
 * compressed streams cross PCIe on the way back.  Image i's stream is h_blob[h_offsets[i] .. h_offsets[i+1]);
 * h_offsets has batch + 1 entries; blob_capacity >= batch * lrfb_qmf_pack_bound always suffices (LRFB_E_WORKSPACE
 * otherwise); h_images and h_blob should be pinned.  Same shape limits as lrfb_qmf_pack_device. */
int32_t lrfb_qmf_encode_bytes_host(lrfb_ctx* ctx, const lrfb_qmf_config* cfg, int32_t batch, const void* h_images,
                                   const char* metadata_json, int64_t metadata_len, uint8_t* h_blob,
                                   int64_t blob_capacity, int64_t* h_offsets);

/* The whole of lrf.qmf_decode for a batch with HOST buffers (lrf/compression/qmf.py:295-353): encoded images in
 * (h_blob[h_offsets[i] .. h_offsets[i+1]), h_offsets[0] = 0, one shape), uint8 images out ([batch][3][H][W]).
 * lrfb_qmf_unpack_device + lrfb_qmf_decode with the copy-back of the images chunked behind the kernels; h_images should be
 * pinned.  Returns LRFB_E_ARG naming the first malformed image. */
int32_t lrfb_qmf_decode_bytes_host(lrfb_ctx* ctx, const lrfb_qmf_config* cfg, int32_t batch, const uint8_t* h_blob,
                                   const int64_t* h_offsets, uint8_t* h_images);

/* Test hook: select a kernel variant process-wide.  Knobs: "decode_v1" (1 = per-row float decoder instead of the
 * int8 dot-product one).  The shipped library reads no environment variables. */
int32_t lrfb_debug_set(const char* knob, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* LRFB_H_ */
