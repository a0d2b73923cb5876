// Device inverse of the lossless stage: un-frame a batch of encoded images and inflate every factor column back into the
// int8 record lrfb_qmf_decode reads — the head of lrf.qmf_decode (lrf/compression/qmf.py:313-327: separate_bytes,
// decode_tensor -> decode_matrix: zlib.decompress per column, lrf/compression/utils.py:393-426, :458-490).
//
// Inflate is RFC 1951 as zlib implements it (stored, fixed and dynamic blocks, any number of blocks per stream, the zlib
// wrapper with its adler32).  Any valid deflate stream of the right length decodes to the same bytes, so streams written
// by the reference (CPython's zlib), by lrfb_qmf_pack_host and by the device deflate are all accepted; the result is
// checked against zlib.decompress in tests/test_inflate.py.  Canonical-Huffman decoding follows the counting form
// (per-length code counts and a symbol table sorted by code): no large lookup tables, a few hundred bytes per stream.
//
// Mapping: one thread per image walks the framing (nested big-endian length prefixes) and notes where each column's
// stream starts; then one warp per column: lane 0 decodes the stream into a shared output buffer (match copies included),
// all lanes verify the adler32 and copy the column into the record.
#pragma once
#include "lrfb_common.cuh"

namespace lrfb {
namespace d9i {

struct Params {
  const unsigned char* blob;   // encoded images back to back
  const long long* offsets;    // [batch + 1]
  int batch;
  int n_mat;                   // 2 * planes: U_0, V_0, U_1, ...
  int ncols[6], len[6], col0[6];
  long long rec_off[6];
  int cols_total;
  signed char* rec;            // [batch][rec_stride]
  long long rec_stride;
  unsigned* col_pos;           // [batch][cols_total] offset of the column's zlib stream inside the image
  unsigned* col_len;           // [batch][cols_total]
  int* error;                  // 0, or 1 + index of the first image found malformed
  int max_len;                 // longest column (sizes the shared buffers)
};

__host__ __device__ inline int smem_bytes(int max_len) { return ((max_len + 64 + 15) & ~15) + 1024; }

__device__ inline unsigned be32(const unsigned char* p) {
  return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | (unsigned)p[3];
}
__device__ inline void flag_error(const Params& P, int img) { atomicCAS(P.error, 0, img + 1); }

// combine_bytes(parts) = k - 1 nested BE32 prefixes (outermost first: acc[k-2] ... acc[0], acc[j] = 4 + acc[j-1] + size_j,
// acc[0] = size_0), then the parts.  One thread per image.
__global__ void __launch_bounds__(128) unframe_kernel(Params P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.batch) return;
  const long long o0 = P.offsets[i], size = P.offsets[i + 1] - o0;
  const unsigned char* p = P.blob + o0;
  unsigned* cpos = P.col_pos + (long long)i * P.cols_total;
  unsigned* clen = P.col_len + (long long)i * P.cols_total;
  bool ok = size >= 8;
  long long pos = 0;
  if (ok) {
    const long long meta_len = be32(p);
    pos = 4 + meta_len;  // body starts here
    ok = pos + 4ll * (P.n_mat - 1) <= size;
  }
  long long msize[6] = {0, 0, 0, 0, 0, 0};
  if (ok) {
    // prefixes acc[k-2] ... acc[0]
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    for (int j = P.n_mat - 2; j >= 0; --j) acc[j] = be32(p + pos), pos += 4;
    long long used = 0;
    for (int j = 0; j < P.n_mat - 1; ++j) {
      msize[j] = j == 0 ? acc[0] : acc[j] - acc[j - 1] - 4;
      ok = ok && msize[j] >= 8;
      used += msize[j];
    }
    msize[P.n_mat - 1] = size - pos - used;
    ok = ok && msize[P.n_mat - 1] >= 8;
  }
  for (int mtx = 0; mtx < P.n_mat && ok; ++mtx) {
    const long long mend = pos + msize[mtx];
    ok = mend <= size;
    if (!ok) break;
    const long long hdr_len = be32(p + pos);
    long long q = pos + 4 + hdr_len;  // columns body
    const int R = P.ncols[mtx];
    ok = q + 4ll * (R - 1) <= mend;
    if (!ok) break;
    long long acc_prev = 0, used = 0;
    // prefixes come outermost first: read them into the sizes from the innermost end
    for (int r = 0; r < R - 1; ++r) {
      const long long a = be32(p + q + 4ll * (R - 2 - r));  // acc[r]
      const long long sz = r == 0 ? a : a - acc_prev - 4;
      ok = ok && sz >= 8 && sz <= 0x7fffffffll;
      clen[P.col0[mtx] + r] = (unsigned)sz;
      acc_prev = a, used += sz;
    }
    q += 4ll * (R - 1);
    const long long last = mend - q - used;
    ok = ok && last >= 8 && last <= 0x7fffffffll;
    clen[P.col0[mtx] + R - 1] = (unsigned)last;
    for (int r = 0; r < R && ok; ++r) {
      cpos[P.col0[mtx] + r] = (unsigned)q;
      q += clen[P.col0[mtx] + r];
    }
    ok = ok && q == mend;
    pos = mend;
  }
  if (!ok) flag_error(P, i);
}

// ---- inflate, serial on lane 0 --------------------------------------------------------------------------------------------
struct Huff {
  unsigned short* count;   // [16] codes per length
  unsigned short* symbol;  // symbols in canonical order
};
struct Bits {
  const unsigned char* in;
  int n, pos;              // bytes available / consumed
  unsigned long long buf;
  int cnt;
  bool bad;
  __device__ inline void fill() {
    while (cnt <= 56) {
      const unsigned long long b = pos < n ? in[pos] : 0;  // reading past the end is caught by `bad` below
      if (pos >= n + 8) bad = true;
      ++pos;
      buf |= b << cnt, cnt += 8;
    }
  }
  __device__ inline unsigned get(int k) {  // k <= 16
    if (cnt < k) fill();
    const unsigned v = (unsigned)(buf & ((1ull << k) - 1ull));
    buf >>= k, cnt -= k;
    return v;
  }
};
__device__ inline int build(Huff& h, const unsigned char* length, int n) {  // returns < 0 for an over-subscribed set
  for (int l = 0; l < 16; ++l) h.count[l] = 0;
  for (int s = 0; s < n; ++s) h.count[length[s]]++;
  if (h.count[0] == n) return 0;  // no codes: legal for the distance code of a block without matches
  int left = 1;
  for (int l = 1; l < 16; ++l) {
    left <<= 1;
    left -= h.count[l];
    if (left < 0) return left;
  }
  unsigned short offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (unsigned short)(offs[l] + h.count[l]);
  for (int s = 0; s < n; ++s)
    if (length[s] != 0) h.symbol[offs[length[s]]++] = (unsigned short)s;
  return left;
}
__device__ inline int decode_sym(Bits& b, const Huff& h) {
  if (b.cnt < 15) b.fill();  // a code is at most 15 bits: walk them in the buffer, consume what was used
  unsigned long long bits = b.buf;
  int code = 0, first = 0, index = 0;
  for (int l = 1; l < 16; ++l) {
    code |= (int)(bits & 1ull);
    bits >>= 1;
    const int count = h.count[l];
    if (code - count < first) {
      b.buf = bits, b.cnt -= l;
      return h.symbol[index + (code - first)];
    }
    index += count, first += count;
    first <<= 1, code <<= 1;
  }
  return -1;
}
__device__ inline int length_base(int idx) { return idx < 8 ? idx + 3 : idx == 28 ? 258 : ((4 + (idx & 3)) << ((idx >> 2) - 1)) + 3; }
__device__ inline int length_extra(int idx) { return (idx < 8 || idx == 28) ? 0 : (idx >> 2) - 1; }
__device__ inline int dist_base(int c) { return c < 4 ? c + 1 : ((2 + (c & 1)) << ((c >> 1) - 1)) + 1; }
__device__ inline int dist_extra(int c) { return c < 4 ? 0 : (c >> 1) - 1; }

// Decodes one zlib stream (in[0..n)) into out[0..want); returns true when the stream is well formed and exactly `want`
// bytes long.  The adler32 trailer is checked by the caller (all lanes).  trailer_pos: where the 4 checksum bytes start.
__device__ inline bool inflate_stream(const unsigned char* in, int n, unsigned char* out, int want, unsigned char* scratch,
                                      int& trailer_pos) {
  if (n < 6 || (in[0] & 0x0f) != 8 || (((unsigned)in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return false;
  Bits b{in, n, 2, 0ull, 0, false};
  Huff lc{reinterpret_cast<unsigned short*>(scratch), reinterpret_cast<unsigned short*>(scratch) + 16};
  Huff dc{reinterpret_cast<unsigned short*>(scratch) + 16 + 288, reinterpret_cast<unsigned short*>(scratch) + 32 + 288};
  unsigned char* lengths = scratch + 2 * (32 + 288 + 32);  // [320]
  int op = 0, last;
  do {
    last = (int)b.get(1);
    const int type = (int)b.get(2);
    if (type == 0) {  // stored: skip to the byte boundary (the reader buffers whole bytes), LEN, NLEN, the bytes
      b.get(b.cnt & 7);
      const unsigned len = b.get(16), nlen = b.get(16);
      if ((len ^ 0xffffu) != nlen || op + (int)len > want) return false;
      for (unsigned k = 0; k < len; ++k) out[op++] = (unsigned char)b.get(8);
      if (b.bad) return false;
      continue;
    }
    if (type == 3) return false;
    if (type == 1) {
      for (int s = 0; s < 144; ++s) lengths[s] = 8;
      for (int s = 144; s < 256; ++s) lengths[s] = 9;
      for (int s = 256; s < 280; ++s) lengths[s] = 7;
      for (int s = 280; s < 288; ++s) lengths[s] = 8;
      build(lc, lengths, 288);
      for (int s = 0; s < 30; ++s) lengths[s] = 5;
      build(dc, lengths, 30);
    } else {
      const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
      if (nlen > 286 || ndist > 30) return false;
      const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      for (int i = 0; i < 19; ++i) lengths[i] = 0;
      for (int i = 0; i < ncode; ++i) lengths[order[i]] = (unsigned char)b.get(3);
      if (build(lc, lengths, 19) != 0) return false;  // the code-length code must be complete
      int idx = 0;
      while (idx < nlen + ndist) {
        int sym = decode_sym(b, lc);
        if (sym < 0) return false;
        if (sym < 16) {
          lengths[idx++] = (unsigned char)sym;
        } else {
          int len = 0, rep;
          if (sym == 16) {
            if (idx == 0) return false;
            len = lengths[idx - 1], rep = 3 + (int)b.get(2);
          } else if (sym == 17) {
            rep = 3 + (int)b.get(3);
          } else {
            rep = 11 + (int)b.get(7);
          }
          if (idx + rep > nlen + ndist) return false;
          while (rep--) lengths[idx++] = (unsigned char)len;
        }
      }
      if (lengths[256] == 0) return false;
      // the two tables are built from one array: move the distance lengths aside first
      unsigned char* dl = lengths + 288;
      for (int i = ndist - 1; i >= 0; --i) dl[i] = lengths[nlen + i];
      int e = build(lc, lengths, nlen);
      if (e < 0 || (e > 0 && nlen - lc.count[0] != 1)) return false;
      e = build(dc, dl, ndist);
      if (e < 0 || (e > 0 && ndist - dc.count[0] != 1)) return false;
    }
    for (;;) {
      const int sym = decode_sym(b, lc);
      if (sym < 0 || b.bad) return false;
      if (sym < 256) {
        if (op >= want) return false;
        out[op++] = (unsigned char)sym;
      } else if (sym == 256) {
        break;
      } else {
        const int li = sym - 257;
        if (li >= 29) return false;
        const int len = length_base(li) + (int)b.get(length_extra(li));
        const int ds = decode_sym(b, dc);
        if (ds < 0 || ds >= 30) return false;
        const int dist = dist_base(ds) + (int)b.get(dist_extra(ds));
        if (dist > op || op + len > want) return false;
        for (int k = 0; k < len; ++k, ++op) out[op] = out[op - dist];
      }
    }
  } while (!last);
  if (op != want || b.bad) return false;
  // the checksum starts at the next byte boundary: bytes fetched minus whole bytes still buffered
  trailer_pos = b.pos - (b.cnt >> 3);
  return trailer_pos + 4 <= n;
}


// One warp per column stream.
__global__ void __launch_bounds__(32) inflate_kernel(Params P) {
  LRFB_DYN_SMEM(smem);
  const int lane = threadIdx.x;
  const int pad = (P.max_len + 64 + 15) & ~15;
  unsigned char* out = smem;
  unsigned char* scratch = smem + pad;
  const long long total = (long long)P.batch * P.cols_total;
  for (long long s = blockIdx.x; s < total; s += gridDim.x) {
    if (*reinterpret_cast<volatile int*>(P.error)) break;
    const int img = (int)(s / P.cols_total), c = (int)(s - (long long)img * P.cols_total);
    int mtx = 0;
    while (mtx + 1 < P.n_mat && c >= P.col0[mtx] + P.ncols[mtx]) ++mtx;
    const int want = P.len[mtx], r = c - P.col0[mtx];
    const int n = (int)P.col_len[(long long)img * P.cols_total + c];
    // the stream (about a sixth of the column) is read where it lies: lane 0's byte loads hit L1 after the first touch of a line
    const unsigned char* in = P.blob + P.offsets[img] + P.col_pos[(long long)img * P.cols_total + c];
    int ok = 0, tpos = 0;
    if (lane == 0) ok = inflate_stream(in, n, out, want, scratch, tpos) ? 1 : 0;
    ok = __shfl_sync(0xffffffffu, ok, 0);
    tpos = __shfl_sync(0xffffffffu, tpos, 0);
    __syncwarp();
    if (ok) {
      unsigned long long sa = 0, sb = 0;  // adler32: a = 1 + sum d_i, b = n + sum (n - i) d_i
      for (int i = lane; i < want; i += 32) sa += out[i], sb += (unsigned long long)(want - i) * out[i];
#pragma unroll
      for (int o = 16; o; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o), sb += __shfl_xor_sync(0xffffffffu, sb, o);
      const unsigned a = (unsigned)((1 + sa) % 65521ull), b2 = (unsigned)((want + sb) % 65521ull);
      ok = be32(in + tpos) == ((b2 << 16) | a);
    }
    if (!ok) {
      if (lane == 0) flag_error(P, img);
      continue;
    }
    signed char* dst = P.rec + (long long)img * P.rec_stride + P.rec_off[mtx] + (long long)r * want;
    for (int i = lane; i < want; i += 32) dst[i] = (signed char)out[i];
    __syncwarp();
  }
}

}  // namespace d9i
}  // namespace lrfb
