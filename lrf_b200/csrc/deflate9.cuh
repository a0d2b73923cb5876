// Device DEFLATE that is byte-identical to zlib level 9 — the lossless stage of lrf.qmf_encode
// (lrf/compression/utils.py:354-390 encode_matrix: zlib.compress(column, level=9) for every factor column).
//
// zlib is a third-party dependency of the reference (CPython's zlib module, zlib 1.2.x / 1.3: the deflate
// algorithm at level 9 has been output-stable across those versions for inputs of one block); its published algorithm is
// restated here: deflate_slow (lazy matching, good_length 32, max_lazy 258, nice_length 258, max_chain 4096,
// 15-bit rolling hash of 3 bytes, TOO_FAR 4096, matches no farther than MAX_DIST = 32 506), a block flushed whenever the
// symbol buffer holds 16 383 entries (columns up to 16 382 bytes are one block), no window slide (inputs of at most kMaxLen
// = 65 024 bytes sit in zlib's 64 KB window at once), trees.c's heap-ordered Huffman construction with its
// depth tie-break and overflow repair, the run-length coded tree header, the stored / static / dynamic choice, and
// the zlib wrapper (78 DA ... adler32).  Parity anchor: the system zlib itself (tests/test_deflate9.py compares every
// stream byte for byte).
//
// Mapping: one warp per column (four warps per long column when the batch has fewer long columns than single-warp slots).
// The hash chains zlib walks serially (prev[] links, newest first) become a random-access structure: positions are
// radix-sorted by (hash, position), so the chain of position p is the run of entries just below p's rank, and the lanes
// test 64 chain candidates per step: a four-byte filter (the two bytes of zlib's own quick check around the current best
// length plus two offsets that follow where the last survivors parted from the string at p), then the survivors are compared
// in chain order by the whole warp, 128 bytes per round.  "The first candidate that reaches nice_length stops the walk,
// otherwise the closest of the longest" is exactly what the serial walk with its strict > update returns.  The
// chain-length budget (4096, a quarter of it once the previous match is >= 32 long) counts candidates the same way.  Lane 0
// builds the three Huffman trees (small: <= 286 leaves) and the tree header; the symbols are then coded by all lanes
// (prefix sum of code lengths, OR into the staged output).
#pragma once
#include "lrfb_common.cuh"

namespace lrfb {
namespace d9 {

constexpr int kMaxLen = 65024;       // longest column the device path takes: the whole input sits in zlib's 64 KB window, which never slides
constexpr int kOneBlock = 16382;     // columns up to here are one deflate block (fewer symbols than the 16 383-entry buffer)
constexpr int kBlockSymbols = 16383; // zlib flushes a block when its symbol buffer holds lit_bufsize - 1 entries
constexpr int kMaxDist = 32506;      // w_size - MIN_LOOKAHEAD: the farthest match zlib takes
constexpr int kHeap = 573;      // 2 * L_CODES + 1
constexpr int kLCodes = 286, kDCodes = 30, kBLCodes = 19;

#ifdef LRFB_SIM
inline int d9_clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
#else
__host__ __device__ inline int d9_clz(unsigned v) {
#ifdef __CUDA_ARCH__
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
#endif

// ---- the DEFLATE code tables, in closed form ----------------------------------------------------------------------
__host__ __device__ inline int length_code(int lc) {  // lc = match length - 3, 0..255 -> 0..28 (trees.c _length_code)
  if (lc < 8) return lc;
  if (lc == 255) return 28;
  const int e = 29 - d9_clz((unsigned)lc);  // extra bits = floor(log2 lc) - 2
  return 4 * e + (lc >> e);
}
__host__ __device__ inline int extra_lbits(int code) { return (code < 8 || code == 28) ? 0 : (code >> 2) - 1; }
__host__ __device__ inline int dist_code(int d0) {  // d0 = distance - 1, 0..32767 -> 0..29
  if (d0 < 4) return d0;
  const int e = 30 - d9_clz((unsigned)d0);  // floor(log2 d0) - 1
  return 2 * e + 2 + ((d0 >> e) & 1);
}
__host__ __device__ inline int extra_dbits(int code) { return code < 4 ? 0 : (code >> 1) - 1; }
__host__ __device__ inline int extra_blbits(int code) { return code == 16 ? 2 : code == 17 ? 3 : code == 18 ? 7 : 0; }
__host__ __device__ inline int static_llen(int n) { return n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8; }
__host__ __device__ inline unsigned static_lcode(int n) {  // before bit reversal
  return n < 144 ? 0x30u + n : n < 256 ? 0x190u + (n - 144) : n < 280 ? (unsigned)(n - 256) : 0xC0u + (n - 280);
}
__host__ __device__ inline unsigned bi_reverse(unsigned code, int len) {
  unsigned res = 0;
  do {
    res |= code & 1;
    code >>= 1, res <<= 1;
  } while (--len > 0);
  return res >> 1;
}
__host__ __device__ inline int bl_order(int i) {
  const unsigned char t[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  return t[i];
}
__host__ __device__ inline unsigned hash3(const unsigned char* w, int p) {  // zlib's ins_h after three UPDATE_HASH steps
  return (((unsigned)w[p] & 31u) << 10) ^ ((unsigned)w[p + 1] << 5) ^ (unsigned)w[p + 2];
}

// ---- serial part (lane 0 on the device): trees.c -------------------------------------------------------------------
struct Tree {
  unsigned short* freq;  // [nodes]
  unsigned short* dad;   // [nodes]
  unsigned char* len;    // [nodes + 1]
  unsigned short* code;  // [elems]
  int elems, kind, max_length, max_code;  // kind 0 literal/length, 1 distance, 2 bit-length
};
struct Work {
  Tree l, d, b;
  unsigned short* heap;  // [kHeap]
  unsigned char* depth;  // [kHeap]
  unsigned short bl_count[16];
  int heap_len, heap_max;
  unsigned opt_len, static_len;
};
// bytes of scratch behind a Work (everything but the leaf frequencies, which live through the parse)
constexpr int kFreqBytes = 2 * (kHeap + 61 + 39);                                       // l, d, b frequencies
constexpr int kTreeBytes = 2 * (kHeap + 61 + 39) + (kHeap + 1 + 62 + 40) + 2 * (kLCodes + kDCodes + kBLCodes) + 2 * kHeap + kHeap + 16;

__host__ __device__ inline void work_bind(Work& w, unsigned char* freq_mem, unsigned char* tree_mem) {
  unsigned short* f = reinterpret_cast<unsigned short*>(freq_mem);
  w.l.freq = f, w.d.freq = f + kHeap, w.b.freq = f + kHeap + 61;
  unsigned short* s = reinterpret_cast<unsigned short*>(tree_mem);
  w.l.dad = s, s += kHeap;
  w.d.dad = s, s += 61;
  w.b.dad = s, s += 39;
  w.l.code = s, s += kLCodes;
  w.d.code = s, s += kDCodes;
  w.b.code = s, s += kBLCodes;
  w.heap = s, s += kHeap;
  unsigned char* c = reinterpret_cast<unsigned char*>(s);
  w.l.len = c, c += kHeap + 1;
  w.d.len = c, c += 62;
  w.b.len = c, c += 40;
  w.depth = c;
  w.l.elems = kLCodes, w.l.kind = 0, w.l.max_length = 15;
  w.d.elems = kDCodes, w.d.kind = 1, w.d.max_length = 15;
  w.b.elems = kBLCodes, w.b.kind = 2, w.b.max_length = 7;
}

__host__ __device__ inline bool smaller(const Tree& t, const Work& w, int n, int m) {
  return t.freq[n] < t.freq[m] || (t.freq[n] == t.freq[m] && w.depth[n] <= w.depth[m]);
}
__host__ __device__ inline void pqdownheap(const Tree& t, Work& w, int k) {
  const int v = w.heap[k];
  int j = k << 1;
  while (j <= w.heap_len) {
    if (j < w.heap_len && smaller(t, w, w.heap[j + 1], w.heap[j])) j++;
    if (smaller(t, w, v, w.heap[j])) break;
    w.heap[k] = w.heap[j];
    k = j;
    j <<= 1;
  }
  w.heap[k] = (unsigned short)v;
}
__host__ __device__ inline int tree_xbits(int kind, int n) {
  return kind == 0 ? (n >= 257 ? extra_lbits(n - 257) : 0) : kind == 1 ? extra_dbits(n) : extra_blbits(n);
}
__host__ __device__ inline void gen_bitlen(Tree& t, Work& w) {
  const int max_code = t.max_code, max_length = t.max_length;
  int h, overflow = 0;
  for (int bits = 0; bits <= 15; ++bits) w.bl_count[bits] = 0;
  t.len[w.heap[w.heap_max]] = 0;
  for (h = w.heap_max + 1; h < kHeap; ++h) {
    const int n = w.heap[h];
    int bits = t.len[t.dad[n]] + 1;
    if (bits > max_length) bits = max_length, overflow++;
    t.len[n] = (unsigned char)bits;
    if (n > max_code) continue;
    w.bl_count[bits]++;
    const int xbits = tree_xbits(t.kind, n);
    const unsigned f = t.freq[n];
    w.opt_len += f * (unsigned)(bits + xbits);
    if (t.kind == 0) w.static_len += f * (unsigned)(static_llen(n) + xbits);
    if (t.kind == 1) w.static_len += f * (unsigned)(5 + xbits);
  }
  if (overflow == 0) return;
  do {
    int bits = max_length - 1;
    while (w.bl_count[bits] == 0) bits--;
    w.bl_count[bits]--;
    w.bl_count[bits + 1] += 2;
    w.bl_count[max_length]--;
    overflow -= 2;
  } while (overflow > 0);
  for (int bits = max_length; bits != 0; bits--) {
    int n = w.bl_count[bits];
    while (n != 0) {
      const int m = w.heap[--h];
      if (m > max_code) continue;
      if ((int)t.len[m] != bits) {
        w.opt_len += ((unsigned)bits - (unsigned)t.len[m]) * (unsigned)t.freq[m];
        t.len[m] = (unsigned char)bits;
      }
      n--;
    }
  }
}
__host__ __device__ inline void gen_codes(Tree& t, const Work& w) {
  unsigned short next_code[16];
  unsigned code = 0;
  next_code[0] = 0;
  for (int bits = 1; bits <= 15; ++bits) {
    code = (code + w.bl_count[bits - 1]) << 1;
    next_code[bits] = (unsigned short)code;
  }
  for (int n = 0; n <= t.max_code; ++n) {
    const int len = t.len[n];
    if (len == 0) continue;
    t.code[n] = (unsigned short)bi_reverse(next_code[len]++, len);
  }
}
__host__ __device__ inline void build_tree(Tree& t, Work& w) {
  const int elems = t.elems;
  int n, m, max_code = -1, node;
  w.heap_len = 0, w.heap_max = kHeap;
  for (n = 0; n < elems; ++n) {
    if (t.freq[n] != 0) {
      w.heap[++w.heap_len] = (unsigned short)(max_code = n);
      w.depth[n] = 0;
    } else {
      t.len[n] = 0;
    }
  }
  while (w.heap_len < 2) {
    node = w.heap[++w.heap_len] = (unsigned short)(max_code < 2 ? ++max_code : 0);
    t.freq[node] = 1;
    w.depth[node] = 0;
    w.opt_len--;
    if (t.kind == 0) w.static_len -= (unsigned)static_llen(node);
    if (t.kind == 1) w.static_len -= 5u;
  }
  t.max_code = max_code;
  for (n = w.heap_len / 2; n >= 1; --n) pqdownheap(t, w, n);
  node = elems;
  do {
    n = w.heap[1];
    w.heap[1] = w.heap[w.heap_len--];
    pqdownheap(t, w, 1);
    m = w.heap[1];
    w.heap[--w.heap_max] = (unsigned short)n;
    w.heap[--w.heap_max] = (unsigned short)m;
    t.freq[node] = (unsigned short)(t.freq[n] + t.freq[m]);
    w.depth[node] = (unsigned char)((w.depth[n] >= w.depth[m] ? w.depth[n] : w.depth[m]) + 1);
    t.dad[n] = t.dad[m] = (unsigned short)node;
    w.heap[1] = (unsigned short)node++;
    pqdownheap(t, w, 1);
  } while (w.heap_len >= 2);
  w.heap[--w.heap_max] = w.heap[1];
  gen_bitlen(t, w);
  gen_codes(t, w);
}
__host__ __device__ inline void scan_tree(const Tree& t, Work& w, int max_code) {
  int prevlen = -1, curlen, nextlen = t.len[0], count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) max_count = 138, min_count = 3;
  t.len[max_code + 1] = 0xff;  // guard
  for (int n = 0; n <= max_code; ++n) {
    curlen = nextlen, nextlen = t.len[n + 1];
    if (++count < max_count && curlen == nextlen) continue;
    if (count < min_count) {
      w.b.freq[curlen] += (unsigned short)count;
    } else if (curlen != 0) {
      if (curlen != prevlen) w.b.freq[curlen]++;
      w.b.freq[16]++;
    } else if (count <= 10) {
      w.b.freq[17]++;
    } else {
      w.b.freq[18]++;
    }
    count = 0, prevlen = curlen;
    if (nextlen == 0) max_count = 138, min_count = 3;
    else if (curlen == nextlen) max_count = 6, min_count = 3;
    else max_count = 7, min_count = 4;
  }
}

// LSB-first bit writer over zero-initialised 32-bit words (the staged output).  Serial use only.
struct BitW {
  unsigned* words;
  unsigned pos;     // in bits
  int in_global;    // the words are in global memory next to atomicOr traffic: go through the atomics too (L1 may be stale)
  __host__ __device__ inline void put(unsigned v, int nbits) {
    const unsigned wi = pos >> 5, sh = pos & 31;
#if defined(__CUDA_ARCH__)
    if (in_global) {
      atomicOr(words + wi, v << sh);
      if (sh + nbits > 32) atomicOr(words + wi + 1, v >> (32 - sh));
      pos += nbits;
      return;
    }
#endif
    words[wi] |= v << sh;
    if (sh + nbits > 32) words[wi + 1] |= v >> (32 - sh);
    pos += nbits;
  }
};
__host__ __device__ inline void send_tree(const Tree& t, const Work& w, int max_code, BitW& o) {
  int prevlen = -1, curlen, nextlen = t.len[0], count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) max_count = 138, min_count = 3;
  for (int n = 0; n <= max_code; ++n) {
    curlen = nextlen, nextlen = t.len[n + 1];
    if (++count < max_count && curlen == nextlen) continue;
    if (count < min_count) {
      do o.put(w.b.code[curlen], w.b.len[curlen]);
      while (--count != 0);
    } else if (curlen != 0) {
      if (curlen != prevlen) o.put(w.b.code[curlen], w.b.len[curlen]), count--;
      o.put(w.b.code[16], w.b.len[16]);
      o.put((unsigned)(count - 3), 2);
    } else if (count <= 10) {
      o.put(w.b.code[17], w.b.len[17]);
      o.put((unsigned)(count - 3), 3);
    } else {
      o.put(w.b.code[18], w.b.len[18]);
      o.put((unsigned)(count - 11), 7);
    }
    count = 0, prevlen = curlen;
    if (nextlen == 0) max_count = 138, min_count = 3;
    else if (curlen == nextlen) max_count = 6, min_count = 3;
    else max_count = 7, min_count = 4;
  }
}

// _tr_flush_block(last = 1) up to the point where the symbols are coded.  On entry l.freq / d.freq hold the symbol
// counts (END_BLOCK included), internal nodes and b.freq are cleared here.  Returns the block type (0 stored, 1 static,
// 2 dynamic); for 1 and 2 the 3 header bits (+ the tree header) are written at o and l/d code + len describe the
// code to use for every symbol.
__host__ __device__ inline int begin_block(Work& w, int stored_len, BitW& o, bool fill_static = true, int last = 1) {
  w.opt_len = 0, w.static_len = 0;
  for (int i = 0; i < kBLCodes; ++i) w.b.freq[i] = 0;
  build_tree(w.l, w);
  build_tree(w.d, w);
  scan_tree(w.l, w, w.l.max_code);
  scan_tree(w.d, w, w.d.max_code);
  build_tree(w.b, w);
  int max_blindex;
  for (max_blindex = kBLCodes - 1; max_blindex >= 3; max_blindex--)
    if (w.b.len[bl_order(max_blindex)] != 0) break;
  w.opt_len += 3 * ((unsigned)max_blindex + 1) + 5 + 5 + 4;
  unsigned opt_lenb = (w.opt_len + 3 + 7) >> 3;
  const unsigned static_lenb = (w.static_len + 3 + 7) >> 3;
  if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
  if ((unsigned)stored_len + 4 <= opt_lenb) return 0;
  if (static_lenb == opt_lenb) {
    o.put((1 << 1) + (unsigned)last, 3);
    if (!fill_static) return 1;  // the caller writes the fixed code (the kernel does it on all lanes)
    for (int n = 0; n < kLCodes; ++n) w.l.code[n] = (unsigned short)bi_reverse(static_lcode(n), static_llen(n)), w.l.len[n] = (unsigned char)static_llen(n);
    for (int n = 0; n < kDCodes; ++n) w.d.code[n] = (unsigned short)bi_reverse((unsigned)n, 5), w.d.len[n] = 5;
    return 1;
  }
  o.put((2 << 1) + (unsigned)last, 3);
  o.put((unsigned)(w.l.max_code + 1 - 257), 5);
  o.put((unsigned)(w.d.max_code + 1 - 1), 5);
  o.put((unsigned)(max_blindex + 1 - 4), 4);
  for (int rank = 0; rank <= max_blindex; ++rank) o.put(w.b.len[bl_order(rank)], 3);
  send_tree(w.l, w, w.l.max_code, o);
  send_tree(w.d, w, w.d.max_code, o);
  return 2;
}

// bits of one symbol under the block's code: value (LSB-first, <= 48 bits) and its length
__host__ __device__ inline unsigned long long symbol_bits(const Work& w, unsigned dist, unsigned lc, int& nbits) {
  if (dist == 0) {
    nbits = w.l.len[lc];
    return w.l.code[lc];
  }
  const int code = length_code((int)lc);
  unsigned long long v = w.l.code[257 + code];
  int nb = w.l.len[257 + code];
  const int e = extra_lbits(code);
  v |= (unsigned long long)(lc & ((1u << e) - 1u)) << nb;
  nb += e;
  const unsigned d0 = dist - 1;
  const int dc = dist_code((int)d0);
  v |= (unsigned long long)w.d.code[dc] << nb;
  nb += w.d.len[dc];
  const int ed = extra_dbits(dc);
  v |= (unsigned long long)(d0 & ((1u << ed) - 1u)) << nb;
  nb += ed;
  nbits = nb;
  return v;
}


// ---- device part: one warp per column --------------------------------------------------------------------------------
struct ColSeg {
  int rec_off;  // byte offset of the segment's first column inside a record
  int ncols;    // columns of this length in the segment (one factor matrix, fiber-major)
  int col0;     // index of its first column among all columns of an image
  int out_off;  // byte offset of its first output slot inside an image's block of the column buffer
};
struct Params {
  const unsigned char* rec;  // [batch][rec_stride] int8 records
  long long rec_stride;
  int len;             // column length of this launch
  int slot;            // output slot bytes per column (>= len + 16, multiple of 16)
  int n_seg;           // segments (<= 6)
  ColSeg seg[6];
  int cols_per_image;  // sum of seg[].ncols
  int cols_total;      // all columns of an image (all launches)
  int batch;
  unsigned char* cbuf;      // [batch][img_stride]
  long long img_stride;
  unsigned* csize;          // [batch][cols_total]
  unsigned char* scratch;   // [gridDim.x][scratch_per_cta(len)]
  int* counter;             // work queue, zeroed before the launch
};
__host__ __device__ inline int pad_len(int len) { return (len + 63) & ~63; }
__host__ __device__ inline long long scratch_per_cta(int len) { return 7ll * pad_len(len); }
__host__ __device__ inline int smem_data(int len) { return (len + 16 + 15) & ~15; }
constexpr int kFreqPad = (kFreqBytes + 15) & ~15, kCntBytes = 2 * (256 + 128), kTaskBytes = 64, kTreePad = (kTreeBytes + 15) & ~15;
__host__ __device__ inline int slot_bytes(int len) { return (len + 64 + 15) & ~15; }  // output slot of a column (stored blocks: +5 each)
// the sorted positions A (2 bytes each) share their region with the Huffman scratch and the staged output of a one-block
// column (both are needed only after the parse); a column of several blocks flushes blocks while A is alive, so its
// Huffman scratch sits behind A and its output is staged in the column's global slot
__host__ __device__ inline int smem_region(int len) {
  const int a = 2 * pad_len(len), b = kTreePad + slot_bytes(len);
  return len > kOneBlock ? a + kTreePad : (a > b ? a : b);
}
__host__ __device__ inline int smem_bytes(int len) { return smem_data(len) + kFreqPad + kTaskBytes + smem_region(len); }  // the sort's counters live in the frequency area

#ifdef D9_PROF
__device__ unsigned long long g_d9_prof[16];
#define D9_T(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_d9_prof[i], (unsigned long long)(t_ - t_prof)); t_prof = t_; } } while (0)
#define D9_C(i, v) do { if (threadIdx.x == 0) atomicAdd(&g_d9_prof[i], (unsigned long long)(v)); } while (0)
#else
#define D9_T(i)
#define D9_C(i, v)
#endif
constexpr int kWarps = 4;             // warps per long column: warp 0 runs the serial parts, every warp walks its share of long chains
constexpr int kWarpsHuge = 8;         // ... and per column so long that shared memory holds one or two of them per SM
#ifndef D9_LONG
#define D9_LONG 2048
#endif
constexpr int kLongColumn = D9_LONG;  // columns above this many bytes may get kWarps warps (small batches), the others one
#ifndef D9_WIDE
#define D9_WIDE 2
#endif
constexpr int kWide = D9_WIDE;              // chain candidates per lane and step
constexpr int kStep = 32 * kWide;     // candidates per step of one warp
constexpr int kParMin = 4 * kStep;    // chains with more candidates than this are walked by all warps of the CTA
struct Task {
  int p, best, maxlen, rank, left, cmd;  // cmd 1: walk, 2: the parse of this column is over
  unsigned key;                           // result: (length << 12) | (4095 - chain index), atomicMax over the warps
  int nice_idx;                           // smallest chain index that reached nice_length so far (walks stop beyond it)
  int stream;
};
__device__ __forceinline__ unsigned ld4(const unsigned* w, int off) {
  const int i = off >> 2;
  return __funnelshift_r(w[i], w[i + 1], (off & 3) * 8);
}
// Walk the chain candidates with index s0 = w*kStep, (w+nw)*kStep, ... < left (index 0 = newest = A[rank-1]) and return the
// longest match beyond `best` and the chain index of the closest candidate that has it.  A candidate can only beat
// `best` if it agrees with the string at p on bytes 0..best; filter on four of them: best-1, best (zlib's own quick
// check) and 3, 4 (right after the hashed trigram, where most chain members of these smooth columns part ways): ~4.5 %
// get through (12 % with zlib's two bytes alone).  Survivors are compared by the whole warp, 128 bytes per round, in chain
// order, so `best` tightens exactly as in the serial walk; the first candidate that reaches maxlen (nice_length or the
// end of the input) ends the walk — zlib stops there too — and candidates beyond it never matter because no length exceeds it.
__device__ __forceinline__ void walk_chain(const unsigned char* data, const unsigned* data32, const unsigned short* A,
                                           int p, int rank, int left, int maxlen, int w, int nw, int* nice_idx, int lane,
                                           int& best, int& bidx, const int limit) {
  // limit: zlib ends the walk at the first candidate farther than MAX_DIST; 0 (NIL) for every column of one block
  // filter offsets f1, f2 start at 3, 4 and then follow the survivors: where the last ones parted from the string at p
  // the next chain members usually do too (1 % get through instead of 2.7 %)
  unsigned ex = 0, exm = 0;
  int f1 = 3, f2 = 4;
  auto set_filter = [&]() {
    ex = (unsigned)data[p + best] | ((unsigned)data[p + best - 1] << 8) | ((unsigned)data[p + f1] << 16) | ((unsigned)data[p + f2] << 24);
    exm = 0xffffu | (best >= f1 ? 0xff0000u : 0u) | (best >= f2 ? 0xff000000u : 0u);
  };
  auto filter = [&](int q) {
    const unsigned char* dq = data + q;
    const unsigned c = (unsigned)dq[best] | ((unsigned)dq[best - 1] << 8) | ((unsigned)dq[f1] << 16) | ((unsigned)dq[f2] << 24);
    return ((c ^ ex) & exm) == 0u;
  };
  if (w * kStep >= left) return;
  set_filter();
  bool dirty = false;
  // the string at p, 4 bytes per lane (first 128 bytes), with the bytes beyond maxlen masked off
  const int nv0 = maxlen - 4 * lane;
  const unsigned lm = nv0 >= 4 ? 0xffffffffu : nv0 <= 0 ? 0u : (1u << (8 * nv0)) - 1u;
  const unsigned Pw = lm ? ld4(data32, p + 4 * lane) : 0u;
  for (int s0 = w * kStep; s0 < left; s0 += nw * kStep) {
    if (nice_idx && s0 > *reinterpret_cast<volatile int*>(nice_idx)) break;  // a closer candidate already ended the walk
    if (dirty) set_filter(), dirty = false;
    D9_C(12, 1);
    int q[kWide];
    unsigned pm[kWide];
#pragma unroll
    for (int j = 0; j < kWide; ++j) {
      const int idx = s0 + 32 * j + lane;
      q[j] = idx < left ? A[rank - 1 - idx] : 0;
    }
#pragma unroll
    for (int j = 0; j < kWide; ++j) pm[j] = __ballot_sync(0xffffffffu, q[j] > limit && filter(q[j]));
    bool done = false;
#pragma unroll
    for (int j = 0; j < kWide; ++j) {
      unsigned mset = pm[j];
      while (mset) {
        const int src = __ffs((int)mset) - 1;
        mset &= mset - 1;
        const int qq = __shfl_sync(0xffffffffu, q[j], src);
        D9_C(13, 1);
        unsigned x = 0;
        if (lm) x = (Pw ^ ld4(data32, qq + 4 * lane)) & lm;
        int len = __reduce_min_sync(0xffffffffu, x ? 4 * lane + ((__ffs((int)x) - 1) >> 3) : 0x7fff);
        if (len == 0x7fff) {
          len = maxlen;
          for (int base = 128; base < maxlen; base += 128) {  // matches beyond 128 bytes: up to two more rounds
            const int off = base + 4 * lane, nv = maxlen - off;
            unsigned y = 0;
            if (nv > 0) {
              y = ld4(data32, p + off) ^ ld4(data32, qq + off);
              if (nv < 4) y &= (1u << (8 * nv)) - 1u;
            }
            const int mn = __reduce_min_sync(0xffffffffu, y ? off + ((__ffs((int)y) - 1) >> 3) : 0x7fff);
            if (mn != 0x7fff) {
              len = mn;
              break;
            }
          }
        }
        if (len > best) {
          best = len, bidx = s0 + 32 * j + src;
          D9_C(14, 1);
          if (best >= maxlen) {
            done = true;
            break;
          }
        } else {
          f2 = f1, f1 = len;  // this survivor parted from the string at offset len <= best
        }
        dirty = true;
        bool more = mset != 0;
#pragma unroll
        for (int j2 = j + 1; j2 < kWide; ++j2) more |= pm[j2] != 0;
        if (more) {  // the rest of this step is re-filtered against the new best / the new offsets
          set_filter(), dirty = false;
          mset = __ballot_sync(0xffffffffu, lane > src && q[j] > limit && filter(q[j]));
#pragma unroll
          for (int j2 = j + 1; j2 < kWide; ++j2) pm[j2] = __ballot_sync(0xffffffffu, q[j2] > limit && filter(q[j2]));
        }
      }
      if (done) break;
    }
    if (done) {
      if (nice_idx && lane == 0) atomicMin(nice_idx, bidx);
      break;
    }
  }
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// _tr_flush_block: trees and header on lane 0, symbols on all lanes (prefix sum of code lengths, atomicOr into the output),
// or LEN / NLEN and the bytes for a stored block; then init_block.  Out of line: the parse loop calls it from three places.
struct BlockOut {
  unsigned* words;         // zero-initialised output (shared for one-block columns, the global slot otherwise)
  Work* w;
  unsigned char* freq_mem;
  const unsigned char* data;
  const unsigned short* sym_d;
  const unsigned char* sym_l;
  unsigned bitpos;
  int block_start;
  int in_global;
};
__device__ __forceinline__ void flush_block_body(BlockOut& b, int end, int last, int ns, int lane) {
  Work& w = *b.w;
  unsigned* words = b.words;
  unsigned bitpos = b.bitpos;
  const int stored_len = end - b.block_start;
  int type = 0;
  if (lane == 0) {
    w.l.freq[256] = 1;
    BitW o{words, bitpos, b.in_global};
    type = begin_block(w, stored_len, o, false, last);
    if (type == 0) o.put((unsigned)last, 3), o.pos = (o.pos + 7u) & ~7u;
    bitpos = o.pos;
  }
  type = __shfl_sync(0xffffffffu, type, 0);
  bitpos = __shfl_sync(0xffffffffu, bitpos, 0);
  if (type == 1) {  // fixed Huffman code
    for (int c = lane; c < kLCodes; c += 32) w.l.code[c] = (unsigned short)bi_reverse(static_lcode(c), static_llen(c)), w.l.len[c] = (unsigned char)static_llen(c);
    if (lane < kDCodes) w.d.code[lane] = (unsigned short)bi_reverse((unsigned)lane, 5), w.d.len[lane] = 5;
  }
  __syncwarp();
  if (type == 0) {  // stored block: LEN, NLEN, the bytes
    unsigned char* ob = reinterpret_cast<unsigned char*>(words) + (bitpos >> 3);
    if (lane == 0) ob[0] = (unsigned char)stored_len, ob[1] = (unsigned char)(stored_len >> 8), ob[2] = (unsigned char)~stored_len, ob[3] = (unsigned char)(~stored_len >> 8);
    for (int i = lane; i < stored_len; i += 32) ob[4 + i] = b.data[b.block_start + i];
    bitpos += 8u * (4u + (unsigned)stored_len);
  } else {
    for (int base = 0; base <= ns; base += 32) {
      const int i = base + lane;
      int nb = 0;
      unsigned long long v = 0;
      if (i < ns) v = symbol_bits(w, __ldcg(b.sym_d + i), __ldcg(b.sym_l + i), nb);
      else if (i == ns) v = symbol_bits(w, 0, 256, nb);
      const int inc = warp_incl_scan(nb, lane);
      if (nb) {
        const unsigned pos = bitpos + (unsigned)(inc - nb), wi = pos >> 5, sh = pos & 31;
        atomicOr(words + wi, (unsigned)(v << sh));
        const unsigned long long hi = sh ? (v >> (32 - sh)) : (v >> 16 >> 16);
        if (hi) {
          atomicOr(words + wi + 1, (unsigned)hi);
          if (hi >> 32) atomicOr(words + wi + 2, (unsigned)(hi >> 32));
        }
      }
      bitpos += (unsigned)__shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncwarp();
  if (!last) {  // init_block
    for (int i = lane; i < kFreqPad / 4; i += 32) reinterpret_cast<unsigned*>(b.freq_mem)[i] = 0;
    __syncwarp();
  }
  b.bitpos = bitpos, b.block_start = end;
}
__device__ __noinline__ void flush_block_out(BlockOut& b, int end, int last, int ns, int lane) { flush_block_body(b, end, last, ns, lane); }

// NW warps per column: 4 shorten a long column's critical path (shared chain walks), 1 puts more columns on an SM.
// MULTI: columns above kOneBlock bytes (several deflate blocks, flushed in mid-parse).
template <int NW, bool MULTI>
__global__ void __launch_bounds__(32 * NW, NW == 1 ? 32 : NW <= 4 ? 8 : 2) deflate9_kernel(Params P) {
  LRFB_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = P.len, m = n >= 3 ? n - 2 : 0, lp = pad_len(n);
  unsigned char* data = smem;
  const unsigned* data32 = reinterpret_cast<const unsigned*>(smem);
  unsigned char* freq_mem = smem + smem_data(n);
  static_assert(kCntBytes <= kFreqPad, "the sort's counters borrow the frequency area");
  unsigned short* cnt = reinterpret_cast<unsigned short*>(freq_mem);  // [256] low digit, [128] high digit; dead before the parse
  Task* task = reinterpret_cast<Task*>(freq_mem + kFreqPad);
  unsigned char* region = freq_mem + kFreqPad + kTaskBytes;
  unsigned short* A = reinterpret_cast<unsigned short*>(region);  // positions sorted by (hash, position)
  unsigned char* scr = P.scratch + (long long)blockIdx.x * scratch_per_cta(n);
  unsigned short* B = reinterpret_cast<unsigned short*>(scr);                // sort ping-pong, dead once RC is written
  unsigned* RC = reinterpret_cast<unsigned*>(scr);                           // per position: rank | chain candidates << 16
  unsigned short* sym_d = reinterpret_cast<unsigned short*>(scr + 4 * lp);   // match distance, 0 = literal
  unsigned char* sym_l = scr + 6 * lp;                                       // literal byte / match length - 3
  const int n_streams = P.batch * P.cols_per_image;

  for (;;) {
    if (threadIdx.x == 0) task->stream = atomicAdd(P.counter, 1);
    __syncthreads();
    const int s = task->stream;
    if (s >= n_streams) break;
    if (warp != 0) {
      // helper warps: wait for the long chain walks warp 0 publishes, take every kWarps-th step of them
      for (;;) {
        __syncthreads();
        if (*reinterpret_cast<volatile int*>(&task->cmd) == 2) break;
        const int tp = task->p, tmax = task->maxlen, trank = task->rank, tleft = task->left;
        int best = task->best, bidx = -1;
        walk_chain(data, data32, A, tp, trank, tleft, tmax, warp, NW, &task->nice_idx, lane, best, bidx,
                   MULTI && tp > kMaxDist ? tp - kMaxDist : 0);
        if (bidx >= 0 && lane == 0) atomicMax(&task->key, ((unsigned)best << 12) | (unsigned)(4095 - bidx));
        __syncthreads();
      }
      continue;
    }
    // hand out the columns of high rank index first: they carry less signal (long zero runs, long hash chains) and take
    // several times longer than the leading columns, which then fill the tail of the launch
    const int img = s % P.batch;
    int j = P.cols_per_image - 1 - s / P.batch, sg = 0;
    while (sg + 1 < P.n_seg && j >= P.seg[sg].ncols) j -= P.seg[sg].ncols, ++sg;
    const unsigned char* src = P.rec + (long long)img * P.rec_stride + P.seg[sg].rec_off + (long long)j * n;
    unsigned char* dst = P.cbuf + (long long)img * P.img_stride + P.seg[sg].out_off + (long long)j * P.slot;
    unsigned* dsize = P.csize + (long long)img * P.cols_total + P.seg[sg].col0 + j;

#ifdef D9_PROF
    long long t_prof = clock64();
#endif
    // ---- load the column, clear the counters --------------------------------------------------------------------
    if ((((unsigned long long)src) & 15) == 0 && (n & 15) == 0) {
      for (int i = lane * 16; i < n; i += 512) *reinterpret_cast<uint4*>(data + i) = *reinterpret_cast<const uint4*>(src + i);
    } else {
      for (int i = lane; i < n; i += 32) data[i] = src[i];
    }
    for (int i = n + lane; i < smem_data(n); i += 32) data[i] = 0;
    for (int i = lane; i < kCntBytes / 4; i += 32) reinterpret_cast<unsigned*>(freq_mem)[i] = 0;
    __syncwarp();

    // ---- positions sorted by (hash, position): two stable counting passes (8 + 7 bits) ----------------------------
    for (int base = 0; base < m; base += 32) {
      const int i = base + lane;
      const bool valid = i < m;
      const unsigned h = valid ? hash3(data, i) : 0;
      const unsigned k0 = valid ? (h & 255u) : 0x10000u + lane, k1 = valid ? (h >> 8) : 0x10000u + lane;
      const unsigned p0 = __match_any_sync(0xffffffffu, k0), p1 = __match_any_sync(0xffffffffu, k1);
      if (valid && (__ffs((int)p0) - 1) == lane) cnt[k0] += (unsigned short)__popc(p0);
      if (valid && (__ffs((int)p1) - 1) == lane) cnt[256 + k1] += (unsigned short)__popc(p1);
      __syncwarp();
    }
    {  // exclusive prefix sums of the two histograms
      int run = 0;
      for (int b0 = 0; b0 < 256; b0 += 32) {
        const int v = cnt[b0 + lane], inc = warp_incl_scan(v, lane);
        cnt[b0 + lane] = (unsigned short)(run + inc - v);
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      run = 0;
      for (int b0 = 0; b0 < 128; b0 += 32) {
        const int v = cnt[256 + b0 + lane], inc = warp_incl_scan(v, lane);
        cnt[256 + b0 + lane] = (unsigned short)(run + inc - v);
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
    }
    for (int base = 0; base < m; base += 32) {  // pass 1: identity -> B by the low 8 bits
      const int i = base + lane;
      const bool valid = i < m;
      const unsigned k0 = valid ? (hash3(data, i) & 255u) : 0x10000u + lane;
      const unsigned p0 = __match_any_sync(0xffffffffu, k0);
      const int off = valid ? cnt[k0] : 0;
      __syncwarp();
      if (valid) {
        __stcg(B + off + __popc(p0 & ((1u << lane) - 1u)), (unsigned short)i);
        if ((__ffs((int)p0) - 1) == lane) cnt[k0] = (unsigned short)(off + __popc(p0));
      }
      __syncwarp();
    }
    {
      unsigned short qn = lane < m ? __ldcg(B + lane) : (unsigned short)0;
      for (int base = 0; base < m; base += 32) {  // pass 2: B -> A by the high 7 bits
        const int i = base + lane;
        const bool valid = i < m;
        const int q = qn;
        if (i + 32 < m) qn = __ldcg(B + i + 32);
        const unsigned k1 = valid ? (hash3(data, q) >> 8) : 0x10000u + lane;
        const unsigned p1 = __match_any_sync(0xffffffffu, k1);
        const int off = valid ? cnt[256 + k1] : 0;
        __syncwarp();
        if (valid) {
          A[off + __popc(p1 & ((1u << lane) - 1u))] = (unsigned short)q;
          if ((__ffs((int)p1) - 1) == lane) cnt[256 + k1] = (unsigned short)(off + __popc(p1));
        }
        __syncwarp();
      }
    }
    D9_T(0);  // load + sort
    {  // per position: its rank in A and the number of chain candidates below it (same hash, position 0 excluded: NIL)
      int run_start = 0, kzero = -1, carry_h = -1;
      for (int base = 0; base < m; base += 32) {
        const int k = base + lane;
        const bool valid = k < m;
        const int q = valid ? A[k] : 0;
        const int hk = valid ? (int)hash3(data, q) : -2;
        int hp = __shfl_up_sync(0xffffffffu, hk, 1);
        if (lane == 0) hp = carry_h;
        carry_h = __shfl_sync(0xffffffffu, hk, 31);
        const unsigned zb = __ballot_sync(0xffffffffu, valid && q == 0);
        if (zb) kzero = base + __ffs((int)zb) - 1;
        int st = (valid && hk != hp) ? k : -1;  // bucket starts, then an inclusive max-scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, st, o);
          if (lane >= o) st = st > t ? st : t;
        }
        if (st < 0) st = run_start;
        run_start = __shfl_sync(0xffffffffu, st, 31);
        if (valid) {
          int cl = k - st - (st == kzero ? 1 : 0);
          if (cl < 0) cl = 0;
          __stcg(RC + q, (unsigned)k | ((unsigned)cl << 16));
        }
      }
    }
    __syncwarp();

    for (int i = lane; i < kFreqPad / 4; i += 32) reinterpret_cast<unsigned*>(freq_mem)[i] = 0;  // symbol frequencies
    __syncwarp();
    D9_T(1);  // ranks
    // ---- deflate_slow ------------------------------------------------------------------------------------------------
    unsigned short* lfreq = reinterpret_cast<unsigned short*>(freq_mem);
    unsigned short* dfreq = lfreq + kHeap;
    int strstart = 0, lookahead = n, match_length = 2, match_start = 0, match_available = 0, ns = 0;
    // ---- block output: _tr_flush_block.  One block per column up to kOneBlock bytes; longer columns flush whenever the
    // symbol buffer holds 16 383 entries, as zlib does, with the Huffman scratch behind A and the output in the global slot
    constexpr bool multi = MULTI;
    unsigned* words = multi ? reinterpret_cast<unsigned*>(dst) : reinterpret_cast<unsigned*>(region + kTreePad);
    Work w;
    work_bind(w, freq_mem, multi ? region + 2 * lp : region);
    const unsigned bitpos = 16;
    if (multi) {
      for (int i = lane; i < slot_bytes(n) / 4; i += 32) words[i] = i == 0 ? 0xDA78u : 0u;
      __syncwarp();
    }
    BlockOut bo{words, &w, freq_mem, data, sym_d, sym_l, bitpos, 0, multi ? 1 : 0};
    auto flush_block = [&](int end, int last) {
      if (MULTI) flush_block_out(bo, end, last, ns, lane);  // three call sites in the parse loop
      else flush_block_body(bo, end, last, ns, lane);
      ns = 0;
    };
    // rank | candidates of 128 positions ahead of the parse live in four registers per lane; the far half is
    // requested from L2 64 positions before it is needed
    int wbase = -1000000;
    unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    auto ldrc = [&](int i) { return i < m ? __ldcg(RC + i) : 0u; };
    while (lookahead > 0) {
      const int prev_length = match_length, prev_match = match_start;
      match_length = 2;
      const int maxlen = lookahead < 258 ? lookahead : 258;
      if (lookahead >= 3 && prev_length < maxlen) {
        const int p = strstart;
        if (p < wbase || p - wbase >= 128) {
          wbase = p & ~31;
          w0 = ldrc(wbase + lane), w1 = ldrc(wbase + 32 + lane), w2 = ldrc(wbase + 64 + lane), w3 = ldrc(wbase + 96 + lane);
        } else if (p - wbase >= 64) {
          wbase += 64;
          w0 = w2, w1 = w3;
          w2 = ldrc(wbase + 64 + lane), w3 = ldrc(wbase + 96 + lane);
        }
        const int wi = (p - wbase) >> 5;
        const unsigned rc = __shfl_sync(0xffffffffu, wi == 0 ? w0 : wi == 1 ? w1 : wi == 2 ? w2 : w3, (p - wbase) & 31);
        // zlib walks at most max_chain (4096; a quarter once the previous match is >= good_length) candidates, newest
        // first; here they are A[rank-1], A[rank-2], ...
        const int rank = (int)(rc & 0xffffu);
        int left = (int)(rc >> 16);
        {
          const int chain = prev_length >= 32 ? 1024 : 4096;
          left = left < chain ? left : chain;
        }
        int best = prev_length, bidx = -1;
        D9_T(2);  // parse control
        if (NW > 1 && left > kParMin) {  // long chain: every warp of the CTA takes every NW-th step; longest-then-closest by atomicMax
          if (lane == 0) {
            task->p = p, task->best = best, task->maxlen = maxlen, task->rank = rank, task->left = left, task->cmd = 1;
            task->key = 0, task->nice_idx = 0x7fffffff;
          }
          __syncthreads();
          walk_chain(data, data32, A, p, rank, left, maxlen, 0, NW, &task->nice_idx, lane, best, bidx, MULTI && p > kMaxDist ? p - kMaxDist : 0);
          if (bidx >= 0 && lane == 0) atomicMax(&task->key, ((unsigned)best << 12) | (unsigned)(4095 - bidx));
          __syncthreads();
          const unsigned key = *reinterpret_cast<volatile unsigned*>(&task->key);
          bidx = -1;
          if (key) best = (int)(key >> 12), bidx = 4095 - (int)(key & 4095u);
          D9_T(3);  // long walks
          D9_C(8, 1);
          D9_C(9, left);
        } else {
          walk_chain(data, data32, A, p, rank, left, maxlen, 0, 1, nullptr, lane, best, bidx, MULTI && p > kMaxDist ? p - kMaxDist : 0);
          D9_T(4);  // short walks
          D9_C(10, 1);
          D9_C(11, left);
        }
        match_length = best;
        if (bidx >= 0) match_start = A[rank - 1 - bidx];
        if (match_length == 3 && p - match_start > 4096) match_length = 2;
      }
      if (prev_length >= 3 && match_length <= prev_length) {
        if (lane == 0) {
          const unsigned dist = (unsigned)(strstart - 1 - prev_match), lc = (unsigned)(prev_length - 3);
          __stcg(sym_d + ns, (unsigned short)dist), __stcg(sym_l + ns, (unsigned char)lc);
          lfreq[257 + length_code((int)lc)]++, dfreq[dist_code((int)dist - 1)]++;
        }
        ++ns;
        lookahead -= prev_length - 1, strstart += prev_length - 1;
        match_available = 0, match_length = 2;
        if (MULTI && ns == kBlockSymbols) {
          __syncwarp();
          flush_block(strstart, 0);
        }
      } else if (match_available) {
        if (lane == 0) {
          const unsigned c = data[strstart - 1];
          __stcg(sym_d + ns, (unsigned short)0), __stcg(sym_l + ns, (unsigned char)c);
          lfreq[c]++;
        }
        ++ns;
        if (MULTI && ns == kBlockSymbols) {  // zlib flushes before it steps over the pending byte
          __syncwarp();
          flush_block(strstart, 0);
        }
        ++strstart, --lookahead;
      } else {
        match_available = 1, ++strstart, --lookahead;
      }
    }
    if (match_available) {
      if (lane == 0) {
        const unsigned c = data[strstart - 1];
        __stcg(sym_d + ns, (unsigned short)0), __stcg(sym_l + ns, (unsigned char)c);
        lfreq[c]++;
      }
      ++ns;
    }
    D9_T(2);
    if (lane == 0) task->cmd = 2;
    __syncthreads();  // the helper warps leave; A may now be overwritten

    // ---- the last block (trees and header on lane 0, symbols on all lanes), adler32, copy out -------------------------
    if (!multi) {
      for (int i = lane; i < slot_bytes(n) / 4; i += 32) words[i] = i == 0 ? 0xDA78u : 0u;
      __syncwarp();
    }
    flush_block(strstart, 1);
    D9_T(5);  // trees + header + symbols
    unsigned long long sa = 0, sb = 0;  // adler32: a = 1 + sum d_i, b = n + sum (n - i) d_i
    for (int i = lane; i < n; i += 32) sa += data[i], sb += (unsigned long long)(n - i) * data[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o), sb += __shfl_xor_sync(0xffffffffu, sb, o);
    const unsigned ad_a = (unsigned)((1 + sa) % 65521ull), ad_b = (unsigned)((n + sb) % 65521ull);
    int total = (int)((bo.bitpos + 7) >> 3);
    if (lane == 0) {
      unsigned char* ob = reinterpret_cast<unsigned char*>(words);
      ob[total] = (unsigned char)(ad_b >> 8), ob[total + 1] = (unsigned char)ad_b, ob[total + 2] = (unsigned char)(ad_a >> 8), ob[total + 3] = (unsigned char)ad_a;
    }
    total += 4;
    __syncwarp();
    if (!multi)
      for (int i = lane; i < (total + 3) / 4; i += 32) reinterpret_cast<unsigned*>(dst)[i] = words[i];
    D9_T(6);  // symbols + copy out
    if (lane == 0) *dsize = (unsigned)total;
    __syncwarp();
  }
}


// ---- framing: encode_matrix / combine_bytes / metadata header of one image (lrf/compression/utils.py:246-300, :354-390,
// lrf/compression/qmf.py:288-292).  combine_bytes(parts) = k - 1 nested BE32 length prefixes (outermost first), then the
// parts; every matrix is combine(json header, combine_bytes(compressed columns)). ------------------------------------------
struct FrameParams {
  int n_mat;           // 2 * planes: U_0, V_0, U_1, ...
  int ncols[6];        // columns (= rank) per matrix
  int col0[6];         // index of the matrix's first column among all columns of an image
  int slot_off[6];     // byte offset of its first slot in the image's column-buffer block
  int slot[6];         // slot bytes per column
  int hdr_len[6];
  char hdr[6][64];     // {"num_fibers": R, "mode": "col", "dtype": "int8"}
  int meta_len;
  char meta[1024];     // the image metadata json
  int cols_total;
  int batch;
  const unsigned char* cbuf;
  long long img_stride;
  const unsigned* csize;
  long long* sizes;    // [batch] framed bytes per image (sizes kernel), input of the scan
  long long* offsets;  // [batch + 1]
  unsigned char* blob;
  long long capacity;
};
__host__ __device__ inline long long matrix_bytes(const FrameParams& F, const unsigned* cs, int mtx) {
  long long b = 4 + F.hdr_len[mtx] + 4ll * (F.ncols[mtx] - 1);
  for (int r = 0; r < F.ncols[mtx]; ++r) b += cs[F.col0[mtx] + r];
  return b;
}
__global__ void __launch_bounds__(128) frame_sizes_kernel(FrameParams F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F.batch) return;
  const unsigned* cs = F.csize + (long long)i * F.cols_total;
  long long body = 4ll * (F.n_mat - 1);
  for (int mtx = 0; mtx < F.n_mat; ++mtx) body += matrix_bytes(F, cs, mtx);
  F.sizes[i] = 4 + F.meta_len + body;
}
// exclusive prefix sum of sizes -> offsets[0..batch]; one CTA
__global__ void __launch_bounds__(1024) frame_scan_kernel(const long long* sizes, long long* offsets, int batch) {
  __shared__ long long part[1024];
  const int t = threadIdx.x, per = (batch + (int)blockDim.x - 1) / (int)blockDim.x;
  const int lo = t * per < batch ? t * per : batch, hi = lo + per < batch ? lo + per : batch;
  long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += sizes[i];
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int k = 0; k < (int)blockDim.x; ++k) {
      const long long v = part[k];
      part[k] = run, run += v;
    }
    offsets[batch] = run;
  }
  __syncthreads();
  long long run = part[t];
  for (int i = lo; i < hi; ++i) offsets[i] = run, run += sizes[i];
}
__device__ inline void put_be32(unsigned char* p, long long v) {
  p[0] = (unsigned char)(v >> 24), p[1] = (unsigned char)(v >> 16), p[2] = (unsigned char)(v >> 8), p[3] = (unsigned char)v;
}
constexpr int kFrameThreads = 256, kMaxJobs = 6 * 64 + 8;
__global__ void __launch_bounds__(kFrameThreads) frame_write_kernel(FrameParams F) {
  __shared__ int job_dst[kMaxJobs], job_src[kMaxJobs], job_len[kMaxJobs], n_jobs;  // src: offset into the image's cbuf block
  const int i = blockIdx.x, t = threadIdx.x;
  const long long o0 = F.offsets[i], o1 = F.offsets[i + 1];
  if (o1 > F.capacity) return;  // the host checks offsets[batch] against the capacity
  unsigned char* out = F.blob + o0;
  const unsigned* cs = F.csize + (long long)i * F.cols_total;
  if (t == 0) {
    int nj = 0;
    long long pos = 0;
    put_be32(out, F.meta_len), pos = 4 + F.meta_len;
    // body = combine_bytes(matrices): prefixes acc[k-2] ... acc[0], acc[j] = 4 + acc[j-1] + size_j
    long long msize[6], acc = 0;
    for (int mtx = 0; mtx < F.n_mat; ++mtx) msize[mtx] = matrix_bytes(F, cs, mtx);
    long long accs[6];
    for (int mtx = 0; mtx < F.n_mat; ++mtx) accs[mtx] = acc = (mtx ? 4 + acc : 0) + msize[mtx];
    for (int mtx = F.n_mat - 2; mtx >= 0; --mtx) put_be32(out + pos, accs[mtx]), pos += 4;
    for (int mtx = 0; mtx < F.n_mat; ++mtx) {
      put_be32(out + pos, F.hdr_len[mtx]), pos += 4;
      for (int c = 0; c < F.hdr_len[mtx]; ++c) out[pos + c] = (unsigned char)F.hdr[mtx][c];
      pos += F.hdr_len[mtx];
      const int R = F.ncols[mtx];
      // columns: prefixes of the running sums, outermost first
      long long run = 0;
      for (int r = 0; r < R - 1; ++r) run = (r ? 4 + run : 0) + cs[F.col0[mtx] + r];  // acc[R-2]
      for (int r = R - 2; r >= 0; --r) {
        put_be32(out + pos, run), pos += 4;
        run -= cs[F.col0[mtx] + r] + (r ? 4 : 0);
      }
      for (int r = 0; r < R; ++r) {
        job_dst[nj] = (int)pos, job_src[nj] = F.slot_off[mtx] + r * F.slot[mtx], job_len[nj] = (int)cs[F.col0[mtx] + r];
        pos += cs[F.col0[mtx] + r], ++nj;
      }
    }
    n_jobs = nj;
  }
  for (int c = t; c < F.meta_len; c += kFrameThreads) out[4 + c] = (unsigned char)F.meta[c];
  __syncthreads();
  const unsigned char* src = F.cbuf + (long long)i * F.img_stride;
  for (int jb = 0; jb < n_jobs; ++jb) {
    const unsigned char* sp = src + job_src[jb];
    unsigned char* dp = out + job_dst[jb];
    for (int c = t; c < job_len[jb]; c += kFrameThreads) dp[c] = sp[c];
  }
}

#ifdef LRFB_SIM
// ---- test tooling (CPU shim build only): the textbook serial form of the same stream, used by tests/test_deflate9.py
// to check the shared tree / header code against zlib on thousands of inputs in milliseconds ----------------------
inline long long deflate9_serial(const unsigned char* in, int n, unsigned char* out, long long* probes) {
  std::vector<unsigned char> win(n + 300, 0);
  if (n) memcpy(win.data(), in, n);
  std::vector<unsigned short> head(32768, 0), prev(32768, 0), sd(n + 1);
  std::vector<unsigned char> sl(n + 1);
  std::vector<unsigned char> fm(kFreqBytes, 0), tm(kTreeBytes, 0);
  Work w;
  work_bind(w, fm.data(), tm.data());
  std::vector<unsigned> words((n + 256) / 4 + 8, 0);
  BitW o{words.data(), 16, 0};
  words[0] = 0xDA78u;
  int ns = 0, block_start = 0;
  auto flush_block = [&](int end, int last) {  // _tr_flush_block(window + block_start, end - block_start, last)
    w.l.freq[256] = 1;
    const int stored_len = end - block_start;
    const int type = begin_block(w, stored_len, o, true, last);
    if (type == 0) {
      o.put((unsigned)last, 3);
      o.pos = (o.pos + 7) & ~7u;
      unsigned char* ob = reinterpret_cast<unsigned char*>(words.data()) + (o.pos >> 3);
      ob[0] = (unsigned char)stored_len, ob[1] = (unsigned char)(stored_len >> 8), ob[2] = (unsigned char)~stored_len, ob[3] = (unsigned char)(~stored_len >> 8);
      if (stored_len) memcpy(ob + 4, in + block_start, stored_len);
      o.pos += 8u * (4 + stored_len);
    } else {
      for (int i = 0; i <= ns; ++i) {
        int nb;
        const unsigned long long v = i < ns ? symbol_bits(w, sd[i], sl[i], nb) : symbol_bits(w, 0, 256, nb);
        o.put((unsigned)v, nb > 32 ? 32 : nb);
        if (nb > 32) o.put((unsigned)(v >> 32), nb - 32);
      }
    }
    memset(fm.data(), 0, fm.size());
    ns = 0, block_start = end;
  };
  auto tally = [&](unsigned dist, unsigned lc) {
    sd[ns] = (unsigned short)dist, sl[ns] = (unsigned char)lc, ++ns;
    if (dist == 0) w.l.freq[lc]++;
    else w.l.freq[length_code((int)lc) + 257]++, w.d.freq[dist_code((int)dist - 1)]++;
    return ns == kBlockSymbols;
  };
  int strstart = 0, lookahead = n, match_length = 2, prev_length, match_start = 0, prev_match, match_available = 0;
  long long np = 0;
  while (lookahead > 0) {
    int hash_head = 0;
    if (lookahead >= 3) {
      const unsigned h = hash3(win.data(), strstart);
      hash_head = prev[strstart & 32767] = head[h];
      head[h] = (unsigned short)strstart;
    }
    prev_length = match_length, prev_match = match_start, match_length = 2;
    if (hash_head != 0 && prev_length < 258 && strstart - hash_head <= kMaxDist) {
      unsigned chain = prev_length >= 32 ? 1024 : 4096;
      int best = prev_length, nice = std::min(258, lookahead), cur = hash_head;
      const int limit = strstart > kMaxDist ? strstart - kMaxDist : 0;
      do {
        ++np;
        int len = 0;
        while (len < 258 && win[cur + len] == win[strstart + len]) ++len;
        if (len > best) {
          match_start = cur, best = len;
          if (len >= nice) break;
        }
      } while ((cur = prev[cur & 32767]) > limit && --chain != 0);
      match_length = std::min(best, lookahead);
      if (match_length == 3 && strstart - match_start > 4096) match_length = 2;
    }
    if (prev_length >= 3 && match_length <= prev_length) {
      const int max_insert = strstart + lookahead - 3;
      const bool bflush = tally((unsigned)(strstart - 1 - prev_match), (unsigned)(prev_length - 3));
      lookahead -= prev_length - 1;
      prev_length -= 2;
      do {
        if (++strstart <= max_insert) {
          const unsigned h = hash3(win.data(), strstart);
          prev[strstart & 32767] = head[h];
          head[h] = (unsigned short)strstart;
        }
      } while (--prev_length != 0);
      match_available = 0, match_length = 2, strstart++;
      if (bflush) flush_block(strstart, 0);
    } else if (match_available) {
      if (tally(0, win[strstart - 1])) flush_block(strstart, 0);
      strstart++, lookahead--;
    } else {
      match_available = 1, strstart++, lookahead--;
    }
  }
  if (match_available) tally(0, win[strstart - 1]);
  if (probes) *probes = np;
  flush_block(strstart, 1);
  long long nbytes = (o.pos + 7) >> 3;
  memcpy(out, words.data(), nbytes);
  unsigned a = 1, b = 0;
  for (int i = 0; i < n; ++i) a = (a + in[i]) % 65521u, b = (b + a) % 65521u;
  out[nbytes] = (unsigned char)(b >> 8), out[nbytes + 1] = (unsigned char)b, out[nbytes + 2] = (unsigned char)(a >> 8), out[nbytes + 3] = (unsigned char)a;
  return nbytes + 4;
}
#endif

}  // namespace d9
}  // namespace lrfb
