/*
 * oracle/myyuv_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
 *
 * A plain-C CPU restatement of the reference's (mahbhlddnhakkh/yuv-manipulations-2) hot path:
 * XRGB8888 -> IYUV colour conversion, DCT-q compression, decompression and the compressed payload
 * layout.  It is the checker the CUDA path is compared against; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it.  The product (yuv-manipulations-2_b200/) never does.
 *
 * PARITY STATUS: PINNED.  tests/test_oracle.py checks this file byte-for-byte against the reference's
 * own golden files (images/chef-with-trumpet{.myyuv,-DCT-50.myyuv,-DCT-90.myyuv,-big-DCT-50.myyuv}) and
 * against the unmodified reference compiled into oracle/_ref/ (random blocks through
 * Huffman::fromData/dump/fromDump, whole frames through YUV::compress/decompress).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * All float arithmetic is written as separate IEEE binary32 operations; build with -ffp-contract=off
 * and without -march/-ffast-math (oracle/Makefile) so no FMA is formed -- the reference binary has none
 * (CMakeLists.txt:3-4: no arch flags).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_OK 0
#define ORA_ERR_QUALITY 1      /* "Level of quality must be between 1 and 100"  DCT.cpp:378-382,438-442 */
#define ORA_ERR_WIDTH 2        /* "Error. width % 8 must be 0"                  DCT.cpp:280-282,338-340 */
#define ORA_ERR_HEIGHT 3       /* "Error. height % 8 must be 0"                 DCT.cpp:283-285,341-343 */
#define ORA_ERR_CAPACITY 4     /* output buffer too small (oracle API only)                             */
#define ORA_ERR_DCTYUV_SIZE 5  /* "DCTYUV load bad size"                        DCT.cpp:132-134,143-145 */
#define ORA_ERR_PLANE_SIZE 6   /* "DCTYUVPlane load ... bad size"               DCT.cpp:41-55           */
#define ORA_ERR_HUFF_CODE 7    /* "Huffman bad code" / "Huffman unknown symbol" Huffman.cpp:121,130,139 */
#define ORA_ERR_CHUNK 8        /* chunk overruns its plane content (UB in the reference)                */

/* ------------------------------------------------------------------------------------------------
 * Colour conversion  (myyuv_yuv.cpp:19-27 divide_roundnearest, :34-52 getYUV444FromRGB2x2,
 *                     :88-128 the IYUV lambda; myyuv_bmp.cpp:80-103 colorData row flip)
 * ---------------------------------------------------------------------------------------------- */
static inline void ora_pixel_yuv444(const uint8_t* px, uint8_t* y, uint8_t* cb, uint8_t* cr) {
  /* myyuv_yuv.cpp:43-49.  Bytes are B,G,R,X.  The float->uint8_t casts of possibly negative values are
   * what x86-64/GCC emits: cvttss2si to int32, keep the low byte; "+ 128" is then done in int and the
   * store to uint8_t yuv444[] wraps again. */
  const float B = (float)px[0];
  const float G = (float)px[1];
  const float R = (float)px[2];
  const float t0 = 0.299f * R;
  const float t1 = 0.587f * G;
  const float t2 = 0.114f * B;
  const float s0 = t0 + t1;
  const float Y = s0 + t2;
  const float db = B - Y;
  const float dr = R - Y;
  const float fb = db * 0.564f;
  const float fr = dr * 0.713f;
  *y = (uint8_t)(int32_t)Y;
  *cb = (uint8_t)((int)(uint8_t)(int32_t)fb + 128);
  *cr = (uint8_t)((int)(uint8_t)(int32_t)fr + 128);
}

static inline unsigned ora_div4_nearest(uint8_t v) {
  /* divide_roundnearest(v, 4_uchar), myyuv_yuv.cpp:19-27: both operands non-negative -> (v + 2) / 4 in int */
  return ((unsigned)v + 2u) / 4u;
}

/* px: rows as they lie in the BMP file, pixel_bytes bytes per pixel (4: B,G,R,X; 3: B,G,R -- getYUV444FromRGB2x2
 * addresses pixels as index * (bit_count / 8), myyuv_yuv.cpp:34-41; the 32-bit assert at :92 is compiled out of the
 * Release build the reference's README asks for).  bottom_up != 0: file row 0 is the bottom image row
 * (BMP height > 0, myyuv_bmp.cpp:95-98); bottom_up == 0: rows already top-down (height < 0, :87-88). */
void ora_bgr_to_iyuv_pb(const uint8_t* bgrx, uint32_t w, uint32_t h, int bottom_up, uint32_t pixel_bytes, uint8_t* out) {
  uint8_t* yp = out;
  uint8_t* up = out + (size_t)w * h;          /* myyuv_yuv.cpp:105-107 */
  uint8_t* vp = out + (size_t)w * h * 5 / 4;
  for (uint32_t j = 0; j < h; j += 2) {        /* :108-124 */
    for (uint32_t i = 0; i < w; i += 2) {
      uint8_t y4[4], cb4[4], cr4[4];
      for (int s = 0; s < 4; s++) {
        const uint32_t row = j + (uint32_t)(s >> 1), col = i + (uint32_t)(s & 1);
        const uint32_t frow = bottom_up ? (h - 1 - row) : row;
        ora_pixel_yuv444(bgrx + ((size_t)frow * w + col) * pixel_bytes, &y4[s], &cb4[s], &cr4[s]);
      }
      /* :114-115: the four rounded quarters are summed and stored to uint8_t (wraps at 256) */
      const uint8_t Cb = (uint8_t)(ora_div4_nearest(cb4[0]) + ora_div4_nearest(cb4[1]) + ora_div4_nearest(cb4[2]) +
                                   ora_div4_nearest(cb4[3]));
      const uint8_t Cr = (uint8_t)(ora_div4_nearest(cr4[0]) + ora_div4_nearest(cr4[1]) + ora_div4_nearest(cr4[2]) +
                                   ora_div4_nearest(cr4[3]));
      const size_t loc = (size_t)i + (size_t)j * w;
      yp[loc] = y4[0];
      yp[loc + 1] = y4[1];
      yp[loc + w] = y4[2];
      yp[loc + w + 1] = y4[3];
      const size_t k = ((size_t)i + (size_t)j * w / 2) / 2; /* :120 */
      up[k] = Cb;
      vp[k] = Cr;
    }
  }
}

void ora_bgrx_to_iyuv(const uint8_t* bgrx, uint32_t w, uint32_t h, int bottom_up, uint8_t* out) {
  ora_bgr_to_iyuv_pb(bgrx, w, h, bottom_up, 4, out);
}

/* ------------------------------------------------------------------------------------------------
 * Tables (DCT.cpp:199-230, Huffman.cpp:32-34)
 * The two q50 tables are the JPEG Annex K tables and the zigzag is the JPEG zigzag (public standard
 * data).  The 8x8 DCT matrix is NOT exactly symmetric in the reference (DCT.cpp:221-230); its 64 float
 * bit patterns are parity-critical data and are carried here as IEEE-754 hex words.
 * ---------------------------------------------------------------------------------------------- */
static const uint8_t ORA_Q50_LUMA[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                         14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                         18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                         49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t ORA_Q50_CHROMA[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                           24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t ORA_ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                       12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                       35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                       58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
#include "dct_matrix_bits.h" /* static const uint32_t ORA_DCT_BITS[64]: float bit patterns of DCT.cpp:221-230 */

static inline float ora_C(int row, int col) {
  float f;
  memcpy(&f, &ORA_DCT_BITS[row * 8 + col], 4);
  return f;
}

/* DCT.cpp:286-290 (dup :344-348): quality-scaled table.  std::round = half away from zero. */
void ora_qtable(uint8_t q, int chroma, float qt[64]) {
  const float qf = (float)q;
  const float mul = (qf >= 50.5f) ? (100.0f - qf) / 50.0f : 50.0f / qf;
  const uint8_t* base = chroma ? ORA_Q50_CHROMA : ORA_Q50_LUMA;
  for (int i = 0; i < 64; i++) {
    float v = roundf((float)base[i] * mul);
    if (v < 1.0f) v = 1.0f;
    if (v > 255.0f) v = 255.0f;
    qt[i] = v;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Forward DCT + quantisation of one 8x8 block  (DCT.cpp:301-305 load, :269-277 applyDCTBlock,
 * :232-242 squareMatrixMul, :244-254 squareMatrixMulT)
 * Each output is acc = 0.0f; for k ascending: acc = acc + (a*b)  (product rounded, then sum rounded).
 * ---------------------------------------------------------------------------------------------- */
void ora_fdct_quant_block(const uint8_t* px, uint32_t stride, const float qt[64], int16_t coef[64]) {
  float X[64], T[64], Yc[64];
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) X[r * 8 + c] = (float)px[(size_t)r * stride + c] - 128.0f;
  for (int a = 0; a < 8; a++)   /* T = C . X */
    for (int c = 0; c < 8; c++) {
      float acc = 0.0f;
      for (int k = 0; k < 8; k++) {
        const float p = ora_C(a, k) * X[k * 8 + c];
        acc = acc + p;
      }
      T[a * 8 + c] = acc;
    }
  for (int a = 0; a < 8; a++)   /* Y = T . C^T */
    for (int b = 0; b < 8; b++) {
      float acc = 0.0f;
      for (int k = 0; k < 8; k++) {
        const float p = T[a * 8 + k] * ora_C(b, k);
        acc = acc + p;
      }
      Yc[a * 8 + b] = acc;
    }
  for (int i = 0; i < 64; i++) {
    const float d = Yc[i] / qt[i];
    coef[i] = (int16_t)roundf(d); /* DCT.cpp:274 */
  }
}

/* Dequantise + inverse DCT + round/clamp of one block (DCT.cpp:330-334, :256-266 squareMatrixMulT2,
 * :232-242 squareMatrixMul, :358-362 store). */
void ora_dequant_idct_block(const int16_t coef[64], const float qt[64], uint8_t* px, uint32_t stride) {
  float Bq[64], D[64], P[64];
  for (int i = 0; i < 64; i++) Bq[i] = (float)coef[i] * qt[i];
  for (int a = 0; a < 8; a++)   /* D = C^T . B */
    for (int c = 0; c < 8; c++) {
      float acc = 0.0f;
      for (int k = 0; k < 8; k++) {
        const float p = ora_C(k, a) * Bq[k * 8 + c];
        acc = acc + p;
      }
      D[a * 8 + c] = acc;
    }
  for (int a = 0; a < 8; a++)   /* P = D . C */
    for (int b = 0; b < 8; b++) {
      float acc = 0.0f;
      for (int k = 0; k < 8; k++) {
        const float p = D[a * 8 + k] * ora_C(k, b);
        acc = acc + p;
      }
      P[a * 8 + b] = acc;
    }
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) {
      int v = (int)roundf(P[r * 8 + c]) + 128; /* DCT.cpp:360 */
      if (v < 0) v = 0;
      if (v > 255) v = 255;
      px[(size_t)r * stride + c] = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Per-block Huffman coder.  Huffman.cpp:172-241 (fromData), :71-83 (generateCodeLength), :86-103
 * (generateCanonicalTree), :279-326 (dump), :36-52 (pack11bit).
 *
 * The code LENGTHS the reference produces under frequency ties depend on two libstdc++ (GCC 13)
 * behaviours that the reference inherits (Huffman.cpp:173 std::unordered_map<int16_t,uint8_t>,
 * :204 std::priority_queue): the iteration order of the hash map and the push_heap/pop_heap
 * algorithms.  Both are restated below from their published semantics (libstdc++ hashtable.h:
 * _M_insert_bucket_begin / _M_rehash_aux / _Prime_rehash_policy::_M_need_rehash; stl_heap.h:
 * __push_heap / __adjust_heap) and pinned by the golden files and by oracle/_ref.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int16_t key[66];
  int n;
  unsigned nbuckets;      /* 1 (nothing allocated) -> 13 -> 29 -> 59 -> 127 */
  unsigned next_resize;   /* _Prime_rehash_policy::_M_next_resize */
} ora_hashlist;

static unsigned ora_bucket(int16_t key, unsigned nb) {
  /* std::hash<short> is static_cast<size_t>(v): sign-extended to 64 bits, then % bucket_count */
  return (unsigned)((uint64_t)(int64_t)key % (uint64_t)nb);
}

/* libstdc++ list rule shared by insert and rehash: a key whose bucket already holds nodes goes
 * immediately before the first node of that bucket's run, otherwise to the front of the whole list. */
static void ora_list_place(int16_t* list, int* n, int16_t key, unsigned nb) {
  const unsigned b = ora_bucket(key, nb);
  int pos = 0;
  for (int i = 0; i < *n; i++)
    if (ora_bucket(list[i], nb) == b) { pos = i; break; }
  memmove(list + pos + 1, list + pos, (size_t)(*n - pos) * sizeof(int16_t));
  list[pos] = key;
  (*n)++;
}

static unsigned ora_next_prime_bkt(unsigned want) {
  /* the only sizes reachable with <= 65 keys: _M_next_bkt(12)=13, (26)=29, (58)=59, (118)=127 */
  static const unsigned primes[] = {13, 29, 59, 127};
  for (int i = 0; i < 4; i++)
    if (primes[i] >= want) return primes[i];
  return 257;
}

static void ora_hash_insert(ora_hashlist* h, int16_t key) {
  /* _M_insert_unique_node: _M_need_rehash(bkt_count, element_count, 1) first, then _M_insert_bucket_begin */
  const unsigned n_after = (unsigned)h->n + 1u;
  if (n_after > h->next_resize) {
    unsigned min_bkts = n_after;
    if (h->next_resize == 0 && min_bkts < 11) min_bkts = 11;
    if (min_bkts >= h->nbuckets) {
      unsigned want = min_bkts + 1;
      if (h->nbuckets * 2 > want) want = h->nbuckets * 2;
      const unsigned nb = ora_next_prime_bkt(want);
      /* _M_rehash_aux(unique keys): walk old list front to back, re-place each node */
      int16_t old[66];
      const int on = h->n;
      memcpy(old, h->key, sizeof(int16_t) * (size_t)on);
      h->n = 0;
      for (int i = 0; i < on; i++) ora_list_place(h->key, &h->n, old[i], nb);
      h->nbuckets = nb;
      h->next_resize = nb; /* floor(nb * max_load_factor 1.0) */
    } else {
      h->next_resize = h->nbuckets;
    }
  }
  ora_list_place(h->key, &h->n, key, h->nbuckets);
}

static int ora_hash_find(const ora_hashlist* h, int16_t key) {
  for (int i = 0; i < h->n; i++)
    if (h->key[i] == key) return i;
  return -1;
}

typedef struct {
  uint8_t freq[131];
  uint8_t parent[131];
} ora_tree;

/* std::push_heap with Compare(a,b) = a.freq > b.freq (Huffman.hpp:41-45): sift the last element up
 * while the parent's freq is strictly greater. */
static void ora_heap_sift_up(uint8_t* heap, int hole, int top, uint8_t value, const ora_tree* t) {
  int parent = (hole - 1) / 2;
  while (hole > top && t->freq[heap[parent]] > t->freq[value]) {
    heap[hole] = heap[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  heap[hole] = value;
}

/* std::pop_heap + pop_back: returns the old top; heap shrinks by one (stl_heap.h __pop_heap/__adjust_heap). */
static uint8_t ora_heap_pop(uint8_t* heap, int* size, const ora_tree* t) {
  const uint8_t top = heap[0];
  const int len = *size - 1;
  const uint8_t value = heap[len];
  *size = len;
  if (len == 0) return top;
  int hole = 0, child = 0;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (t->freq[heap[child]] > t->freq[heap[child - 1]]) child--;
    heap[hole] = heap[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    heap[hole] = heap[child - 1];
    hole = child - 1;
  }
  ora_heap_sift_up(heap, hole, 0, value, t);
  return top;
}

static unsigned ora_group_bytes(unsigned cnt) { return 1u + (cnt * 11u + 7u) / 8u; }

/* coef: 64 quantised coefficients, row-major.  out: >= 256 bytes.  Returns the chunk size (7..~173). */
int ora_huff_encode_block(const int16_t coef[64], uint8_t* out) {
  int16_t z[64];
  ora_hashlist h;
  uint8_t cnt[66];
  memset(&h, 0, sizeof h);
  h.nbuckets = 1;
  memset(cnt, 0, sizeof cnt);
  /* Huffman.cpp:176-189: zigzag walk, freq[d]++ in scan order, trailing-zero run length.
   * counts are kept per KEY (indexable by list position only after the order settles), so keep a
   * side table keyed by first-seen order and map through find(). */
  int16_t seen_key[66];
  int seen_n = 0;
  unsigned trailing = 0;
  for (int i = 0; i < 64; i++) {
    const int16_t d = coef[ORA_ZIGZAG[i]];
    z[i] = d;
    int s = -1;
    for (int k = 0; k < seen_n; k++)
      if (seen_key[k] == d) { s = k; break; }
    if (s < 0) {
      s = seen_n++;
      seen_key[s] = d;
      ora_hash_insert(&h, d);
    }
    cnt[s]++;
    trailing = (d == 0) ? trailing + 1 : 0;
  }
  unsigned msg_len = 64 - trailing; /* :190 */
  /* :192-203: drop the trailing zeros from freq[0]; freq[0] (operator[]) inserts key 0 if absent */
  int zs = -1;
  for (int k = 0; k < seen_n; k++)
    if (seen_key[k] == 0) { zs = k; break; }
  if (zs >= 0) cnt[zs] = (uint8_t)(cnt[zs] - trailing);
  if (zs < 0) {
    zs = seen_n++;
    seen_key[zs] = 0;
    cnt[zs] = 0;
    ora_hash_insert(&h, 0);
  }
  if (cnt[zs] == 0) {
    if (msg_len == 0) {
      cnt[zs] = 1;
      msg_len = 1;
    } else {
      const int p = ora_hash_find(&h, 0); /* freq.erase(0): order of the rest unchanged */
      memmove(h.key + p, h.key + p + 1, sizeof(int16_t) * (size_t)(h.n - p - 1));
      h.n--;
    }
  }
  /* leaves in map iteration order (:207-209) */
  const int nleaf = h.n;
  ora_tree t;
  int16_t leaf_sym[66];
  uint8_t heap[66];
  int hsize = 0;
  for (int i = 0; i < nleaf; i++) {
    leaf_sym[i] = h.key[i];
    int s = 0;
    while (seen_key[s] != h.key[i]) s++;
    t.freq[i] = cnt[s];
    t.parent[i] = 0xff;
    heap[hsize++] = (uint8_t)i;
    ora_heap_sift_up(heap, hsize - 1, 0, (uint8_t)i, &t);
  }
  /* :210-217: merge the two tops until one node is left */
  int nnode = nleaf;
  while (hsize > 1) {
    const uint8_t l = ora_heap_pop(heap, &hsize, &t);
    const uint8_t r = ora_heap_pop(heap, &hsize, &t);
    t.freq[nnode] = (uint8_t)(t.freq[l] + t.freq[r]);
    t.parent[nnode] = 0xff;
    t.parent[l] = (uint8_t)nnode;
    t.parent[r] = (uint8_t)nnode;
    heap[hsize++] = (uint8_t)nnode;
    ora_heap_sift_up(heap, hsize - 1, 0, (uint8_t)nnode, &t);
    nnode++;
  }
  /* :71-83: code length = leaf depth, a lone root leaf gets length 1 */
  uint8_t depth[131];
  uint8_t len_of[66];
  for (int i = nnode - 1; i >= 0; i--) depth[i] = (t.parent[i] == 0xff) ? 0 : (uint8_t)(depth[t.parent[i]] + 1);
  for (int i = 0; i < nleaf; i++) len_of[i] = depth[i] ? depth[i] : 1;
  /* tree_data: std::map<len, ascending symbols> (:76-78) and canonical codes (:86-103) */
  int order[66];
  for (int i = 0; i < nleaf; i++) order[i] = i;
  for (int i = 1; i < nleaf; i++) { /* sort by (len, symbol) */
    const int v = order[i];
    int j = i - 1;
    while (j >= 0 && (len_of[order[j]] > len_of[v] || (len_of[order[j]] == len_of[v] && leaf_sym[order[j]] > leaf_sym[v]))) {
      order[j + 1] = order[j];
      j--;
    }
    order[j + 1] = v;
  }
  uint8_t code_of[66];
  {
    uint8_t code = 0, prev = 0;
    for (int i = 0; i < nleaf; i++) {
      const int s = order[i];
      code = (uint8_t)(code << (len_of[s] - prev));
      code_of[s] = code;
      code++;
      prev = len_of[s];
    }
  }
  /* :279-316 dump header + code table */
  unsigned pos = 3;
  for (int i = 0; i < nleaf;) {
    int j = i;
    while (j < nleaf && len_of[order[j]] == len_of[order[i]]) j++;
    int left = j - i, at = i;
    while (left > 0) { /* groups of <= 32 with the same length (:307-315) */
      const int c = left > 32 ? 32 : left;
      out[pos++] = (uint8_t)(((len_of[order[i]] - 1) << 5) | (c - 1));
      const unsigned nb = ((unsigned)c * 11u + 7u) / 8u;
      memset(out + pos, 0, nb);
      for (int m = 0; m < c; m++) { /* pack11bit :36-52 */
        const int16_t sv = leaf_sym[order[at + m]];
        const unsigned num = (sv < 0) ? (unsigned)(2048 + sv) : (unsigned)sv;
        const unsigned bit = (unsigned)m * 11u;
        for (unsigned bb = 0; bb < 11; bb++)
          if (num & (1u << bb)) out[pos + ((bit + bb) >> 3)] |= (uint8_t)(1u << ((bit + bb) & 7));
      }
      pos += nb;
      at += c;
      left -= c;
    }
    i = j;
  }
  const unsigned tree_bytes = pos - 3;
  /* :227-236 code stream: each code MSB first; stream bit p lives in bit p%8 of byte p/8 (:319-325) */
  unsigned bits = 0;
  uint8_t data[64];
  memset(data, 0, sizeof data);
  for (unsigned i = 0; i < msg_len; i++) {
    int s = 0;
    while (leaf_sym[s] != z[i]) s++;
    for (int j = 0; j < len_of[s]; j++) {
      if ((code_of[s] >> (len_of[s] - 1 - j)) & 1) data[bits >> 3] |= (uint8_t)(1u << (bits & 7));
      bits++;
    }
  }
  const unsigned data_bytes = (bits + 7) / 8;
  out[0] = (uint8_t)(bits & 0xff);
  out[1] = (uint8_t)(bits >> 8);
  out[2] = (uint8_t)tree_bytes;
  memcpy(out + pos, data, data_bytes);
  (void)ora_group_bytes;
  return (int)(pos + data_bytes);
}

/* Huffman.cpp:243-277 fromDump, :54-69 unpack11bit, :143-154 decodeFromTreeData, :106-141 decodeSymbol.
 * Reads outside [chunk, chunk+size) (undefined behaviour in the reference's NDEBUG build) are
 * reported as ORA_ERR_CHUNK instead. */
int ora_huff_decode_block(const uint8_t* chunk, uint32_t size, int16_t coef[64]) {
  memset(coef, 0, 64 * sizeof(int16_t));
  if (size < 3) return ORA_ERR_CHUNK;
  const unsigned bits = (unsigned)chunk[0] | ((unsigned)chunk[1] << 8);
  const unsigned tree_bytes = chunk[2];
  const unsigned data_bytes = (bits + 7) / 8;
  if (bits > 512 || 3 + tree_bytes + data_bytes > size) return ORA_ERR_CHUNK;
  int16_t sym[9][72];
  unsigned count[9];
  memset(count, 0, sizeof count);
  unsigned i = 3;
  while (i - 3 < tree_bytes) {
    const uint8_t info = chunk[i++];
    const unsigned len = (unsigned)(info >> 5) + 1, c = (unsigned)(info & 31) + 1;
    const unsigned nb = (c * 11u + 7u) / 8u;
    if (i - 3 + nb > tree_bytes) return ORA_ERR_CHUNK;
    for (unsigned m = 0; m < c; m++) {
      const unsigned bit = m * 11u;
      unsigned v = 0;
      for (unsigned bb = 0; bb < 11; bb++)
        if (chunk[i + ((bit + bb) >> 3)] & (1u << ((bit + bb) & 7))) v |= 1u << bb;
      if (count[len] >= 72) return ORA_ERR_CHUNK;
      sym[len][count[len]++] = (int16_t)((v >= 1024) ? (int)v - 2048 : (int)v);
    }
    i += nb;
  }
  const uint8_t* data = chunk + 3 + tree_bytes;
  unsigned p = 0, j = 0;
  while (p < bits && j < 64) {
    uint8_t code = 0, first = 0; /* uint8_t on purpose: Huffman.cpp:107-108 */
    int found = 0;
    for (unsigned len = 1; len <= 8; len++) {
      const unsigned c = count[len];
      if (p >= bits) return ORA_ERR_HUFF_CODE;            /* :120-122 */
      code |= (uint8_t)((data[p >> 3] >> (p & 7)) & 1);
      p++;
      if ((unsigned)code < c + (unsigned)first) {         /* :126 */
        coef[ORA_ZIGZAG[j++]] = sym[len][code - first];
        found = 1;
        break;
      }
      first = (uint8_t)(first + c);
      first = (uint8_t)(first << 1);
      code = (uint8_t)(code << 1);
    }
    if (!found) return ORA_ERR_HUFF_CODE;                  /* :139 */
  }
  return ORA_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Whole-frame compress / decompress and the payload layout
 * (DCT.cpp:371-430 compress_DCT_planar, :279-323 applyDCTPlane, :112-173 DCTYUV, :16-73 DCTYUVPlane,
 *  :432-488 decompress_DCT_planar, :337-365 restoreDCTPlane)
 *   payload := u32 planes_sizes[3]  plane[0] plane[1] plane[2]
 *   plane   := u32 n_chunks  u32 content_size  u8 chunk_size[n]  u8 content[content_size]
 * ---------------------------------------------------------------------------------------------- */
static void ora_put32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static uint32_t ora_get32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

uint32_t ora_compress_bound(uint32_t w, uint32_t h) {
  const uint64_t nblk = (uint64_t)w * h / 64 * 3 / 2;
  return (uint32_t)(12 + 24 + nblk * 256);
}

int ora_compress(const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t q[3], uint8_t* out, uint32_t cap,
                 uint32_t* out_size) {
  for (int i = 0; i < 3; i++)
    if (q[i] < 1 || q[i] > 100) return ORA_ERR_QUALITY;
  const uint32_t pw[3] = {w, w / 2, w / 2}, ph[3] = {h, h / 2, h / 2};
  const uint8_t* src[3] = {iyuv, iyuv + (size_t)w * h, iyuv + (size_t)w * h * 5 / 4};
  size_t pos = 12;
  for (int p = 0; p < 3; p++) {
    if (pw[p] % 8) return ORA_ERR_WIDTH;
    if (ph[p] % 8) return ORA_ERR_HEIGHT;
    float qt[64];
    ora_qtable(q[p], p != 0, qt);
    const uint32_t bw = pw[p] / 8, n = bw * (ph[p] / 8);
    uint8_t* chunks = (uint8_t*)malloc((size_t)n * 256);
    uint8_t* sizes = (uint8_t*)malloc(n);
    if (!chunks || !sizes) { free(chunks); free(sizes); return ORA_ERR_CAPACITY; }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)n; k++) {
      int16_t coef[64];
      const uint32_t bx = (uint32_t)k % bw, by = (uint32_t)k / bw;
      ora_fdct_quant_block(src[p] + (size_t)by * 8 * pw[p] + (size_t)bx * 8, pw[p], qt, coef);
      sizes[k] = (uint8_t)ora_huff_encode_block(coef, chunks + (size_t)k * 256);
    }
    size_t content = 0;
    for (uint32_t k = 0; k < n; k++) content += sizes[k];
    if (pos + 8 + n + content > cap) { free(chunks); free(sizes); return ORA_ERR_CAPACITY; }
    ora_put32(out + 4 * p, (uint32_t)(8 + n + content));
    ora_put32(out + pos, n);
    ora_put32(out + pos + 4, (uint32_t)content);
    memcpy(out + pos + 8, sizes, n);
    size_t c = pos + 8 + n;
    for (uint32_t k = 0; k < n; k++) {
      memcpy(out + c, chunks + (size_t)k * 256, sizes[k]);
      c += sizes[k];
    }
    pos = c;
    free(chunks);
    free(sizes);
  }
  *out_size = (uint32_t)pos;
  return ORA_OK;
}

/* quantised coefficients of one plane, [n_blocks][64] row-major (test helper, same maths as above) */
int ora_plane_coefs(const uint8_t* plane, uint32_t w, uint32_t h, uint8_t q, int chroma, int16_t* coefs) {
  if (q < 1 || q > 100) return ORA_ERR_QUALITY;
  if (w % 8) return ORA_ERR_WIDTH;
  if (h % 8) return ORA_ERR_HEIGHT;
  float qt[64];
  ora_qtable(q, chroma, qt);
  const uint32_t bw = w / 8, n = bw * (h / 8);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)n; k++)
    ora_fdct_quant_block(plane + (size_t)((uint32_t)k / bw) * 8 * w + (size_t)((uint32_t)k % bw) * 8, w, qt, coefs + 64 * k);
  return ORA_OK;
}

int ora_decompress(const uint8_t* payload, uint32_t size, uint32_t w, uint32_t h, const uint8_t q[3], uint8_t* iyuv) {
  for (int i = 0; i < 3; i++)
    if (q[i] < 1 || q[i] > 100) return ORA_ERR_QUALITY;
  if (size <= 12) return ORA_ERR_DCTYUV_SIZE; /* DCT.cpp:132-134 */
  uint32_t psz[3];
  uint64_t tot = 12;
  for (int p = 0; p < 3; p++) { psz[p] = ora_get32(payload + 4 * p); tot += psz[p]; }
  if (size < tot) return ORA_ERR_DCTYUV_SIZE;  /* :138-146 */
  const uint32_t pw[3] = {w, w / 2, w / 2}, ph[3] = {h, h / 2, h / 2};
  uint8_t* dst[3] = {iyuv, iyuv + (size_t)w * h, iyuv + (size_t)w * h * 5 / 4};
  size_t ppos = 12;
  int rc = ORA_OK;
  for (int p = 0; p < 3 && rc == ORA_OK; p++) {
    if (pw[p] % 8) return ORA_ERR_WIDTH;
    if (ph[p] % 8) return ORA_ERR_HEIGHT;
    const uint8_t* pl = payload + ppos;
    if (psz[p] <= 8) return ORA_ERR_PLANE_SIZE; /* :41-43 */
    const uint32_t n = ora_get32(pl), content = ora_get32(pl + 4);
    if (n == 0 || content == 0) return ORA_ERR_PLANE_SIZE;       /* :47-52 */
    if ((uint64_t)psz[p] < 8ull + n + content) return ORA_ERR_PLANE_SIZE; /* :53-55 */
    const uint32_t bw = pw[p] / 8, need = bw * (ph[p] / 8);
    if (n < need) return ORA_ERR_PLANE_SIZE;
    const uint8_t* sizes = pl + 8;
    const uint8_t* cont = sizes + n;
    uint32_t* off = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)need);
    if (!off) return ORA_ERR_CAPACITY;
    uint64_t run = 0;
    for (uint32_t k = 0; k < need; k++) { off[k] = (uint32_t)run; run += sizes[k]; } /* getContentPos :21-33 */
    if (run > content) { free(off); return ORA_ERR_CHUNK; }
    float qt[64];
    ora_qtable(q[p], p != 0, qt);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)need; k++) {
      int16_t coef[64];
      const int e = ora_huff_decode_block(cont + off[k], sizes[k], coef);
      if (e != ORA_OK) {
#pragma omp critical
        rc = e;
        continue;
      }
      ora_dequant_idct_block(coef, qt, dst[p] + (size_t)((uint32_t)k / bw) * 8 * pw[p] + (size_t)((uint32_t)k % bw) * 8, pw[p]);
    }
    free(off);
    ppos += psz[p];
  }
  return rc;
}

/* decode every chunk of a payload to coefficients (test helper: [all blocks in file order][64]) */
int ora_payload_coefs(const uint8_t* payload, uint32_t size, uint32_t w, uint32_t h, int16_t* coefs) {
  if (size <= 12) return ORA_ERR_DCTYUV_SIZE;
  const uint32_t pw[3] = {w, w / 2, w / 2}, ph[3] = {h, h / 2, h / 2};
  size_t ppos = 12, blk = 0;
  for (int p = 0; p < 3; p++) {
    const uint8_t* pl = payload + ppos;
    const uint32_t n = ora_get32(pl);
    const uint32_t need = (pw[p] / 8) * (ph[p] / 8);
    if (n < need) return ORA_ERR_PLANE_SIZE;
    const uint8_t* sizes = pl + 8;
    const uint8_t* c = sizes + n;
    for (uint32_t k = 0; k < need; k++) {
      const int e = ora_huff_decode_block(c, sizes[k], coefs + 64 * blk);
      if (e != ORA_OK) return e;
      c += sizes[k];
      blk++;
    }
    ppos += ora_get32(payload + 4 * p);
  }
  return ORA_OK;
}

/* batch helpers used by the tests to push many blocks through the block-level functions */
void ora_huff_encode_blocks(const int16_t* coef, uint32_t n, uint8_t* out, uint8_t* sizes) {
  for (uint32_t b = 0; b < n; b++) {
    uint8_t tmp[256];
    const int s = ora_huff_encode_block(coef + 64 * (size_t)b, tmp);
    memcpy(out, tmp, (size_t)s);
    out += s;
    sizes[b] = (uint8_t)s;
  }
}

int ora_huff_decode_blocks(const uint8_t* chunks, const uint8_t* sizes, uint32_t n, int16_t* coef) {
  for (uint32_t b = 0; b < n; b++) {
    const int e = ora_huff_decode_block(chunks, sizes[b], coef + 64 * (size_t)b);
    if (e != ORA_OK) return e;
    chunks += sizes[b];
  }
  return ORA_OK;
}
