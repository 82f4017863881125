// block_codec.cuh -- per-8x8-block entropy coding of the reference's DCT payload, written for one GPU
// thread per block.  How the 32 lanes of a warp go through the data dependent loops is a POLICY argument of every function:
//   WarpFree      (the kernels' choice) each lane runs its own trip counts and the hardware reconverges the warp behind the
//                 loop; only choices the whole warp makes together use a warp reduction (umax);
//   WarpLockstep  (round 1; -DMYB_LOCKSTEP) every loop runs to the warp maximum with a predicated body and a warp barrier per
//                 iteration -- written after a first version whose lanes had drifted apart (2.6 active lanes per instruction,
//                 profiles/r01_notes.md); with one call per phase and sorted blocks it only costs instructions;
//   NoWarp        every "lane" runs alone: the host build.
// Everything here is integer/byte work, so it is plain __host__ __device__ code: tests/hostemu/hostemu.cpp compiles
// the same header with g++ (policy NoWarp) so the logic can be checked against the oracle without a GPU
// (test infrastructure only -- the product never runs it on the CPU).
//
// Format and semantics follow the reference (paths relative to /root/reference):
//   Huffman.cpp:172-241  fromData   zigzag, trailing-zero trim, histogram, tree, canonical codes, code stream
//   Huffman.cpp:279-326  dump       u16 bits | u8 table_bytes | groups {(len-1)<<5|(cnt-1), 11-bit symbols} | stream
//   Huffman.cpp:243-277  fromDump,  :106-154 decodeSymbol / decodeFromTreeData
// Byte-exactness with the reference needs its tie-breaking, which comes from libstdc++ (GCC 13):
// std::unordered_map<int16_t,uint8_t> iteration order and std::priority_queue's heap algorithms
// (Huffman.cpp:173,204-217).  list_place()/heap_*() below re-implement those published semantics with
// fixed-size arrays; see SURVEY.md App. A.5 and DESIGN.md "Tie-breaking".
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MYB_HD __host__ __device__ __forceinline__
#define MYB_NOUNROLL _Pragma("unroll 1")  // keep the (large, divergent) coder loops rolled: code size is I-cache bound
#define MYB_UNROLL4 _Pragma("unroll 4")
#else
#define MYB_HD inline
#define MYB_NOUNROLL
#define MYB_UNROLL4
#endif

namespace myyuvb {

// zigzag scan position -> row-major coefficient index (JPEG zigzag; Huffman.cpp:32-34)
#define MYB_ZIGZAG_LIST                                                                                       \
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,   \
      14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, \
      46, 53, 60, 61, 54, 47, 55, 62, 63

// ---- warp cooperation policies -------------------------------------------------------------------
struct NoWarp {  // host emulation and the out-of-line fallback path: every "lane" runs alone
  MYB_HD int max(int v) const { return v; }
  MYB_HD int umax(int v) const { return v; }
  MYB_HD bool any(bool p) const { return p; }
  MYB_HD void sync() const {}
};
#if defined(__CUDACC__)
struct WarpFree {  // every lane runs its own trip counts; the hardware reconverges the warp behind each loop
  __device__ __forceinline__ int max(int v) const { return v; }
  __device__ __forceinline__ int umax(int v) const { return __reduce_max_sync(0xffffffffu, v); }  // for choices the whole warp makes together
  __device__ __forceinline__ bool any(bool p) const { return p; }
  __device__ __forceinline__ void sync() const {}
};
struct WarpLockstep {  // all 32 lanes of the warp call the codec functions together
  __device__ __forceinline__ int max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
  __device__ __forceinline__ int umax(int v) const { return __reduce_max_sync(0xffffffffu, v); }
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
#endif

MYB_HD uint32_t bit_reverse32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
#endif
}

MYB_HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

// ---------------------------------------------------------------------------------------------------
// Scratch memory of one block's Huffman build.  All arrays are bytes addressed as b[(off + i) * stride]
// (symbols: 16-bit, h[i * stride]; the stride is a template constant so index arithmetic folds into the
// addressing).  On the GPU the fast instance lives in shared memory with stride = threads per CTA, so the lanes of a warp hit consecutive bytes; the fallback instance for
// blocks with many distinct symbols lives in per-thread local memory with stride 1.
// CAP = max distinct symbols (+1 slot for the key 0 that freq[0] may insert, Huffman.cpp:195).
// ---------------------------------------------------------------------------------------------------
template <int CAP, int STRIDE>
struct HuffScratch {
  uint8_t* b;
  int16_t* h;
  static constexpr int kCap = CAP;
  // Arrays that are never live at the same time share their bytes (the scratch of 64-symbol blocks sits in shared memory
  // in heavy_blocks_kernel, where every byte per lane costs occupancy):
  //   list order -> (dead after the code lengths are assigned) -> canonical codes
  //   bucket list, then heap -> (dead after the merges) -> sorted slots
  //   map emulation's table and links -> heap weights (upper half, dead after the merges); lower half: code lengths
  //   value table of the histogram -> (dead after the histogram) -> parent links / depths
  static constexpr int kCnt = 0;                 // [CAP+1] occurrences of slot s in the message
  static constexpr int kOrd = kCnt + CAP + 1;    // [CAP+1] hash-list order: slot at list position p
  static constexpr int kCode = kOrd;             // [CAP]   bit-reversed canonical code of slot s
  static constexpr int kBkt = kOrd + CAP + 1;    // [CAP+1] bucket of list position p; later: heap
  static constexpr int kSorted = kBkt;           // [CAP]   slots ordered by (length, symbol value)
  static constexpr int kFreq = kBkt + CAP + 1;   // [2*CAP] hash table + links of the map emulation, then upper half: heap weights
  static constexpr int kLen = kFreq;             // [CAP]   code length of slot s
  static constexpr int kPar = kFreq + 2 * CAP;   // [2*CAP] parent node, then depth
  // [128] slot of the value v at index v + 64, 0xff = not seen: over the parent links when those are 128 bytes, over
  // node weights + parent links when the two together are (CAP = 32), else behind them
  static constexpr int kLut = 2 * CAP >= 128 ? kPar : 4 * CAP >= 128 ? kFreq : kPar + 2 * CAP;
  static constexpr int kBytes = 4 * CAP >= 128 ? kPar + 2 * CAP : kLut + 128;  // bytes per block
  static constexpr int kSyms = CAP + 1;          // int16 per block
  MYB_HD uint8_t& at(int off, int i) const { return b[(off + i) * STRIDE]; }
  MYB_HD int16_t& sym(int i) const { return h[i * STRIDE]; }
};

// Accessor pair for coefficients that stay where they are (read only, e.g. the queue of deferred blocks in global
// memory) while the slot numbers go to a byte column of the scratch: huff_plan takes ZSplitValues, huff_emit ZSplitSlots.
template <int STRIDE>
struct ZSplitValues {
  const uint16_t* v;  // 64 coefficient words, zigzag order (low 11 bits: the value)
  uint8_t* col;       // slot of message position i at col[i * STRIDE]
  static constexpr bool kKeepsValues = true;
  MYB_HD int get(int i) const { return ((int)((uint32_t)v[i] << 21)) >> 21; }
  MYB_HD void set(int i, int s) { col[i * STRIDE] = (uint8_t)s; }
};
template <int STRIDE>
struct ZSplitSlots {
  const uint8_t* col;
  MYB_HD int get(int i) const { return col[i * STRIDE]; }
};
template <class Z, class = void>
struct z_keeps_values { static constexpr bool value = false; };
template <class Z>
struct z_keeps_values<Z, decltype((void)Z::kKeepsValues)> { static constexpr bool value = Z::kKeepsValues; };

struct HuffPlan {
  int n;            // distinct symbols (leaves); < 0: scratch capacity exceeded, retry with the large instance
  int msg_len;      // coded symbols (1..64)
  int bits;         // code stream bits
  int table_bytes;  // bytes of the serialised code table
  uint64_t per_len; // symbols per code length 1..8, one byte each
  MYB_HD int size() const { return 3 + table_bytes + ((bits + 7) >> 3); }
};

// std::hash<short>(v) % nb with v sign-extended to 64 bits (libstdc++ functional_hash.h); nb in {13,29,59,127}.
// 2^64 mod nb = 3, 24, 5, 2 respectively.
MYB_HD int hash_bucket13(int v) {
  const int m = (v < 0 ? -v : v) % 13;
  if (v >= 0) return m;
  const int r = 3 - m;
  return r < 0 ? r + 13 : r;
}
MYB_HD int hash_bucket(int v, int nb) {
  int m, two64;
  switch (nb) {
    case 13: return hash_bucket13(v);
    case 29: two64 = 24; m = (v < 0 ? -v : v) % 29; break;
    case 59: two64 = 5; m = (v < 0 ? -v : v) % 59; break;
    default: two64 = 2; m = (v < 0 ? -v : v) % 127; break;
  }
  if (v >= 0) return m;
  const int r = two64 - m;
  return r < 0 ? r + nb : r;
}

// General case of the map's iteration order (more than 13 keys): 13 -> 29 -> 59 -> 127 buckets, rehash before
// inserting key number 14, 30, 60 (_Prime_rehash_policy::_M_need_rehash, max_load_factor 1, growth 2).
// Keys are slots 0..m-1 in first-occurrence order; erase_slot (>= 0) is removed at the end.  Result in kOrd.
// The container is emulated as it is built (hashtable.h): a singly linked list of nodes and a bucket table whose entry is
// the node BEFORE the bucket's first node.  _M_insert_bucket_begin: a key whose bucket already holds nodes goes right
// before the first of them, otherwise to the front of the whole list (and the bucket of the former first node now hangs
// off the new one).  _M_rehash_aux walks the old list front to back and re-inserts every node the same way.  Every
// insertion is a handful of accesses, where shifting an array-shaped list cost O(list length).
// Table and links live in the weight / parent area, which is not in use yet.
template <int CAP, int STRIDE>
MYB_HD void hash_list_order_general(const HuffScratch<CAP, STRIDE>& S, int m, int erase_slot) {
  using SC = HuffScratch<CAP, STRIDE>;
  constexpr int kTbl = SC::kFreq;                        // [nb] 0xff: empty bucket, 0xfe: before-begin, else a slot
  constexpr int kNxt = SC::kFreq + (CAP <= 32 ? 64 : 128);  // [CAP+1] next slot in the list, 0xff: end
  static_assert(kNxt + CAP + 1 <= SC::kBytes, "table and links must fit the weight / parent area");
  static_assert(CAP <= 32 ? CAP + 1 <= 59 : true, "a 64-byte table holds at most 59 buckets");
  int head = 0xff, nb = 13;
  auto insert = [&](int s) {
    const int b = hash_bucket(S.sym(s), nb);
    S.at(SC::kBkt, s) = (uint8_t)b;  // bucket of slot s
    const int prev = S.at(kTbl, b);
    if (prev == 0xff) {
      S.at(kNxt, s) = (uint8_t)head;
      if (head != 0xff) S.at(kTbl, S.at(SC::kBkt, head)) = (uint8_t)s;
      head = s;
      S.at(kTbl, b) = 0xfe;
    } else if (prev == 0xfe) {
      S.at(kNxt, s) = (uint8_t)head;
      head = s;
    } else {
      S.at(kNxt, s) = S.at(kNxt, prev);
      S.at(kNxt, prev) = (uint8_t)s;
    }
  };
  MYB_NOUNROLL
  for (int i = 0; i < 13; i++) S.at(kTbl, i) = 0xff;
  MYB_NOUNROLL
  for (int s = 0; s < m; s++) {
    if (s == 13 || s == 29 || s == 59) {
      nb = (s == 13) ? 29 : (s == 29) ? 59 : 127;
      MYB_NOUNROLL
      for (int i = 0; i < nb; i++) S.at(kTbl, i) = 0xff;
      int p = head;
      head = 0xff;
      MYB_NOUNROLL
      while (p != 0xff) {
        const int nx = S.at(kNxt, p);
        insert(p);
        p = nx;
      }
    }
    insert(s);
  }
  int w = 0;
  MYB_NOUNROLL
  for (int p = head; p != 0xff; p = S.at(kNxt, p))
    if (p != erase_slot) S.at(SC::kOrd, w++) = (uint8_t)p;
}

// std::push_heap with Compare(a,b) = a.freq > b.freq (Huffman.hpp:41-45; stl_heap.h __push_heap).
// The heap is two parallel byte arrays: node ids in kBkt (free after the list order is final) and their weights in the
// upper half of kFreq, so a comparison reads the weight at the heap position itself instead of going through the node id
// (this code is bound by the latency of such chains).  Weights are at most 64 (a block has 64 coefficients).
template <int CAP, int STRIDE>
MYB_HD void heap_sift_up(const HuffScratch<CAP, STRIDE>& S, int hole, int node, int wnode) {
  constexpr int kHeapW = HuffScratch<CAP, STRIDE>::kFreq + CAP;
  MYB_NOUNROLL
  while (hole > 0) {
    const int parent = (hole - 1) >> 1;
    const int pw = S.at(kHeapW, parent);
    if (!(pw > wnode)) break;
    S.at(S.kBkt, hole) = S.at(S.kBkt, parent);
    S.at(kHeapW, hole) = (uint8_t)pw;
    hole = parent;
  }
  S.at(S.kBkt, hole) = (uint8_t)node;
  S.at(kHeapW, hole) = (uint8_t)wnode;
}

// std::pop_heap + pop_back (stl_heap.h __pop_heap / __adjust_heap); returns the removed top node, its weight in wtop.
template <int CAP, int STRIDE>
MYB_HD int heap_pop(const HuffScratch<CAP, STRIDE>& S, int& hsize, int& wtop) {
  constexpr int kHeapW = HuffScratch<CAP, STRIDE>::kFreq + CAP;
  const int top = S.at(S.kBkt, 0);
  wtop = S.at(kHeapW, 0);
  const int len = hsize - 1;
  hsize = len;
  if (len == 0) return top;
  const int value = S.at(S.kBkt, len), wvalue = S.at(kHeapW, len);
  int hole = 0, child = 0;
  MYB_NOUNROLL
  while (child < ((len - 1) >> 1)) {
    child = 2 * (child + 1);
    int wc = S.at(kHeapW, child);
    const int wl = S.at(kHeapW, child - 1);
    if (wc > wl) { child--; wc = wl; }
    S.at(S.kBkt, hole) = S.at(S.kBkt, child);
    S.at(kHeapW, hole) = (uint8_t)wc;
    hole = child;
  }
  if ((len & 1) == 0 && child == ((len - 2) >> 1)) {
    child = 2 * (child + 1);
    S.at(S.kBkt, hole) = S.at(S.kBkt, child - 1);
    S.at(kHeapW, hole) = S.at(kHeapW, child - 1);
    hole = child - 1;
  }
  heap_sift_up(S, hole, value, wvalue);
  return top;
}

MYB_HD int group_table_bytes(int cnt) {  // one code length with cnt symbols, split in groups of <= 32 (Huffman.cpp:284-293)
  int bytes = 0;
  if (cnt > 32) {
    bytes = 45;
    cnt -= 32;
  }
  if (cnt > 0) bytes += 1 + ((cnt * 11 + 7) >> 3);
  return bytes;
}

template <int CAP, int STRIDE, class W>
MYB_HD HuffPlan huff_plan_tail(int L, int n, int zero_slot, bool bail, const HuffScratch<CAP, STRIDE>& S, const W& warp);

// Build the code of one block.  Z: accessor with  int get(int i)  (zigzag coefficient i) and
// void set(int i, int v); on success the first msg_len entries are overwritten with slot numbers
// (the coefficient of slot s is S.sym(s)).  L = index of the last non-zero zigzag coefficient + 1 (0: all zero).
// `warp`: cooperation policy; with WarpLockstep all 32 lanes must call this together (idle lanes pass L = 0).
template <int CAP, int STRIDE, class Z, class W>
MYB_HD HuffPlan huff_plan(Z& z, int L, const HuffScratch<CAP, STRIDE>& S, const W& warp) {
  // ---- histogram in first-occurrence order (Huffman.cpp:176-189, message part only; the trailing zeros
  // only matter through the key 0 they may add to the map, handled below).  Coefficients in [-64, 63] find
  // their slot through a 128-byte table in the scratch; the rest by a linear search.
  int n = 0, zero_slot = -1;
  bool bail = false;
  const int Lw = warp.max(L);
  if (Lw > 0) {
    for (int i = 0; i < 128; i++) S.at(S.kLut, i) = 0xff;
  }
  MYB_NOUNROLL
  for (int i = 0; i < Lw; i++) {
    if (i < L && !bail) {
      const int v = z.get(i);
      const unsigned vi = (unsigned)(v + 64);
      int s = -1;
      if (vi < 128u) {
        const int t = S.at(S.kLut, (int)vi);
        if (t != 0xff) s = t;
      } else {
        int k = 0;
        while (k < n && S.sym(k) != v) k++;
        if (k < n) s = k;
      }
      if (s < 0) {
        if (n == CAP) {
          bail = true;  // does not fit this scratch instance: undo and let the caller retry with the big one
          if constexpr (!z_keeps_values<Z>::value)
            for (int j = 0; j < i; j++) z.set(j, S.sym(z.get(j)));
        } else {
          s = n++;
          S.sym(s) = (int16_t)v;
          S.at(S.kCnt, s) = 0;
          if (v == 0) zero_slot = s;
          if (vi < 128u) S.at(S.kLut, (int)vi) = (uint8_t)s;
        }
      }
      if (!bail) {
        S.at(S.kCnt, s)++;
        z.set(i, s);
      }
    }
    warp.sync();
  }
  if (L == 0) {  // all-zero block: the single symbol 0, one bit (Huffman.cpp:195-199)
    S.sym(0) = 0;
    S.at(S.kCnt, 0) = 1;
    z.set(0, 0);
    n = 1;
    zero_slot = 0;
  }
  return huff_plan_tail(L, n, zero_slot, bail, S, warp);
}

// Everything after the histogram: S.sym(s) / S.at(kCnt, s) hold the n distinct symbols of the message in
// first-occurrence order and their counts, zero_slot the slot of the value 0 (-1: not in the message).
template <int CAP, int STRIDE, class W>
MYB_HD HuffPlan huff_plan_tail(int L, int n, int zero_slot, bool bail, const HuffScratch<CAP, STRIDE>& S, const W& warp) {
  HuffPlan pl;
  pl.n = bail ? -1 : n;
  pl.msg_len = L == 0 ? 1 : L;
  const bool tree = !bail && n > 2;  // one or two symbols: every code has length 1 (Huffman.cpp:76, :218-221)
  if (!bail && n <= 2) {
    int lo = 0;
    if (n == 2 && S.sym(1) < S.sym(0)) lo = 1;
    S.at(S.kSorted, 0) = (uint8_t)lo;
    S.at(S.kLen, lo) = 1;
    S.at(S.kCode, lo) = 0;
    if (n == 2) {
      S.at(S.kSorted, 1) = (uint8_t)(1 - lo);
      S.at(S.kLen, 1 - lo) = 1;
      S.at(S.kCode, 1 - lo) = 1;
    }
  }
  // ---- map iteration order.  Key 0 is always in the reference's map while it is filled (trailing zeros or
  // freq[0], Huffman.cpp:192-195); when the message itself has no zero it is erased again (:201) and can
  // only have mattered by triggering a rehash as key number 14, 30 or 60.
  int m = tree ? n : 0, erase_slot = -1;
  if (tree && zero_slot < 0 && (n == 13 || n == 29 || n == 59)) {
    S.sym(n) = 0;
    erase_slot = n;
    m = n + 1;
  }
  {
    // up to 13 keys: one table size (13 buckets); list and buckets are 4-bit fields of two 64-bit registers
    const bool fast = m <= 13;
    const int mw = warp.max(fast ? m : 0);
    uint64_t ord = 0, bkt = ~0ull;  // empty fields hold 0xF, which is no bucket
    MYB_NOUNROLL
    for (int s = 0; s < mw; s++) {
      if (s < m && fast) {
        const uint64_t b = (uint64_t)hash_bucket13(S.sym(s));
        const uint64_t x = bkt ^ (b * 0x1111111111111111ull);
        const uint64_t zero_nib = (x - 0x1111111111111111ull) & ~x & 0x8888888888888888ull;  // lowest hit is exact
        const int p4 = zero_nib ? (ctz64(zero_nib) & ~3) : 0;
        const uint64_t low = (1ull << p4) - 1ull;
        ord = (ord & low) | ((ord & ~low) << 4) | ((uint64_t)s << p4);
        bkt = (bkt & low) | ((bkt & ~low) << 4) | (b << p4);
      }
    }
    if (fast) {
      MYB_NOUNROLL
      for (int i = 0; i < mw; i++)
        if (i < m) S.at(S.kOrd, i) = (uint8_t)((ord >> (4 * i)) & 15u);
    } else {
      hash_list_order_general(S, m, erase_slot);
    }
    warp.sync();
  }
  // ---- leaves pushed in list order (Huffman.cpp:207-209); node id = list position, weights travel with the heap entries
  const int nt = tree ? n : 0;
  const int nw = warp.max(nt);
  int hsize = 0;
  MYB_NOUNROLL
  for (int j = 0; j < nw; j++) {
    if (j < nt) {
      const int w = S.at(S.kCnt, S.at(S.kOrd, j));
      hsize++;
      heap_sift_up(S, hsize - 1, j, w);
    }
    warp.sync();
  }
  int nnode = nt;
  MYB_NOUNROLL
  for (int t = 0; t + 1 < nw; t++) {  // Huffman.cpp:210-217: n - 1 merges
    if (t + 1 < nt) {
      int wl, wr;
      const int l = heap_pop(S, hsize, wl);
      const int r = heap_pop(S, hsize, wr);
      const int w = wl + wr;
      S.at(S.kPar, l) = (uint8_t)nnode;
      S.at(S.kPar, r) = (uint8_t)nnode;
      hsize++;
      heap_sift_up(S, hsize - 1, nnode, w);
      nnode++;
    }
    warp.sync();
  }
  // ---- code length = leaf depth (Huffman.cpp:71-83); parents always have larger ids than children
  if (tree) S.at(S.kPar, nnode - 1) = 0;
  MYB_NOUNROLL
  for (int k = 2; k <= 2 * nw - 1; k++) {
    const int i = nnode - k;
    if (tree && i >= 0) S.at(S.kPar, i) = (uint8_t)(S.at(S.kPar, S.at(S.kPar, i)) + 1);
  }
  MYB_NOUNROLL
  for (int j = 0; j < nw; j++)
    if (j < nt) S.at(S.kLen, S.at(S.kOrd, j)) = S.at(S.kPar, j);
  // ---- tree_data: lengths ascending, symbols ascending inside a length (Huffman.cpp:76-78).  The keys are distinct, so a
  // slot's place is the number of smaller keys: n^2 compares whose loads do not depend on one another, where an insertion
  // sort walks a chain of dependent loads (this code runs with few warps per SM and is bound by exactly that latency).
  MYB_NOUNROLL
  for (int i = 0; i < nw; i++) {
    if (i < nt) {
      const int key = ((int)S.at(S.kLen, i) << 12) + (S.sym(i) + 2048);
      int rank = 0;
      MYB_UNROLL4
      for (int j = 0; j < nt; j++) rank += ((((int)S.at(S.kLen, j) << 12) + (S.sym(j) + 2048)) < key) ? 1 : 0;
      S.at(S.kSorted, rank) = (uint8_t)i;
    }
    warp.sync();
  }
  // ---- canonical codes (Huffman.cpp:86-103) stored bit-reversed; sizes.  Runs for every block (n <= 2 too).
  const int nc = bail ? 0 : n;
  const int ncw = warp.max(nc);
  int code = 0, prev = 0, bits = 0;
  uint64_t per_len = 0;
  MYB_NOUNROLL
  for (int i = 0; i < ncw; i++) {
    if (i < nc) {
      const int s = S.at(S.kSorted, i);
      const int len = S.at(S.kLen, s);
      code = (code << (len - prev)) & 0xff;
      S.at(S.kCode, s) = (uint8_t)(bit_reverse32((uint32_t)code) >> (32 - len));
      code = (code + 1) & 0xff;
      prev = len;
      per_len += 1ull << (8 * (len - 1));
      bits += len * (int)S.at(S.kCnt, s);
    }
  }
  int table = 0;
  for (int len = 0; len < 8; len++) table += group_table_bytes((int)((per_len >> (8 * len)) & 0xff));
  pl.bits = bits;
  pl.table_bytes = table;
  pl.per_len = per_len;
  return pl;
}

// little-endian bit writer into bytes; put() takes at most 16 bits while fewer than 8 are pending
struct BitSink {
  uint8_t* p;
  uint32_t acc;
  int nb;
  MYB_HD void put(uint32_t v, int len) {
    acc |= v << nb;
    nb += len;
    if (nb >= 8) { *p++ = (uint8_t)acc; acc >>= 8; nb -= 8; }
    if (nb >= 8) { *p++ = (uint8_t)acc; acc >>= 8; nb -= 8; }
  }
  MYB_HD void flush() {
    if (nb > 0) *p++ = (uint8_t)acc;
    acc = 0;
    nb = 0;
  }
};

// Serialise the chunk planned by huff_plan into dst[0 .. pl.size()).  Lanes without a chunk pass pl.n = 0.
template <int CAP, int STRIDE, class Z, class W>
MYB_HD void huff_emit(Z& z, const HuffPlan& pl, const HuffScratch<CAP, STRIDE>& S, uint8_t* dst, const W& warp) {
  const int n = pl.n > 0 ? pl.n : 0;
  if (n > 0) {
    dst[0] = (uint8_t)(pl.bits & 0xff);
    dst[1] = (uint8_t)(pl.bits >> 8);
    dst[2] = (uint8_t)pl.table_bytes;
  }
  BitSink w{dst + 3, 0, 0};
  // code table, Huffman.cpp:300-316: symbols in (length, value) order; a group header before the first symbol of
  // a length and after every 32 symbols of the same length; every group is padded to whole bytes
  const int nw = warp.max(n);
  int run_len = 0, in_run = 0;
  MYB_NOUNROLL
  for (int i = 0; i < nw; i++) {
    if (i < n) {
      const int s = S.at(S.kSorted, i);
      const int len = S.at(S.kLen, s);
      if (len != run_len) { run_len = len; in_run = 0; }
      if ((in_run & 31) == 0) {
        w.flush();
        const int left = (int)((pl.per_len >> (8 * (len - 1))) & 0xff) - in_run;
        w.put((uint32_t)(((len - 1) << 5) | ((left > 32 ? 32 : left) - 1)), 8);
      }
      w.put((uint32_t)S.sym(s) & 0x7ffu, 11);  // pack11bit :36-52
      in_run++;
    }
  }
  w.flush();
  const int L = n > 0 ? pl.msg_len : 0;
  const int Lw = warp.max(L);
  MYB_NOUNROLL
  for (int k = 0; k < Lw; k++) {  // code stream, Huffman.cpp:227-236, :319-325
    if (k < L) {
      const int s = z.get(k);
      w.put(S.at(S.kCode, s), S.at(S.kLen, s));
    }
  }
  w.flush();
}

// ---------------------------------------------------------------------------------------------------
// Fast path: blocks with at most 15 distinct symbols (all blocks of typical content up to q ~ 90; natural images at
// q 50 have 2..8).  Same results as huff_plan / huff_emit, organised for few instructions per block:
//  * histogram through a 32-entry open-addressing table, one 32-bit word per slot (symbol << 16 | count);
//  * up to 13 keys the reference's unordered_map never rehashes (13 buckets) and its iteration order is one SWAR
//    list insertion per key on 4-bit fields; 14..16 keys (one rehash to 29 buckets) replay the list rules on a small
//    byte list; the heap of the initial leaves, the (length, value) sort and the canonical codes are straight-line
//    code on registers (a 15-input sorting network);
//  * heap entries carry weight << 16 | leaf set, so a merge needs no parent links: it adds 1 to the depth
//    nibble of every leaf below it, and the code-stream size is the sum of the merged weights.
// Every loop over slots is unrolled with a warp-uniform guard (k < warp maximum of n), so a warp of 4-symbol blocks
// does not pay for the capacity.
// ---------------------------------------------------------------------------------------------------
constexpr int kFastCap = 15;   // distinct symbols of the fast path = the most huff_hist counts
template <int STRIDE>
struct FastScratch {
  uint32_t* sc;    // [16][STRIDE] + lane: symbol << 16 | occurrences, slots in first-occurrence order; after the code
                   //      assignment the low half holds length << 8 | bit-reversed code instead of the count
  uint8_t* aux;    // 64 * STRIDE bytes shared by the lanes, viewed as
                   //   uint16[32][STRIDE]  hash table (value & 0x7ff) << 4 | slot, 0xffff = empty; all empty when huff_hist starts
                   //   uint8[60][STRIDE]   three byte lists of the rehash replay
                   //   uint32[16][STRIDE]  the heap
  int lane;        // this thread's column
  MYB_HD uint32_t& slot(int s) const { return sc[s * STRIDE]; }
  MYB_HD uint16_t& codeword(int s) const { return reinterpret_cast<uint16_t*>(sc + s * STRIDE)[0]; }  // low half (little endian)
  MYB_HD uint16_t& tab(int h) const { return reinterpret_cast<uint16_t*>(aux)[h * STRIDE + lane]; }
  MYB_HD uint32_t& heap(int i) const { return reinterpret_cast<uint32_t*>(aux)[i * STRIDE + lane]; }
  MYB_HD uint8_t& lst(int a, int i) const { return aux[(a * 20 + i) * STRIDE + lane]; }
};

// Histogram of the message z[0 .. L) in first-occurrence order.  Z here is an accessor of raw 16-bit words:
// raw(i) / setraw(i, w).  A coefficient needs 11 bits, so the slot of its value is written into bits 11..14 of
// the same word and the value stays readable (sign-extend the low 11 bits) for the code that takes over when
// the block has more symbols than this path handles.
// Returns the number of distinct symbols, or -1 when there are more than CAP (<= kFastCap; a lane stops counting at
// symbol number CAP + 1; the loop runs to the warp's longest message - a vote every fourth step to leave early cost
// more than it saved, 1.6 % of the headline compress time in a same-box A/B).  Idle lanes pass
// live = false and get 0.
template <int CAP, int STRIDE, class Z, class W>
MYB_HD int huff_hist(Z& z, int L, bool live, const FastScratch<STRIDE>& F, const W& warp) {
  static_assert(CAP <= kFastCap, "slot numbers are 4 bits");
  int n = 0;
  if (!live) L = 0;
  const int Lw = warp.max(L);
  MYB_NOUNROLL
  for (int i = 0; i < Lw; i++) {
    if (i < L && n <= CAP) {
      const uint32_t raw = z.raw(i);
      const uint32_t tag = raw & 0x7ffu;
      uint32_t h = raw & 31u;
      uint32_t e = F.tab((int)h);
      if (e != 0xffffu && (e >> 4) != tag) {  // rare: two values of the block share their low five bits
        do {
          h = (h + 1u) & 31u;
          e = F.tab((int)h);
        } while (e != 0xffffu && (e >> 4) != tag);
      }
      const bool isnew = e == 0xffffu;
      const uint32_t s = isnew ? (uint32_t)n : (e & 15u);  // n == 16 cannot get here
      uint32_t word = raw << 16;
      if (isnew) F.tab((int)h) = (uint16_t)((tag << 4) | s);
      else word = F.slot((int)s);
      F.slot((int)s) = word + 1u;
      n += isnew ? 1 : 0;
      z.setraw(i, tag | (s << 11));
    }
    warp.sync();
  }
  if (live && L == 0) {  // all-zero block: the single symbol 0, one bit (Huffman.cpp:195-199)
    F.slot(0) = 1u;
    z.setraw(0, 0u);
    n = 1;
  }
  return n > CAP ? -1 : n;
}

struct FastPlan {
  int n;           // distinct symbols, 0 = idle lane
  int msg_len;     // coded symbols
  int bits;        // code stream bits
  int table_bytes; // bytes of the serialised code table
  uint32_t dlo, dhi;  // code length of slot s in nibble s (slots 0..7, 8..14)
  MYB_HD int size() const { return n > 0 ? 3 + table_bytes + ((bits + 7) >> 3) : 0; }
};

MYB_HD uint32_t spread8(uint32_t m) {  // bit k of m -> bit 4k
  uint32_t x = m;
  x = (x | (x << 12)) & 0x000f000fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}
MYB_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
// position (in bits) of the lowest 4-bit field of w that equals b, or -1
MYB_HD int nib_find(uint32_t w, uint32_t b) {
  const uint32_t x = w ^ (b * 0x11111111u);
  const uint32_t z = (x - 0x11111111u) & ~x & 0x88888888u;  // lowest hit is exact
  return z ? (ctz32(z) & ~3) : -1;
}
// insert the 4-bit value v at bit position p4, moving the higher fields up by one (the top field is dropped)
MYB_HD uint32_t nib_insert(uint32_t w, int p4, uint32_t v) {
  const uint32_t low = (1u << p4) - 1u;
  return (w & low) | ((w & ~low) << 4) | (v << p4);
}

// std::hash<short> = the value sign-extended to 64 bits; 2^64 mod 13 = 3, 2^64 mod 29 = 24
MYB_HD uint32_t bucket13(int v) {
  const uint32_t x = (uint32_t)(v + 1040 + (v < 0 ? 3 : 0));  // >= 0, same residue as the 64-bit hash
  return x - 13u * ((x * 5042u) >> 16);                        // x % 13 for x < 6547
}
MYB_HD uint32_t bucket29(int v) {
  const uint32_t x = (uint32_t)(v + 1044 + (v < 0 ? 24 : 0));  // 1044 = 36 * 29
  return x - 29u * ((x * 2260u) >> 16);                        // x % 29 for x < 3276... checked in tests
}

// std::push_heap of `ent` as element number J of a heap held in registers (J is a compile-time constant, so the
// path to the root is fixed and only the stopping point is data dependent).  Entries compare by weight (bits 16..).
template <int J, int CAP>
MYB_HD void heap_push_static(uint32_t (&H)[CAP], uint32_t ent) {
  const uint32_t key = ent | 0xffffu;
  bool go = true;
  int hole = J;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int lvl = 0; lvl < 3; lvl++) {
    if (hole > 0) {
      const int parent = (hole - 1) >> 1;
      const bool up = go && H[parent] > key;
      H[hole] = up ? H[parent] : (go ? ent : H[hole]);
      go = up;
      hole = parent;
    }
  }
  if (go) H[0] = ent;
}

template <int STRIDE>
MYB_HD void heap32_sift_up(const FastScratch<STRIDE>& F, int hole, uint32_t ent) {
  const uint32_t key = ent | 0xffffu;
  MYB_NOUNROLL
  while (hole > 0) {
    const int parent = (hole - 1) >> 1;
    const uint32_t pe = F.heap(parent);
    if (!(pe > key)) break;
    F.heap(hole) = pe;
    hole = parent;
  }
  F.heap(hole) = ent;
}

// std::pop_heap + pop_back (stl_heap.h __adjust_heap, then __push_heap of the former last element)
template <int STRIDE>
MYB_HD uint32_t heap32_pop(const FastScratch<STRIDE>& F, int& hsize) {
  const uint32_t top = F.heap(0);
  const int len = hsize - 1;
  hsize = len;
  if (len == 0) return top;
  const uint32_t value = F.heap(len);
  int hole = 0, child = 0;
  const int half = (len - 1) >> 1;
  MYB_NOUNROLL
  while (child < half) {
    child = 2 * child + 2;
    uint32_t r = F.heap(child);
    const uint32_t l = F.heap(child - 1);
    if (r > (l | 0xffffu)) {
      child--;
      r = l;
    }
    F.heap(hole) = r;
    hole = child;
  }
  if ((len & 1) == 0 && child == ((len - 2) >> 1)) {
    child = 2 * child + 2;
    F.heap(hole) = F.heap(child - 1);
    hole = child - 1;
  }
  heap32_sift_up(F, hole, value);
  return top;
}

// libstdc++ _M_insert_bucket_begin on byte lists (list a = slots, list b = buckets): a key whose bucket already holds
// nodes goes right before the first node of that bucket's run, otherwise to the front of the whole list.
template <int STRIDE>
MYB_HD void lst_place(const FastScratch<STRIDE>& F, int la, int lb, int& ln, int slot, int bucket) {
  int p = 0;
  MYB_NOUNROLL
  for (int i = 0; i < ln; i++)
    if (F.lst(lb, i) == bucket) { p = i; break; }
  MYB_NOUNROLL
  for (int i = ln; i > p; i--) {
    F.lst(la, i) = F.lst(la, i - 1);
    F.lst(lb, i) = F.lst(lb, i - 1);
  }
  F.lst(la, p) = (uint8_t)slot;
  F.lst(lb, p) = (uint8_t)bucket;
  ln++;
}

// CAP = 8 or 15: capacity of this instantiation; every lane of the warp has n <= CAP (nw = warp maximum of n)
// ROLL15: the 15-symbol instantiation with rolled loops as well.  Pays where the symbol counts of a warp are mixed (the in-place
// build of the coding kernel: synthetic q90 -2.7 %), costs where every block has 9..15 (heavy15_kernel: natural q90 +0.9 %).
template <int CAP, int STRIDE, class W, bool ROLL15 = true>
MYB_HD FastPlan huff_fast_plan_n(int n, int nw, int msg_len, const FastScratch<STRIDE>& F, const W& warp) {
  FastPlan pl;
  pl.n = n;
  pl.msg_len = msg_len;
#ifndef MYB_PLAN_UNROLLED
  constexpr bool kRolled = CAP <= 8;
#else
  constexpr bool kRolled = false;
#endif
  uint32_t count0 = 0;  // occurrences of slot 0 (the whole message when n == 1)
  if constexpr (kRolled) {
    // Up to 8 symbols: 13 buckets, never a rehash, the list fits one word.  Two rolled loops of n steps each -- the list
    // insertions, then the leaves pushed in list order straight into the shared-memory heap -- where the unrolled version ran
    // 4 or 8 predicated copies of each body whatever n was, built the heap in registers and copied it out.
    uint32_t olo = 0, blo = ~0u;  // bucket of each list position; 0xF = empty, which is no bucket
    MYB_NOUNROLL
    for (int k = 0; k < n; k++) {
      const uint32_t b = bucket13((int)(int16_t)(F.slot(k) >> 16));
      const int f = nib_find(blo, b);
      const int p4 = f < 0 ? 0 : f;
      olo = nib_insert(olo, p4, (uint32_t)k);
      blo = nib_insert(blo, p4, b);
    }
    MYB_NOUNROLL
    for (int j = 0; j < n; j++) {  // Huffman.cpp:207-209
      const uint32_t sl = (olo >> (4 * j)) & 15u;
      heap32_sift_up(F, j, ((F.slot((int)sl) & 0xffu) << 16) | (1u << sl));
    }
    if (n > 0) count0 = F.slot(0) & 0xffu;
#ifndef MYB_PLAN_UNROLLED
  } else if constexpr (CAP > 8 && ROLL15) {
    // Up to 15 symbols, the same with rolled loops: the list is two words of 4-bit fields, 14 keys and more replay the
    // reference's rehash on byte lists as before.
    bool has_zero = false;
    MYB_NOUNROLL
    for (int k = 0; k < n; k++)
      if ((F.slot(k) >> 16) == 0) has_zero = true;
    // The reference's map also holds the key 0 while it is filled (trailing zeros or freq[0], Huffman.cpp:192-195); when
    // the message has no zero it is inserted last and erased again (:201), which only matters when it is key number 14
    // and triggers the rehash to 29 buckets.
    const int keys = n + ((n > 0 && !has_zero) ? 1 : 0);
    const bool rehash = warp.any(keys >= 14);
    uint32_t olo = 0, ohi = 0;
    if (!rehash) {
      uint32_t blo = ~0u, bhi = ~0u;  // bucket of each list position; 0xF = empty, which is no bucket
      MYB_NOUNROLL
      for (int k = 0; k < n; k++) {
        const uint32_t b = bucket13((int)(int16_t)(F.slot(k) >> 16));
        int f = nib_find(blo, b);
        int p = f >= 0 ? (f >> 2) : -1;
        if (p < 0) {
          f = nib_find(bhi, b);
          p = f >= 0 ? 8 + (f >> 2) : 0;
        }
        if (p < 8) {
          ohi = (ohi << 4) | (olo >> 28);
          bhi = (bhi << 4) | (blo >> 28);
          olo = nib_insert(olo, 4 * p, (uint32_t)k);
          blo = nib_insert(blo, 4 * p, b);
        } else {
          ohi = nib_insert(ohi, 4 * (p - 8), (uint32_t)k);
          bhi = nib_insert(bhi, 4 * (p - 8), b);
        }
      }
    } else {
      // 14..16 keys: replay the list rules on byte lists (list 0 = slots, 1 = buckets, 2 = copy)
      const int m = (n == 13 && !has_zero) ? 14 : n;  // the appended key 0 only matters as key number 14
      const int mw = warp.max(m);
      int ln = 0;
      MYB_NOUNROLL
      for (int s = 0; s < mw; s++) {
        if (s < m) {
          if (s == 13) {  // _M_rehash_aux: walk the old list front to back and re-place every node among 29 buckets
            for (int i = 0; i < 13; i++) F.lst(2, i) = F.lst(0, i);
            ln = 0;
            for (int i = 0; i < 13; i++) {
              const int t = F.lst(2, i);
              lst_place(F, 0, 1, ln, t, (int)bucket29((int)(int16_t)(F.slot(t) >> 16)));
            }
          }
          const int v = s < n ? (int)(int16_t)(F.slot(s) >> 16) : 0;
          lst_place(F, 0, 1, ln, s, (int)(s < 13 ? bucket13(v) : bucket29(v)));
        }
        warp.sync();
      }
      // to the nibble lists, dropping the appended key (slot number n) again
      int wpos = 0;
      MYB_NOUNROLL
      for (int i = 0; i < mw; i++) {
        if (i < m) {
          const uint32_t t = F.lst(0, i);
          if ((int)t < n) {
            if (wpos < 8) olo |= t << (4 * wpos);
            else ohi |= t << (4 * (wpos - 8));
            wpos++;
          }
        }
      }
      warp.sync();
    }
    MYB_NOUNROLL
    for (int j = 0; j < n; j++) {  // Huffman.cpp:207-209
      const uint32_t sl = ((j < 8 ? olo >> (4 * (j & 7)) : ohi >> (4 * (j & 7)))) & 15u;
      heap32_sift_up(F, j, ((F.slot((int)sl) & 0xffu) << 16) | (1u << sl));
    }
    if (n > 0) count0 = F.slot(0) & 0xffu;
#endif
  } else {
  // ---- slot words into registers; is the value 0 part of the message?
  uint32_t sw[CAP];
  bool has_zero = false;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < CAP; k++) sw[k] = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++) {
    if (k < n) {
      sw[k] = F.slot(k);
      if ((sw[k] >> 16) == 0) has_zero = true;
    }
  }
  if (CAP > 8) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 8; k < CAP; k++) {
      if (k < n) {
        sw[k] = F.slot(k);
        if ((sw[k] >> 16) == 0) has_zero = true;
      }
    }
  }
  // The reference's map also holds the key 0 while it is filled (trailing zeros or freq[0], Huffman.cpp:192-195); when
  // the message has no zero it is inserted last and erased again (:201), which only matters when it is key number 14
  // and triggers the rehash to 29 buckets.
  const int keys = n + ((n > 0 && !has_zero) ? 1 : 0);
  const bool rehash = CAP > 8 && warp.any(keys >= 14);
  // ---- iteration order of the reference's map as a list of slots, 4 bits each (olo: positions 0..7, ohi: 8..14)
  uint32_t olo = 0, ohi = 0;
  if (!rehash) {
    uint32_t blo = ~0u, bhi = ~0u;  // bucket of each list position; 0xF = empty, which is no bucket
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 8; k++) {  // the list still fits the low word
      if (k < n) {
        const uint32_t b = bucket13((int)(int16_t)(sw[k] >> 16));
        const int f = nib_find(blo, b);
        const int p4 = f < 0 ? 0 : f;
        olo = nib_insert(olo, p4, (uint32_t)k);
        blo = nib_insert(blo, p4, b);
      }
    }
    if (CAP > 8) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int k = 8; k < (CAP < 13 ? CAP : 13); k++) {
        if (k < n) {
          const uint32_t b = bucket13((int)(int16_t)(sw[k] >> 16));
          int f = nib_find(blo, b);
          int p = f >= 0 ? (f >> 2) : -1;
          if (p < 0) {
            f = nib_find(bhi, b);
            p = f >= 0 ? 8 + (f >> 2) : 0;
          }
          if (p < 8) {
            ohi = (ohi << 4) | (olo >> 28);
            bhi = (bhi << 4) | (blo >> 28);
            olo = nib_insert(olo, 4 * p, (uint32_t)k);
            blo = nib_insert(blo, 4 * p, b);
          } else {
            ohi = nib_insert(ohi, 4 * (p - 8), (uint32_t)k);
            bhi = nib_insert(bhi, 4 * (p - 8), b);
          }
        }
      }
    }
  } else {
    // 14..16 keys somewhere in the warp: replay the list rules on byte lists (list 0 = slots, 1 = buckets, 2 = copy)
    const int m = (n == 13 && !has_zero) ? 14 : n;  // the appended key 0 only matters as key number 14
    const int mw = warp.max(m);
    int ln = 0;
    MYB_NOUNROLL
    for (int s = 0; s < mw; s++) {
      if (s < m) {
        if (s == 13) {  // _M_rehash_aux: walk the old list front to back and re-place every node among 29 buckets
          for (int i = 0; i < 13; i++) F.lst(2, i) = F.lst(0, i);
          ln = 0;
          for (int i = 0; i < 13; i++) {
            const int t = F.lst(2, i);
            lst_place(F, 0, 1, ln, t, (int)bucket29((int)(int16_t)(F.slot(t) >> 16)));
          }
        }
        const int v = s < n ? (int)(int16_t)(F.slot(s) >> 16) : 0;
        lst_place(F, 0, 1, ln, s, (int)(s < 13 ? bucket13(v) : bucket29(v)));
      }
      warp.sync();
    }
    // to the nibble lists, dropping the appended key (slot number n) again
    int wpos = 0;
    MYB_NOUNROLL
    for (int i = 0; i < mw; i++) {
      if (i < m) {
        const uint32_t t = F.lst(0, i);
        if ((int)t < n) {
          if (wpos < 8) olo |= t << (4 * wpos);
          else ohi |= t << (4 * (wpos - 8));
          wpos++;
        }
      }
    }
    warp.sync();
  }
  // ---- leaves pushed in list order (Huffman.cpp:207-209) into a heap held in registers
  uint32_t H[CAP];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < CAP; k++) H[k] = 0;
#define MYB_PUSH_LEAF(J)                                                                 \
  if (J < n) {                                                                           \
    const uint32_t sl = ((J < 8 ? olo >> (4 * (J & 7)) : ohi >> (4 * (J & 7)))) & 15u;   \
    heap_push_static<J, CAP>(H, ((F.slot((int)sl) & 0xffu) << 16) | (1u << sl));         \
  }
  MYB_PUSH_LEAF(0) MYB_PUSH_LEAF(1) MYB_PUSH_LEAF(2) MYB_PUSH_LEAF(3)
  if (nw > 4) { MYB_PUSH_LEAF(4) MYB_PUSH_LEAF(5) MYB_PUSH_LEAF(6) MYB_PUSH_LEAF(7) }
  if constexpr (CAP > 8) {
    MYB_PUSH_LEAF(8) MYB_PUSH_LEAF(9) MYB_PUSH_LEAF(10) MYB_PUSH_LEAF(11)
    MYB_PUSH_LEAF(12) MYB_PUSH_LEAF(13) MYB_PUSH_LEAF(14)
  }
#undef MYB_PUSH_LEAF
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++)
    if (k < nw) F.heap(k) = H[k];
  if constexpr (CAP > 8) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 8; k < CAP; k++) F.heap(k) = H[k];
  }
    count0 = sw[0] & 0xffu;
  }
  // ---- n - 1 merges (Huffman.cpp:210-217)
  int hsize = n, bits = 0;
  uint32_t dlo = 0, dhi = 0;
  MYB_NOUNROLL
  for (int t = 0; t + 1 < nw; t++) {
    if (t + 1 < n) {
      const uint32_t a = heap32_pop(F, hsize);
      const uint32_t b = heap32_pop(F, hsize);
      const uint32_t w = (a >> 16) + (b >> 16);
      const uint32_t m = (a | b) & 0x7fffu;
      dlo += spread8(m & 0xffu);
      if (CAP > 8) dhi += spread8(m >> 8);
      bits += (int)w;
      hsize++;
      heap32_sift_up(F, hsize - 1, (w << 16) | m);
    }
    warp.sync();
  }
  if (n == 1) {  // a single symbol gets a one-bit code (Huffman.cpp:76, :218-221)
    dlo = 1;
    bits = (int)count0;
  }
  pl.dlo = dlo;
  pl.dhi = dhi;
  pl.bits = bits;
  // ---- size of the code table: one group per used length, 1 + ceil(11 c / 8) bytes for c symbols (Huffman.cpp:284-293);
  // lengths are 1..8 (64 coefficients cannot make a deeper tree), at most 15 symbols per length
  uint32_t per_len = 0, len8 = 0;  // symbols per length 0..7 in 4-bit fields (length 0 = unused slots, ignored); length 8
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++) per_len += 1u << (((dlo >> (4 * k)) & 15u) * 4u);  // at most 8 symbols: lengths <= 7
  if constexpr (CAP > 8) {
    per_len = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < CAP; k++) {
      const uint32_t d = ((k < 8 ? dlo >> (4 * (k & 7)) : dhi >> (4 * (k & 7)))) & 15u;
      per_len += d < 8u ? 1u << (4u * d) : 0u;
      len8 += d == 8u ? 1u : 0u;
    }
  }
  uint32_t lo = per_len & 0xfff0u, hi = per_len >> 16;
  lo = (lo | (lo << 8)) & 0x00ff00ffu;
  lo = (lo | (lo << 4)) & 0x0f0f0f0fu;
  hi = (hi | (hi << 8)) & 0x00ff00ffu;
  hi = (hi | (hi << 4)) & 0x0f0f0f0fu;
  uint32_t f = 0;
  {
    const uint32_t t = ((lo * 3u + 0x07070707u) >> 3) & 0x1f1f1f1fu;
    const uint32_t nz = ((lo + 0x7f7f7f7fu) >> 7) & 0x01010101u;
    f += lo + t + nz;
  }
  {
    const uint32_t t = ((hi * 3u + 0x07070707u) >> 3) & 0x1f1f1f1fu;
    const uint32_t nz = ((hi + 0x7f7f7f7fu) >> 7) & 0x01010101u;
    f += hi + t + nz;
  }
  pl.table_bytes = (int)((f * 0x01010101u) >> 24) + (len8 ? (int)(1u + len8 + ((3u * len8 + 7u) >> 3)) : 0);
  return pl;
}

// Two instantiations, chosen per warp: the compact one when no block of the warp has more than 8 symbols.
template <bool ROLL15 = true, int STRIDE, class W>
MYB_HD FastPlan huff_fast_plan(int n, int msg_len, const FastScratch<STRIDE>& F, const W& warp) {
  const int nw = warp.umax(n);
  if (nw <= 8) return huff_fast_plan_n<8>(n, nw, msg_len, F, warp);
  return huff_fast_plan_n<kFastCap, STRIDE, W, ROLL15>(n, nw, msg_len, F, warp);
}

#define MYB_CSWAP(a, b)                          \
  {                                              \
    const uint32_t lo_ = a < b ? a : b;          \
    b = a < b ? b : a;                           \
    a = lo_;                                     \
  }

// Serialise the chunk planned by huff_fast_plan into dst[0 .. pl.size()): header, code table in (length, value)
// order with canonical codes assigned on the way (Huffman.cpp:86-103, :300-316), code stream (:227-236, :319-325).
template <int CAP, int STRIDE, class Z, class W>
MYB_HD void huff_fast_emit_n(Z& z, const FastPlan& pl, int nw, const FastScratch<STRIDE>& F, uint8_t* dst, const W& warp) {
  const int n = pl.n;
  if (n > 0) {
    dst[0] = (uint8_t)(pl.bits & 0xff);
    dst[1] = (uint8_t)(pl.bits >> 8);
    dst[2] = (uint8_t)pl.table_bytes;
  }
  // sort keys: length << 16 | (value + 1024) << 4 | slot; unused slots sort last
  uint32_t K[CAP];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < CAP; k++) K[k] = 0xffffffffu;
#define MYB_KEY(k)                                                                                        \
  if (k < n) {                                                                                            \
    const int v = (int)(int16_t)(F.slot(k) >> 16);                                                        \
    const uint32_t d = ((k < 8 ? pl.dlo >> (4 * (k & 7)) : pl.dhi >> (4 * (k & 7)))) & 15u;               \
    K[k] = (d << 16) | ((uint32_t)(v + 1024) << 4) | (uint32_t)k;                                         \
  }
  MYB_KEY(0) MYB_KEY(1) MYB_KEY(2) MYB_KEY(3)
  if (nw > 4) { MYB_KEY(4) MYB_KEY(5) MYB_KEY(6) MYB_KEY(7) }
  if constexpr (CAP > 8) { MYB_KEY(8) MYB_KEY(9) MYB_KEY(10) MYB_KEY(11) MYB_KEY(12) MYB_KEY(13) MYB_KEY(14) }
#undef MYB_KEY
  // Batcher's odd-even merge sort for 16 inputs without the exchanges that touch input 15; its first 19 exchanges sort K[0..7]
  MYB_CSWAP(K[0], K[1]) MYB_CSWAP(K[2], K[3]) MYB_CSWAP(K[0], K[2]) MYB_CSWAP(K[1], K[3]) MYB_CSWAP(K[1], K[2])
  if (nw > 4) {
    MYB_CSWAP(K[4], K[5]) MYB_CSWAP(K[6], K[7]) MYB_CSWAP(K[4], K[6]) MYB_CSWAP(K[5], K[7]) MYB_CSWAP(K[5], K[6])
    MYB_CSWAP(K[0], K[4]) MYB_CSWAP(K[2], K[6]) MYB_CSWAP(K[2], K[4]) MYB_CSWAP(K[1], K[5]) MYB_CSWAP(K[3], K[7])
    MYB_CSWAP(K[3], K[5]) MYB_CSWAP(K[1], K[2]) MYB_CSWAP(K[3], K[4]) MYB_CSWAP(K[5], K[6])
  }
  if constexpr (CAP > 8) {
    MYB_CSWAP(K[8], K[9]) MYB_CSWAP(K[10], K[11]) MYB_CSWAP(K[8], K[10]) MYB_CSWAP(K[9], K[11]) MYB_CSWAP(K[9], K[10])
    MYB_CSWAP(K[12], K[13]) MYB_CSWAP(K[12], K[14]) MYB_CSWAP(K[13], K[14]) MYB_CSWAP(K[8], K[12]) MYB_CSWAP(K[10], K[14])
    MYB_CSWAP(K[10], K[12]) MYB_CSWAP(K[9], K[13]) MYB_CSWAP(K[11], K[13]) MYB_CSWAP(K[9], K[10]) MYB_CSWAP(K[11], K[12])
    MYB_CSWAP(K[13], K[14]) MYB_CSWAP(K[0], K[8]) MYB_CSWAP(K[4], K[12]) MYB_CSWAP(K[4], K[8]) MYB_CSWAP(K[2], K[10])
    MYB_CSWAP(K[6], K[14]) MYB_CSWAP(K[6], K[10]) MYB_CSWAP(K[2], K[4]) MYB_CSWAP(K[6], K[8]) MYB_CSWAP(K[10], K[12])
    MYB_CSWAP(K[1], K[9]) MYB_CSWAP(K[5], K[13]) MYB_CSWAP(K[5], K[9]) MYB_CSWAP(K[3], K[11]) MYB_CSWAP(K[7], K[11])
    MYB_CSWAP(K[3], K[5]) MYB_CSWAP(K[7], K[9]) MYB_CSWAP(K[11], K[13]) MYB_CSWAP(K[1], K[2]) MYB_CSWAP(K[3], K[4])
    MYB_CSWAP(K[5], K[6]) MYB_CSWAP(K[7], K[8]) MYB_CSWAP(K[9], K[10]) MYB_CSWAP(K[11], K[12]) MYB_CSWAP(K[13], K[14])
  }
  uint8_t* p = dst + 3;
  uint8_t* hdr = dst;  // header byte of the open group (dst: none yet)
  uint32_t acc = 0, code = 0;
  int nb = 0, prev = 0, cnt = 0;
#ifndef MYB_TABLE_UNROLLED
  // The sorted keys go through the (now free) heap words so that the table can be written by a ROLLED loop of n steps: the
  // unrolled version ran 4 or 8 predicated copies of this body whatever n was, and was 360 instructions of code.
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < CAP; k++)
    if (k < nw) F.heap(k) = K[k];
  MYB_NOUNROLL
  for (int i = 0; i < n; i++) {
    const uint32_t key = F.heap(i);
    const int len = (int)(key >> 16);
    code <<= (len - prev);
    /* the slot word keeps its symbol and gets length << 8 | bit-reversed code in place of the count */
    const int sl = (int)(key & 15u);
    F.codeword(sl) = (uint16_t)((bit_reverse32(code) >> (32 - len)) | ((uint32_t)len << 8));
    code++;
    if (len != prev) { /* a new group: close the previous one, pad to a whole byte, reserve the header byte */
      if (prev) *hdr = (uint8_t)(((prev - 1) << 5) | (cnt - 1));
      if (nb > 0) *p++ = (uint8_t)acc;
      acc = 0;
      nb = 0;
      hdr = p++;
      cnt = 0;
      prev = len;
    }
    /* pack11bit (Huffman.cpp:36-52): 11 bits on top of nb < 8 pending ones always complete one byte, sometimes two */
    acc |= (((key >> 4) + 1024u) & 0x7ffu) << nb;
    nb += 11;
    *p++ = (uint8_t)acc;
    acc >>= 8;
    nb -= 8;
    if (nb >= 8) {
      *p++ = (uint8_t)acc;
      acc >>= 8;
      nb -= 8;
    }
    cnt++;
  }
#else
#define MYB_TABLE_SYMBOL(i)                                                                                              \
  if (i < n) {                                                                                                           \
    const uint32_t key = K[i];                                                                                           \
    const int len = (int)(key >> 16);                                                                                    \
    code <<= (len - prev);                                                                                               \
    /* the slot word keeps its symbol and gets length << 8 | bit-reversed code in place of the count */                 \
    const int sl = (int)(key & 15u);                                                                                     \
    F.codeword(sl) = (uint16_t)((bit_reverse32(code) >> (32 - len)) | ((uint32_t)len << 8));                             \
    code++;                                                                                                              \
    if (len != prev) { /* a new group: close the previous one, pad to a whole byte, reserve the header byte */          \
      if (prev) *hdr = (uint8_t)(((prev - 1) << 5) | (cnt - 1));                                                         \
      if (nb > 0) *p++ = (uint8_t)acc;                                                                                   \
      acc = 0;                                                                                                           \
      nb = 0;                                                                                                            \
      hdr = p++;                                                                                                         \
      cnt = 0;                                                                                                           \
      prev = len;                                                                                                        \
    }                                                                                                                    \
    /* pack11bit (Huffman.cpp:36-52): 11 bits on top of nb < 8 pending ones always complete one byte, sometimes two */   \
    acc |= (((key >> 4) + 1024u) & 0x7ffu) << nb;                                                                        \
    nb += 11;                                                                                                            \
    *p++ = (uint8_t)acc;                                                                                                 \
    acc >>= 8;                                                                                                           \
    nb -= 8;                                                                                                             \
    if (nb >= 8) {                                                                                                       \
      *p++ = (uint8_t)acc;                                                                                               \
      acc >>= 8;                                                                                                         \
      nb -= 8;                                                                                                           \
    }                                                                                                                    \
    cnt++;                                                                                                               \
  }
  MYB_TABLE_SYMBOL(0) MYB_TABLE_SYMBOL(1) MYB_TABLE_SYMBOL(2) MYB_TABLE_SYMBOL(3)
  if (nw > 4) { MYB_TABLE_SYMBOL(4) MYB_TABLE_SYMBOL(5) MYB_TABLE_SYMBOL(6) MYB_TABLE_SYMBOL(7) }
  if constexpr (CAP > 8) {
    MYB_TABLE_SYMBOL(8) MYB_TABLE_SYMBOL(9) MYB_TABLE_SYMBOL(10) MYB_TABLE_SYMBOL(11)
    MYB_TABLE_SYMBOL(12) MYB_TABLE_SYMBOL(13) MYB_TABLE_SYMBOL(14)
  }
#undef MYB_TABLE_SYMBOL
#endif
  if (n > 0) {
    *hdr = (uint8_t)(((prev - 1) << 5) | (cnt - 1));
    if (nb > 0) *p++ = (uint8_t)acc;
  }
  acc = 0;
  nb = 0;
  warp.sync();
  const int L = n > 0 ? pl.msg_len : 0;
  const int Lw = warp.max(L);
  MYB_NOUNROLL
  for (int k = 0; k < Lw; k++) {
    if (k < L) {
      const uint32_t ce = F.codeword(z.slot(k));
      acc |= (ce & 0xffu) << nb;
      nb += (int)(ce >> 8);
      if (nb >= 8) {
        *p++ = (uint8_t)acc;
        acc >>= 8;
        nb -= 8;
      }
    }
  }
  if (nb > 0) *p++ = (uint8_t)acc;
}

template <int STRIDE, class Z, class W>
MYB_HD void huff_fast_emit(Z& z, const FastPlan& pl, const FastScratch<STRIDE>& F, uint8_t* dst, const W& warp) {
  const int nw = warp.umax(pl.n);
  if (nw <= 8) huff_fast_emit_n<8>(z, pl, nw, F, dst, warp);
  else huff_fast_emit_n<kFastCap>(z, pl, nw, F, dst, warp);
}
#undef MYB_CSWAP

// ---------------------------------------------------------------------------------------------------
// Decoder (Huffman.cpp:243-277, :54-69, :106-154).  chunk/size: one block's bytes.  emit(j, v) receives
// the value v of zigzag position j for every decoded symbol (positions never emitted are 0).
// Returns 0, or non-zero for the conditions on which the reference throws "Huffman bad code" /
// "Huffman unknown symbol" and for reads the reference would do outside the chunk.
// With WarpLockstep all 32 lanes call it together (idle lanes pass size = 0 and ignore the result).
// ---------------------------------------------------------------------------------------------------
MYB_HD int table_symbol(const uint8_t* groups, int table_bytes, int len, int idx, int* out) {
  int i = 0;
  while (i < table_bytes) {  // groups of the same length are concatenated in file order (Huffman.cpp:258-266)
    const int info = groups[i];
    const int glen = (info >> 5) + 1, c = (info & 31) + 1;
    if (glen == len) {
      if (idx < c) {
        const int bit = idx * 11, byte = i + 1 + (bit >> 3);
        uint32_t v = groups[byte] | ((uint32_t)groups[byte + 1] << 8);
        if ((bit & 7) > 5) v |= (uint32_t)groups[byte + 2] << 16;
        v = (v >> (bit & 7)) & 0x7ffu;
        *out = (v >= 1024u) ? (int)v - 2048 : (int)v;
        return 0;
      }
      idx -= c;
    }
    i += 1 + ((c * 11 + 7) >> 3);
  }
  return 1;
}

template <class Emit, class W>
MYB_HD int huff_decode_block(const uint8_t* chunk, int size, Emit&& emit, const W& warp) {
  int err = 0;
  int bits = 0, table_bytes = 0;
  if (size >= 3) {
    bits = chunk[0] | (chunk[1] << 8);
    table_bytes = chunk[2];
    if (bits > 512 || 3 + table_bytes + ((bits + 7) >> 3) > size) err = 1;
  } else if (size > 0) {
    err = 1;
  }
  if (err || size == 0) { bits = 0; table_bytes = 0; }
  const uint8_t* groups = chunk + 3;
  // code table: symbols per length (one byte each) and, when the groups come in non-decreasing length order
  // with at most 32 symbols per length (every stream the reference writes for <= 32 symbols of a length),
  // the byte offset of each length's single group, so a symbol is one 11-bit extract away.
  uint64_t counts = 0, goff = 0;
  bool direct = true;
  {
    int i = 0, last_len = 0;
    const int tw = warp.max(table_bytes);
    for (int guard = 0; guard < tw; guard++) {  // at most one group per iteration, each group is >= 3 bytes
      if (i < table_bytes && !err) {
        const int info = groups[i];
        const int len = (info >> 5) + 1, c = (info & 31) + 1;
        const int sh = 8 * (len - 1);
        if (len <= last_len) direct = false;
        last_len = len;
        goff |= (uint64_t)(i + 1) << sh;
        i += 1 + ((c * 11 + 7) >> 3);
        if (i > table_bytes || ((counts >> sh) & 0xff) + (uint64_t)c > 64) err = 1;  // more symbols than a block holds
        counts += (uint64_t)c << sh;
      }
      if (!warp.any(i < table_bytes && !err)) break;
    }
  }
  if (err) bits = 0;
  const uint8_t* data = groups + table_bytes;
  // bit reader: `acc` holds `have` not yet consumed bits of the stream, LSB first
  uint32_t acc = 0;
  int have = 0, nbyte = 0, p = 0, j = 0;
  const int data_bytes = (bits + 7) >> 3;
  while (warp.any(p < bits && j < 64)) {
    if (p < bits && j < 64) {
      if (have <= 8 && nbyte < data_bytes) { acc |= (uint32_t)data[nbyte++] << have; have += 8; }
      if (have <= 8 && nbyte < data_bytes) { acc |= (uint32_t)data[nbyte++] << have; have += 8; }
      uint32_t code = 0, first = 0;  // uint8_t in the reference (Huffman.cpp:107-108): keep the 8-bit wrap
      int len = 1, found = 0;
        for (; len <= 8; len++) {
        const uint32_t c = (uint32_t)(counts >> (8 * (len - 1))) & 0xff;
        if (p + len - 1 >= bits) break;  // "Huffman bad code" :120-122
        code |= (acc >> (len - 1)) & 1u;
        if (code < c + first) { found = 1; break; }
        first = ((first + c) << 1) & 0xff;
        code = (code << 1) & 0xff;
      }
      if (!found) {
        err = 1;  // ran out of bits, or "Huffman unknown symbol" :139
        bits = 0;
      } else {
        int v = 0;
        const int idx = (int)(code - first);
        if (direct) {
          const int bit = idx * 11, byte = (int)((goff >> (8 * (len - 1))) & 0xff) + (bit >> 3);
          uint32_t t = groups[byte] | ((uint32_t)groups[byte + 1] << 8);
          if ((bit & 7) > 5) t |= (uint32_t)groups[byte + 2] << 16;
          t = (t >> (bit & 7)) & 0x7ffu;
          v = (t >= 1024u) ? (int)t - 2048 : (int)t;
        } else if (table_symbol(groups, table_bytes, len, idx, &v)) {
          err = 1;
          bits = 0;
        }
        if (!err) {
          emit(j, v);
          j++;
          p += len;
          acc >>= len;
          have -= len;
        }
      }
    }
    warp.sync();
  }
  return err;
}

// ---------------------------------------------------------------------------------------------------
// Fast decoder for the chunks the reference writes for blocks with at most 32 distinct symbols: one group per code
// length, lengths strictly increasing, a prefix code that is not over-subscribed.  Anything else (and every
// malformed table) is handed to huff_decode_block, which follows the reference's decoder step by step.
//  * the code table is unpacked once into an array of symbols in canonical order;
//  * a code is decoded without a bit loop: the next 8 stream bits, MSB first, are compared with the left-aligned
//    end of every length's code range (lim[l] = (first_l + count_l) << (8 - l), non-decreasing in l), the number of
//    ranges passed is the length, and the symbol index is  code + (symbols before this length - first code);
//  * the stream is read through a 32-bit window that is reloaded only after 24 consumed bits.
// ---------------------------------------------------------------------------------------------------
constexpr int kDecFastSyms = 32;
template <int STRIDE>
struct DecScratch {
  int16_t* symtab;  // [32] symbols in canonical order
  int16_t* base;    // [8]  index of the first symbol of length l+1 minus its first code
  MYB_HD int16_t& sym(int k) const { return symtab[k * STRIDE]; }
  MYB_HD int16_t& bs(int l) const { return base[l * STRIDE]; }
};

template <class BP>
MYB_HD uint32_t load_window(BP data, int byte0, int data_bytes) {
  uint32_t w = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int t = 0; t < 4; t++)
    if (byte0 + t < data_bytes) w |= (uint32_t)data[byte0 + t] << (8 * t);
  return w;
}

// State of the stream reader: rwin holds 32 stream bits starting at byte byte0 of the stream, the first one in bit 31;
// sh bits of it are consumed, rem = stream bits from the start of the window to the end of the stream.
struct DecStream {
  uint32_t rwin;
  int sh, rem, j, err, byte0;
};

// PAIRS = code lengths of the table / 2, rounded up to 1, 2 or 4 (warp uniform): how many of the range compares are needed
template <int PAIRS, int STRIDE, class BP, class Emit, class W>
MYB_HD void decode_stream(DecStream& st, const uint32_t (&kk)[4], int maxlen, BP data, int data_bytes,
                          const DecScratch<STRIDE>& D, Emit& emit, const W& warp) {
  // A plain per-lane loop: lanes that run out of symbols wait where the hardware reconverges the warp, behind the loop.
  // Holding the lanes in step by hand (a vote per symbol, an inner test, a warp barrier) cost 4-5 % of the decoder's time.
  while (st.sh < st.rem && st.j < 64) {
    {
      if (st.sh > 24) {  // reload the window at the byte that holds the next bit
        st.byte0 += st.sh >> 3;
        st.rem -= st.sh & ~7;
        st.sh &= 7;
        st.rwin = bit_reverse32(load_window(data, st.byte0, data_bytes));
      }
      const uint32_t r = (st.rwin << st.sh) >> 24;  // next 8 bits, first stream bit on top
      const uint32_t r2 = r * 0x10001u;
      uint32_t acc = (r2 + kk[0]) & 0x01000100u;
      if (PAIRS > 1) acc += (r2 + kk[1]) & 0x01000100u;
      if (PAIRS > 2) acc += ((r2 + kk[2]) & 0x01000100u) + ((r2 + kk[3]) & 0x01000100u);
      const int len = 1 + (int)((acc * 0x10001u) >> 24);  // both 16-bit fields summed: ranges passed
      if (len > maxlen || st.sh + len > st.rem) {
        st.err = 1;  // "Huffman unknown symbol" :139, or the stream ends inside a code: "Huffman bad code" :120-122
        st.rem = 0;
      } else {
        const int idx = (int)(r >> (8 - len)) + D.bs(len - 1);
        emit(st.j, (int)D.sym(idx));
        st.j++;
        st.sh += len;
      }
    }
  }
  warp.sync();
}

// The code-stream loop as a replaceable part of huff_decode_fast: kernels.cu passes one that is written for the kernels'
// shared-memory layout, everything else (the host emulation included) runs decode_stream above.
struct GenericStream {
  // pass 1 of huff_decode_fast, if this policy has a version of its own for the byte source BP (false: it has not)
  template <int STRIDE, class BP>
  MYB_HD bool parse_table(BP, int, const DecScratch<STRIDE>&, int&, bool&, int&, uint32_t&, uint32_t&) const { return false; }
  template <int PAIRS, int STRIDE, class BP, class Emit, class W>
  MYB_HD void run(DecStream& st, const uint32_t (&kk)[4], int maxlen, BP data, int data_bytes, const DecScratch<STRIDE>& D,
                  Emit& emit, const W& warp) const {
    decode_stream<PAIRS>(st, kk, maxlen, data, data_bytes, D, emit, warp);
  }
};

// Returns 0 (ok), 1 (error: the conditions huff_decode_block reports) or 2 (not handled here, nothing emitted).
// BP: where the chunk's bytes come from -- a plain pointer, or anything with operator[] and operator+ (the decoder kernel
// reads chunks that lie in its shared-memory staging area through 32-bit shared addresses).
template <int STRIDE, class BP, class Emit, class W, class S = GenericStream>
MYB_HD int huff_decode_fast(BP chunk, int size, const DecScratch<STRIDE>& D, Emit&& emit, int* n_emitted, const W& warp,
                            const S& stream = S{}) {
  int err = 0;
  int bits = 0, table_bytes = 0;
  if (size >= 3) {
    bits = chunk[0] | (chunk[1] << 8);
    table_bytes = chunk[2];
    if (bits > 512 || 3 + table_bytes + ((bits + 7) >> 3) > size) err = 1;
  } else if (size > 0) {
    err = 1;
  }
  if (err || size == 0) { bits = 0; table_bytes = 0; }
  const BP groups = chunk + 3;
  // ---- pass 1: one table symbol per step (Huffman.cpp:258-266, :54-69)
  bool general = false;
  uint32_t cnt_lo = 0, cnt_hi = 0;  // symbols per length, one byte each (lengths 1..4, 5..8)
  int n = 0;
  if (!stream.parse_table(groups, table_bytes, D, err, general, n, cnt_lo, cnt_hi)) {
    int gi = 0, ci = 0, cnt = 0, glen = 0, symbase = 0;
    while (!err && !general && (ci < cnt || gi < table_bytes)) {
      {
        if (ci == cnt) {  // next group
          const int info = groups[gi];
          const int len = (info >> 5) + 1, c = (info & 31) + 1;
          if (len <= glen || n + c > kDecFastSyms) {
            general = true;
          } else {
            glen = len;
            cnt = c;
            ci = 0;
            symbase = gi + 1;
            gi += 1 + ((c * 11 + 7) >> 3);
            if (gi > table_bytes) err = 1;
            if (len <= 4) cnt_lo += (uint32_t)c << (8 * (len - 1));
            else cnt_hi += (uint32_t)c << (8 * (len - 5));
          }
        }
        if (!err && !general) {
          const int bit = ci * 11, byte = symbase + (bit >> 3);
          uint32_t t = groups[byte] | ((uint32_t)groups[byte + 1] << 8);
          if ((bit & 7) > 5) t |= (uint32_t)groups[byte + 2] << 16;
          t = (t >> (bit & 7)) & 0x7ffu;
          D.sym(n) = (int16_t)((t >= 1024u) ? (int)t - 2048 : (int)t);
          n++;
          ci++;
        }
      }
    }
  }
  // ---- pass 2: code ranges.  k[l] = 256 - lim[l] with lim[l] = (first_l + count_l) << (7 - l) the left-aligned end of the
  // codes of length l + 1, so that  (r + k[l]) >> 8 == (r >= lim[l])  for an 8-bit r; two lengths share a register.
  // Lengths past the longest one get k = 0 (never passed), so an unassigned code shows up as length maxlen + 1.
  uint32_t kk[4] = {0, 0, 0, 0};
  int maxlen = 0;
  {
    uint32_t first = 0, off = 0;
    auto length_step = [&](int l, uint32_t c) {
      D.bs(l) = (int16_t)((int)off - (int)first);
      const uint32_t end = first + c;
      if (end > (2u << l)) general = true;  // more codes of this length than exist: let the general decoder reproduce the reference
      if (c) maxlen = l + 1;
      kk[l >> 1] |= (256u - ((end << (7 - l)) & 0x1ffu)) << (16 * (l & 1));
      first = end << 1;
      off += c;
    };
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int l = 0; l < 4; l++) length_step(l, (cnt_lo >> (8 * l)) & 0xffu);
    if (warp.umax(cnt_hi != 0u ? 1 : 0)) {  // codes of five and more bits somewhere in the warp (rare at q 50)
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int l = 4; l < 8; l++) length_step(l, (cnt_hi >> (8 * (l - 4))) & 0xffu);
    }
    // lengths past the longest one: k = 0
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int w = 0; w < 4; w++) {
      const int keep = maxlen - 2 * w;
      kk[w] = keep >= 2 ? kk[w] : (keep == 1 ? kk[w] & 0xffffu : 0u);
    }
  }
  if (err || general) bits = 0;  // such lanes idle through the lockstep loop below (no early return: the warp stays converged)
  const int maxlw = warp.umax(maxlen);  // the whole warp runs one instantiation of the stream loop
  const BP data = groups + table_bytes;
  const int data_bytes = (bits + 7) >> 3;
  // ---- code stream (Huffman.cpp:106-154); the loop exists three times, for code tables of up to 2, 4 and 8 lengths
  DecStream st;
  st.sh = 0;
  st.rem = bits;
  st.j = 0;
  st.err = err;
  st.byte0 = 0;
  st.rwin = bit_reverse32(load_window(data, 0, data_bytes));
  if (maxlw <= 2) stream.template run<1>(st, kk, maxlen, data, data_bytes, D, emit, warp);
  else if (maxlw <= 4) stream.template run<2>(st, kk, maxlen, data, data_bytes, D, emit, warp);
  else stream.template run<4>(st, kk, maxlen, data, data_bytes, D, emit, warp);
  *n_emitted = st.j;
  return st.err ? 1 : (general ? 2 : 0);
}

}  // namespace myyuvb
