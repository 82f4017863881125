// tests/hostemu/hostemu.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's per-block entropy coder (yuv-manipulations-2_b200/csrc/block_codec.cuh, the exact
// source the CUDA kernels inline) for the host with g++, so its logic -- libstdc++ tie-break emulation,
// canonical codes, serialisation, decoder -- can be checked against the oracle without a GPU.
// It also checks the two arithmetic identities the kernels rely on (exact division by one Newton step,
// round-half-away via a round-toward-zero add).  Nothing in the product links this file.
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>

#include "../../yuv-manipulations-2_b200/csrc/block_codec.cuh"

using namespace myyuvb;

namespace {
const uint8_t kZigzag[64] = {MYB_ZIGZAG_LIST};

struct ZArray {  // same views as the kernel's ZShared
  int16_t* z;
  int get(int i) const { return ((int)((uint32_t)(uint16_t)z[i] << 21)) >> 21; }
  void set(int i, int v) { z[i] = (int16_t)v; }
  uint32_t raw(int i) const { return (uint16_t)z[i]; }
  void setraw(int i, uint32_t w) { z[i] = (int16_t)(uint16_t)w; }
  int slot(int i) const { return ((uint16_t)z[i] >> 11) & 15; }
};
}  // namespace

namespace {
// mode 0: general code on the 64-symbol scratch only; 1: 15-symbol scratch first (the kernel's pre-fast8 flow);
// 2: the kernel's flow -- hash histogram, then the 15-symbol fast path or the general code on the 64-symbol scratch;
// 3: as 2, with the general code as heavy_blocks_kernel runs it (split accessors, 32-symbol scratch, then 64);
// 4, 5: the fast-path instantiations as the two builds of the coding kernel and heavy15_kernel call them, with a warp maximum
//       above the lane's own symbol count (see below)
template <int STRIDE>
int encode_blocks(const int16_t* coef, uint32_t n, int mode, uint8_t* out, uint8_t* sizes) {
  using Fast = HuffScratch<15, STRIDE>;
  using Big = HuffScratch<64, STRIDE>;
  uint8_t* fb = new uint8_t[(size_t)Fast::kBytes * STRIDE]();
  int16_t* fh = new int16_t[(size_t)Fast::kSyms * STRIDE]();
  uint8_t* bb = new uint8_t[(size_t)Big::kBytes * STRIDE]();
  int16_t* bh = new int16_t[(size_t)Big::kSyms * STRIDE]();
  uint32_t* f8sc = new uint32_t[(size_t)16 * STRIDE]();
  uint32_t* f8aux = new uint32_t[(size_t)16 * STRIDE]();
  int big_used = 0;
  for (uint32_t b = 0; b < n; b++) {
    int16_t z[64];
    for (int i = 0; i < 64; i++) z[i] = coef[64 * (size_t)b + kZigzag[i]];
    int L = 64;
    while (L > 0 && z[L - 1] == 0) L--;
    ZArray za{z};
    Fast fs{fb, fh};
    Big bs{bb, bh};
    uint8_t tmp[256];
    int sz;
    bool done = false;
    if (mode == 2 || mode == 3) {
      const int lane = STRIDE > 1 ? (int)(b % STRIDE) : 0;  // exercise the lane interleave
      FastScratch<STRIDE> F{f8sc + lane, reinterpret_cast<uint8_t*>(f8aux), lane};
      for (int i = 0; i < 32; i++) F.tab(i) = 0xffff;
      // the tile pass counts up to 8 symbols; blocks with more go to heavy15_kernel, which starts over on the coefficient
      // words (they may carry the first attempt's slot bits) with a histogram of up to 15
      int ns = huff_hist<8>(za, L, true, F, NoWarp{});
      if (ns < 0) {
        for (int i = 0; i < 64; i++) za.setraw(i, (uint32_t)(((int32_t)(za.raw(i) << 21)) >> 21));  // 11-bit value, sign extended
        for (int i = 0; i < 32; i++) F.tab(i) = 0xffff;
        ns = huff_hist<kFastCap>(za, L, true, F, NoWarp{});
      }
      if (ns >= 0) {
        const FastPlan pl = huff_fast_plan(ns, L == 0 ? 1 : L, F, NoWarp{});
        huff_fast_emit(za, pl, F, tmp, NoWarp{});
        sz = pl.size();
        done = true;
      }
    }
    if (mode == 4 || mode == 5) {
      // The instantiations exactly as the kernels call them, for a lane whose warp holds blocks with MORE symbols than its own
      // (nw = the warp's maximum > n: the loops then run to nw with this lane's part predicated):
      //   mode 4  dct_compress_kernel<queue>: histogram to 8, huff_fast_plan_n<8> / huff_fast_emit_n<8> with nw in n..8; blocks
      //           with more go the way of heavy15_kernel: histogram to 15, the UNROLLED 15-symbol plan, nw in n..15
      //   mode 5  dct_compress_kernel<in place>: histogram to 15, the rolled 15-symbol plan with nw in max(n, 9)..15 also for
      //           lanes with 8 symbols and fewer
      const int lane = STRIDE > 1 ? (int)(b % STRIDE) : 0;
      FastScratch<STRIDE> F{f8sc + lane, reinterpret_cast<uint8_t*>(f8aux), lane};
      for (int i = 0; i < 32; i++) F.tab(i) = 0xffff;
      int ns = mode == 4 ? huff_hist<8>(za, L, true, F, NoWarp{}) : huff_hist<kFastCap>(za, L, true, F, NoWarp{});
      const int ml = L == 0 ? 1 : L;
      if (mode == 4 && ns >= 0) {
        const int nw = ns + (int)((b * 5u) % (uint32_t)(9 - ns));  // n..8
        const FastPlan pl = huff_fast_plan_n<8>(ns, nw, ml, F, NoWarp{});
        huff_fast_emit_n<8>(za, pl, nw, F, tmp, NoWarp{});
        sz = pl.size();
        done = true;
      } else {
        if (mode == 4) {
          for (int i = 0; i < 64; i++) za.setraw(i, (uint32_t)(((int32_t)(za.raw(i) << 21)) >> 21));
          for (int i = 0; i < 32; i++) F.tab(i) = 0xffff;
          ns = huff_hist<kFastCap>(za, L, true, F, NoWarp{});
        }
        if (ns >= 0) {
          const int lo = ns < 9 ? 9 : ns;
          const int nw = lo + (int)((b * 3u) % (uint32_t)(kFastCap + 1 - lo));  // max(n, 9)..15
          const FastPlan pl = mode == 4 ? huff_fast_plan_n<kFastCap, STRIDE, NoWarp, false>(ns, nw, ml, F, NoWarp{})
                                        : huff_fast_plan_n<kFastCap, STRIDE, NoWarp, true>(ns, nw, ml, F, NoWarp{});
          huff_fast_emit_n<kFastCap>(za, pl, nw, F, tmp, NoWarp{});
          sz = pl.size();
          done = true;
        }
      }
      if (!done)  // more than 15 symbols: the general code starts over from the coefficient values
        for (int i = 0; i < 64; i++) za.setraw(i, (uint32_t)(((int32_t)(za.raw(i) << 21)) >> 21));
    }
    if (!done && mode == 3) {
      // heavy_blocks_kernel's flow: coefficients read in place, slot numbers in a byte column; 32-symbol scratch, then 64
      using Mid = HuffScratch<32, STRIDE>;
      static uint8_t* mb = new uint8_t[(size_t)Mid::kBytes * STRIDE]();
      static int16_t* mh = new int16_t[(size_t)Mid::kSyms * STRIDE]();
      static uint8_t* slots = new uint8_t[(size_t)64 * STRIDE]();
      uint16_t words[64];
      for (int i = 0; i < 64; i++) words[i] = (uint16_t)za.raw(i);  // may carry huff_hist's slot bits 11..14
      const int lane = STRIDE > 1 ? (int)(b % STRIDE) : 0;
      Mid ms{mb + lane, mh + lane};
      Big bl{bb + lane, bh + lane};
      ZSplitValues<STRIDE> zv{words, slots + lane};
      ZSplitSlots<STRIDE> zs{slots + lane};
      HuffPlan pl = huff_plan(zv, L, ms, NoWarp{});
      if (pl.n >= 0) {
        huff_emit(zs, pl, ms, tmp, NoWarp{});
      } else {
        big_used++;
        pl = huff_plan(zv, L, bl, NoWarp{});
        huff_emit(zs, pl, bl, tmp, NoWarp{});
      }
      sz = pl.size();
      done = true;
    }
    if (!done) {
      HuffPlan pl;
      pl.n = -1;
      if (mode == 1) pl = huff_plan(za, L, fs, NoWarp{});
      bool big = false;
      if (pl.n < 0) {
        big = true;
        big_used++;
        pl = huff_plan(za, L, bs, NoWarp{});
      }
      if (big) huff_emit(za, pl, bs, tmp, NoWarp{});
      else huff_emit(za, pl, fs, tmp, NoWarp{});
      sz = pl.size();
    }
    memcpy(out, tmp, (size_t)sz);
    out += sz;
    sizes[b] = (uint8_t)sz;
  }
  delete[] fb; delete[] fh; delete[] bb; delete[] bh; delete[] f8sc; delete[] f8aux;
  return big_used;
}
}  // namespace

extern "C" {

// coef: n x 64 row-major.  stride: element stride of the scratch arrays (1, or 128 to mimic the shared-memory
// interleave of the kernel).  mode: see encode_blocks.
int hostemu_encode_blocks(const int16_t* coef, uint32_t n, int stride, int mode, uint8_t* out, uint8_t* sizes) {
  return stride == 128 ? encode_blocks<128>(coef, n, mode, out, sizes) : stride == 64 ? encode_blocks<64>(coef, n, mode, out, sizes)
       : stride == 32 ? encode_blocks<32>(coef, n, mode, out, sizes) : encode_blocks<1>(coef, n, mode, out, sizes);
}

// mode 0: the general decoder; mode 1: the kernel's flow (fast decoder, general decoder when it declines).
// Returns 0, or 1 + index of the first block with an error; *fast_used counts blocks the fast decoder handled.
int hostemu_decode_blocks2(const uint8_t* chunks, const uint8_t* sizes, uint32_t n, int mode, int16_t* coef, uint32_t* fast_used) {
  int16_t symtab[32], base[8];
  DecScratch<1> D{symtab, base};
  uint32_t used = 0;
  for (uint32_t b = 0; b < n; b++) {
    int16_t* c = coef + 64 * (size_t)b;
    memset(c, 0, 128);
    auto emit = [&](int j, int v) { c[kZigzag[j]] = (int16_t)v; };
    int err = 2;
    int ne = 0;
    if (mode == 1) err = huff_decode_fast(chunks, sizes[b], D, emit, &ne, NoWarp{});
    if (err == 2) err = huff_decode_block(chunks, sizes[b], emit, NoWarp{});
    else used++;
    if (err) return (int)b + 1;
    chunks += sizes[b];
  }
  if (fast_used) *fast_used = used;
  return 0;
}

int hostemu_decode_blocks(const uint8_t* chunks, const uint8_t* sizes, uint32_t n, int16_t* coef) {
  return hostemu_decode_blocks2(chunks, sizes, n, 0, coef, nullptr);
}

// kernels.cu fdct_quant_pair: q1 = fma(fma(-q, y*r, y), r, y*r) with r = RN(1/q) must equal RN(y/q)
// for every integer divisor 1..255.  Returns the number of mismatches over `samples` random y per divisor
// plus structured y around rounding ties (k + 0.5) * q.
uint64_t hostemu_division_check(uint32_t samples, uint32_t seed) {
  std::mt19937 rng(seed);
  uint64_t bad = 0;
  for (int q = 1; q <= 255; q++) {
    const float d = (float)q, r = 1.0f / d;
    auto test = [&](float y) {
      const volatile float q0 = y * r;
      const float rem = fmaf(-d, q0, y);
      const float q1 = fmaf(rem, r, q0);
      const volatile float ref = y / d;
      if (!(q1 == ref)) bad++;
    };
    for (uint32_t s = 0; s < samples; s++) {
      uint32_t bits = rng();
      // |y| < 8192 covers every DCT output (|Y| <= 1024 * 8); random mantissa/exponent/sign
      const int e = 87 + (int)((bits >> 23) % 53);  // 2^-40 .. 2^12
      bits = (bits & 0x807fffffu) | ((uint32_t)e << 23);
      float y;
      memcpy(&y, &bits, 4);
      test(y);
    }
    for (int k = -1100; k <= 1100; k++) {
      const float t = ((float)k + 0.5f) * d;
      test(t);
      test(nextafterf(t, 1e9f));
      test(nextafterf(t, -1e9f));
    }
  }
  return bad;
}

// round half away from zero == trunc(RZ(v + copysign(0.5, v)))
uint64_t hostemu_round_check(uint32_t samples, uint32_t seed) {
  std::mt19937 rng(seed);
  uint64_t bad = 0;
  auto test = [&](float v) {
    const float ref = roundf(v);
    fesetround(FE_TOWARDZERO);
    const volatile float t = v + copysignf(0.5f, v);
    fesetround(FE_TONEAREST);
    if ((int)truncf(t) != (int)ref) bad++;
  };
  for (uint32_t s = 0; s < samples; s++) {
    uint32_t bits = rng();
    const int e = 100 + (int)((bits >> 23) % 40);
    bits = (bits & 0x807fffffu) | ((uint32_t)e << 23);
    float v;
    memcpy(&v, &bits, 4);
    test(v);
  }
  for (int k = -2000; k <= 2000; k++) {
    const float t = (float)k + 0.5f;
    test(t);
    test(nextafterf(t, 1e9f));
    test(nextafterf(t, -1e9f));
    test((float)k);
  }
  test(0.49999997f);
  test(-0.49999997f);
  test(0.0f);
  test(-0.0f);
  return bad;
}

}  // extern "C"
