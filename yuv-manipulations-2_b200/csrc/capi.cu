// capi.cu -- the C ABI declared in include/myyuvb200.h: contexts, argument checks that mirror the
// reference's exceptions, quantisation tables, device/pinned buffer management, H2D/D2H, kernel launches.
// There is NO CPU implementation of the codec in this library: every entry point either runs the CUDA
// kernels of kernels.cu or fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>

#include "../../include/myyuvb200.h"
#include "kernels.h"

using namespace myyuvb;

namespace {

thread_local std::string g_err;

// MYYUVB_TRACE=1: where a call spends its host time (profiles/first_call.py, profiles/cli_timing.py)
struct TraceScope {
  const char* what;
  std::chrono::steady_clock::time_point t0;
  bool on;
  explicit TraceScope(const char* w) : what(w), t0(std::chrono::steady_clock::now()) {
    static const bool trace = getenv("MYYUVB_TRACE") != nullptr;
    on = trace;
  }
  ~TraceScope() {
    if (on) fprintf(stderr, "[myyuvb] %s: %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess)                                                                                    \
      return fail(MYYUVB_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call);      \
  } while (0)

// JPEG Annex K quantisation tables at quality 50 (the reference's lum_q_table / chroma_q_table, DCT.cpp:199-219)
const uint8_t kLumaQ50[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                              14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                              18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                              49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kChromaQ50[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// DCT.cpp:286-290: mul = q >= 50.5 ? (100-q)/50 : 50/q ; table = clamp(round(q50 * mul), 1, 255), all in float.
// Host code is compiled without FMA contraction (build.py: -ffp-contract=off), each step is one float op.
void make_qtables(const uint8_t quality[3], QTables* qt) {
  for (int p = 0; p < 3; p++) {
    const volatile float qf = (float)quality[p];
    const volatile float mul = (qf >= 50.5f) ? (100.0f - qf) / 50.0f : 50.0f / qf;
    const uint8_t* base = p == 0 ? kLumaQ50 : kChromaQ50;
    for (int i = 0; i < 64; i++) {
      const volatile float prod = (float)base[i] * mul;
      float v = roundf(prod);
      v = v < 1.0f ? 1.0f : (v > 255.0f ? 255.0f : v);
      qt->q[p][i] = v;
    }
    for (int bp = 0; bp < 4; bp++)
      for (int a = 0; a < 8; a++) {
        const float q0 = qt->q[p][a * 8 + 2 * bp], q1 = qt->q[p][a * 8 + 2 * bp + 1];
        // correctly rounded reciprocals, used by the exact-division step in the kernel
        qt->rq[p][bp * 8 + a] = {1.0f / q0, 1.0f / q1, -q0, -q1};
      }
  }
}

struct Buffer {
  void* p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  int reserve(size_t bytes) {
    if (bytes <= cap) return MYYUVB_OK;
    release();
    const size_t want = bytes + bytes / 8 + 256;
    static const bool trace = getenv("MYYUVB_TRACE") != nullptr;  // where a first call spends its time (profiles/first_call.py)
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = pinned ? cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable) : cudaMalloc(&p, want);
    if (trace)
      fprintf(stderr, "[myyuvb] %s %zu bytes: %.3f ms\n", pinned ? "cudaHostAlloc" : "cudaMalloc", want,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(MYYUVB_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " allocating " + std::to_string(want) + " bytes");
    }
    cap = want;
    return MYYUVB_OK;
  }
  void release() {
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

int check_quality(const uint8_t q[3]) {
  for (int i = 0; i < 3; i++)
    if (q[i] < 1 || q[i] > 100) return fail(MYYUVB_ERR_QUALITY, "Level of quality must be between 1 and 100");
  return MYYUVB_OK;
}

// applyDCTPlane / restoreDCTPlane throw on the first plane whose width or height is not a multiple of 8
// (DCT.cpp:280-285, :338-343); planes are Y (w x h) then U, V (w/2 x h/2).
int check_dims(uint32_t w, uint32_t h) {
  if (w == 0 || h == 0) return fail(MYYUVB_ERR_ARG, "Error. empty image");
  if (w % 8 != 0) return fail(MYYUVB_ERR_WIDTH, "Error. width % 8 must be 0");
  if (h % 8 != 0) return fail(MYYUVB_ERR_HEIGHT, "Error. height % 8 must be 0");
  if ((w / 2) % 8 != 0) return fail(MYYUVB_ERR_WIDTH, "Error. width % 8 must be 0");
  if ((h / 2) % 8 != 0) return fail(MYYUVB_ERR_HEIGHT, "Error. height % 8 must be 0");
  if ((uint64_t)w * h * 3 / 2 > 0xffffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "Error. image does not fit the format's uint32 sizes");
  return MYYUVB_OK;
}

}  // namespace

struct myyuvb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t d2h_ev[2] = {nullptr, nullptr};  // per output slot: the download of the chunk that last used it
  bool own_stream = false;
  int grid = 0, grid_dec = 0;
  Buffer d_in, d_out, d_plane_start, d_counters, d_sizes, d_overflow, d_desc, d_offsets;
  Buffer d_scratch, d_tile_pos, d_tile_total, d_tile_prefix;
  Buffer d_heavy_rec, d_heavy_coef, d_heavy_bytes, d_block_slot, d_heavy_list, d_heavy_list2;
  Buffer h_small, h_stage_in, h_stage_out, h_ring;
  cudaEvent_t ring_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // one per slot of h_ring (pageable <-> device staging)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t kev[2] = {nullptr, nullptr};  // timing events around the last main codec kernel
  uint64_t pending_payload = 0;             // bytes of the payload myyuvb_dct_compress_begin left in d_out
  int enc_mode = 0;                         // 0: choose the coding kernel's build from the last launch's queue share, 1: queue, 2: in place
  int enc_in_place = 0;                     // the current choice in mode 0
  myyuvb_ctx() { h_small.pinned = h_stage_in.pinned = h_stage_out.pinned = h_ring.pinned = true; }
};

namespace {

int ensure_workspace(myyuvb_ctx* c, const FrameGeom& g, bool encoder, Workspace* ws, uint64_t out_capacity = 0) {
  const uint64_t tiles = (uint64_t)g.tiles_per_frame * g.n_frames;
  int rc;
  if ((rc = c->d_plane_start.reserve(((uint64_t)g.n_frames * 4 + 2) * 8))) return rc;  // plane starts, then frame bases
  if (!c->d_counters.p) {
    if ((rc = c->d_counters.reserve(64))) return rc;
    CU(cudaMemsetAsync(c->d_counters.p, 0, 64, c->stream));
  }
  if (encoder) {
    if ((rc = c->d_sizes.reserve((uint64_t)g.nblk_frame * g.n_frames + 16))) return rc;
    if ((rc = c->d_overflow.reserve((uint64_t)c->grid * kEncTile * 256))) return rc;
    // pass-1 parking area: never more than the payload itself, i.e. never more than the caller's capacity
    const uint64_t worst = (uint64_t)g.nblk_frame * g.n_frames * 255;
    ws->scratch_cap = std::min<uint64_t>(out_capacity, worst);
    if ((rc = c->d_scratch.reserve(ws->scratch_cap + 16))) return rc;
    if ((rc = c->d_tile_pos.reserve(tiles * 8))) return rc;
    if ((rc = c->d_tile_total.reserve(tiles * 4))) return rc;
    if ((rc = c->d_tile_prefix.reserve(tiles * 8))) return rc;
    // queue of blocks with more than 8 distinct symbols: room for every second block (beyond that they are coded in place);
    // natural content queues 10 % of its blocks at q 50 and 29 % at q 90, 404 bytes of device memory per slot
    const uint64_t nblk_total = (uint64_t)g.nblk_frame * g.n_frames;
    ws->heavy_cap = nblk_total < 0xfffffff0ull ? (uint32_t)std::max<uint64_t>(1024, nblk_total / 2) : 0u;
    if (ws->heavy_cap) {
      if ((rc = c->d_heavy_rec.reserve((uint64_t)ws->heavy_cap * 16))) return rc;
      if ((rc = c->d_heavy_coef.reserve((uint64_t)ws->heavy_cap * 128))) return rc;
      if ((rc = c->d_heavy_bytes.reserve((uint64_t)ws->heavy_cap * 256))) return rc;
      if ((rc = c->d_block_slot.reserve(nblk_total * 4))) return rc;
      if ((rc = c->d_heavy_list.reserve((uint64_t)ws->heavy_cap * 4))) return rc;
      if ((rc = c->d_heavy_list2.reserve((uint64_t)ws->heavy_cap * 4))) return rc;
    }
  } else {
    if ((rc = c->d_desc.reserve((uint64_t)g.n_frames * 3 * sizeof(PlaneDesc)))) return rc;
    if ((rc = c->d_tile_total.reserve(tiles * 4))) return rc;
    if ((rc = c->d_tile_prefix.reserve(tiles * 8))) return rc;
  }
  ws->plane_start = c->d_plane_start.as<uint64_t>();
  ws->frame_base = c->d_plane_start.as<uint64_t>() + ((uint64_t)g.n_frames * 3 + 1);
  ws->counters = c->d_counters.as<uint32_t>();
  ws->chunk_sizes = c->d_sizes.as<uint8_t>();
  ws->overflow = c->d_overflow.as<uint8_t>();
  ws->scratch = c->d_scratch.as<uint8_t>();
  ws->tile_pos = c->d_tile_pos.as<uint64_t>();
  ws->tile_total = c->d_tile_total.as<uint32_t>();
  ws->tile_prefix = c->d_tile_prefix.as<uint64_t>();
  ws->heavy_rec = c->d_heavy_rec.as<uint4>();
  ws->heavy_coef = c->d_heavy_coef.as<uint16_t>();
  ws->heavy_bytes = c->d_heavy_bytes.as<uint8_t>();
  ws->block_slot = c->d_block_slot.as<uint32_t>();
  ws->heavy_list = c->d_heavy_list.as<uint32_t>();
  ws->heavy_list2 = c->d_heavy_list2.as<uint32_t>();
  if (!encoder) ws->heavy_cap = 0;
  ws->plane_desc = c->d_desc.p;
  ws->code_in_place = 0;
  ws->queue_stats = nullptr;
  if (encoder) {
    // Which build of the coding kernel: blocks with more than 8 symbols are queued unless more than 30 % of the blocks of the
    // previous launch on this context were of that kind (content that is detailed throughout: queueing would only add
    // traffic); below 20 % it goes back.  Every launch reports its count into mapped host memory, read here without any
    // wait -- a stale value only delays the switch by a launch, and the bytes produced are the same either way.
    if ((rc = c->h_small.reserve(256 + 64))) return rc;
    volatile uint32_t* st = c->h_small.as<uint32_t>() + 32;  // bytes 128..135 of the small pinned block
    if (c->enc_mode == 0) {
      const uint32_t queued = st[0], blocks = st[1];  // blocks with more than 8 symbols, all blocks
      if (blocks) {
        if ((uint64_t)queued * 10 > (uint64_t)blocks * 3) c->enc_in_place = 1;
        else if ((uint64_t)queued * 10 < (uint64_t)blocks * 2) c->enc_in_place = 0;
      }
    } else {
      c->enc_in_place = c->enc_mode == 2;
    }
    ws->code_in_place = c->enc_in_place;
    uint32_t* st_dev = nullptr;
    CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&st_dev), c->h_small.as<uint32_t>() + 32, 0));
    ws->queue_stats = st_dev;
  }
  ws->grid = encoder ? c->grid : c->grid_dec;
  ws->k_begin = c->kev[0];
  ws->k_end = c->kev[1];
  return MYYUVB_OK;
}

int ensure_copy_streams(myyuvb_ctx* c) {
  if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (!c->d2h_stream) CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  return MYYUVB_OK;
}

int flags_to_error(uint32_t flags) {
  // a capacity overflow leaves an incomplete payload behind, which a following decompress then rejects: report the cause
  if (flags & kFlagCapacity) return fail(MYYUVB_ERR_CAPACITY, "Error. output buffer is too small for the compressed data");
  if (flags & kFlagDctYuvSize) return fail(MYYUVB_ERR_DCTYUV_SIZE, "DCTYUV load bad size");
  if (flags & kFlagPlaneSize) return fail(MYYUVB_ERR_PLANE_SIZE, "DCTYUVPlane load bad size");
  if (flags & kFlagHuffman) return fail(MYYUVB_ERR_HUFFMAN, "Huffman bad code");
  if (flags & kFlagShardTimeout) return fail(MYYUVB_ERR_SHARD_TIMEOUT, "shard: a rank of the group did not arrive within 2 s");
  if (flags & kFlagBounds) return fail(MYYUVB_ERR_BOUNDS, "Image coordinates are out of bounds");
  return MYYUVB_OK;
}

// reads and clears the device error flags (synchronises the stream)
int read_flags(myyuvb_ctx* c) {
  int rc;
  if ((rc = c->h_small.reserve(256))) return rc;
  volatile uint32_t* h = c->h_small.as<uint32_t>();
  uint32_t* h_dev = nullptr;
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_dev), c->h_small.p, 0));
  launch_publish_words(h_dev, c->d_counters.as<uint32_t>() + 1, 1, c->stream);  // no copy engine: see launch_publish_words
  CU(cudaMemsetAsync(c->d_counters.as<uint32_t>() + 1, 0, 4, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return flags_to_error(*h);
}

bool is_pinned_or_device(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Host <-> device copies of PAGEABLE host memory (what the class API hands over: YUV::data is new uint8_t[]).
// Measured on the same box (profiles/host_staging_ab.py): for uploads the driver's own pageable path beats a hand-made
// ring of pinned slots (one 4K frame: 2.1 vs 2.3 ms per compress call), so uploads go straight to cudaMemcpyAsync;
// for downloads of 32 MB and more a ring of four 2 MB pinned slots (DMA of slice k+1 in flight during the CPU copy of
// slice k) wins (one 8K frame: 10 vs 12 ms per decompress call).  MYYUVB_STAGING=direct switches the ring off.
constexpr size_t kRingSlot = 2u << 20;
constexpr bool kDefaultOwnD2H = true;
constexpr long kDefaultChunkMB = 32;
bool ring_enabled() {
  static const bool on = [] { const char* e = getenv("MYYUVB_STAGING"); return !(e && strcmp(e, "direct") == 0); }();
  return on;
}

int copy_async(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
  if (bytes == 0) return MYYUVB_OK;
  CU(cudaMemcpyAsync(dst, src, bytes, kind, s));  // issuing large copies in 4-16 MB pieces was measured: no gain
  return MYYUVB_OK;
}

// batch_host calls: bytes of IYUV per pipeline chunk (MYYUVB_CHUNK_MB)
uint64_t host_chunk_bytes() {
  static const uint64_t b = [] {
    const char* e = getenv("MYYUVB_CHUNK_MB");
    const long mb = e ? atol(e) : kDefaultChunkMB;
    return (uint64_t)(mb > 0 ? mb : kDefaultChunkMB) << 20;
  }();
  return b;
}

// batch_host calls: downloads on their own stream (MYYUVB_D2H_STREAM=0: on the kernel stream)
bool own_d2h_stream() {
  static const bool on = [] { const char* e = getenv("MYYUVB_D2H_STREAM"); return e ? e[0] == '1' : kDefaultOwnD2H; }();
  return on;
}

// The small side of a batch_host call (payloads, offsets): moved by sm_copy_kernel when the caller's buffer is mapped
// pinned memory, so that it does not queue behind another context's large transfers on the copy engine (a high-priority
// stream did not help, profiles/r01_notes.md).  MYYUVB_SMALL_COPY=0: always the copy engine.
bool sm_copy_enabled() {
  static const bool on = [] { const char* e = getenv("MYYUVB_SMALL_COPY"); return !(e && e[0] == '0'); }();
  return on;
}

// device-side address of h if it lies in mapped pinned host memory, else nullptr
uint8_t* mapped_devptr(const void* h) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, h) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? static_cast<uint8_t*>(a.devicePointer) : nullptr;
}

int small_upload(myyuvb_ctx* c, void* d_dst, const void* h_src, size_t bytes, cudaStream_t s) {
  if (sm_copy_enabled())
    if (const uint8_t* dp = mapped_devptr(h_src)) {
      launch_sm_copy(static_cast<uint8_t*>(d_dst), dp, bytes, s);
      CU(cudaGetLastError());
      return MYYUVB_OK;
    }
  if (bytes == 0) return MYYUVB_OK;
  return copy_async(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s);
}

int staged_download(myyuvb_ctx* c, void* h_dst, const void* d_src, size_t bytes, cudaStream_t s);
int small_download(myyuvb_ctx* c, void* h_dst, const void* d_src, size_t bytes, cudaStream_t s) {
  if (sm_copy_enabled())
    if (uint8_t* dp = mapped_devptr(h_dst)) {
      launch_sm_copy(dp, static_cast<const uint8_t*>(d_src), bytes, s);
      CU(cudaGetLastError());
      return MYYUVB_OK;
    }
  return staged_download(c, h_dst, d_src, bytes, s);
}

int staged_download(myyuvb_ctx* c, void* h_dst, const void* d_src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return MYYUVB_OK;
  TraceScope ts("staged_download (issue; ring: whole copy)");
  if (!ring_enabled() || is_pinned_or_device(h_dst) || bytes < (32u << 20))
    return copy_async(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s);
  int rc;
  if ((rc = c->h_ring.reserve(4 * kRingSlot))) return rc;
  for (auto& ev : c->ring_ev) CU(cudaEventSynchronize(ev));  // earlier users of the ring have left it
  const size_t slices = (bytes + kRingSlot - 1) / kRingSlot;
  auto issue = [&](size_t k) -> int {
    const size_t off = k * kRingSlot, n = std::min(kRingSlot, bytes - off);
    CU(cudaMemcpyAsync(c->h_ring.as<uint8_t>() + (k & 3) * kRingSlot, static_cast<const uint8_t*>(d_src) + off, n, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(c->ring_ev[k & 3], s));
    return MYYUVB_OK;
  };
  for (size_t k = 0; k < std::min<size_t>(3, slices); k++)
    if ((rc = issue(k))) return rc;
  for (size_t k = 0; k < slices; k++) {
    if (k + 3 < slices && (rc = issue(k + 3))) return rc;  // slot (k+3)&3 was drained by the CPU copy of slice k-1
    CU(cudaEventSynchronize(c->ring_ev[k & 3]));
    const size_t off = k * kRingSlot, n = std::min(kRingSlot, bytes - off);
    memcpy(static_cast<uint8_t*>(h_dst) + off, c->h_ring.as<uint8_t>() + (k & 3) * kRingSlot, n);
  }
  return MYYUVB_OK;
}

}  // namespace

extern "C" {

const char* myyuvb_last_error(void) { return g_err.c_str(); }

uint64_t myyuvb_launch_count(void) { return g_launches; }

void myyuvb_phase_clocks(uint64_t* out24, int reset) {
  static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "");
  if (out24) read_phase_clocks(reinterpret_cast<unsigned long long*>(out24), reset);
}

uint64_t myyuvb_compress_bound(uint32_t width, uint32_t height) {
  const uint64_t nblk = (uint64_t)(width / 8) * (height / 8) + 2ull * (width / 16) * (height / 16);
  return 12 + 24 + nblk + nblk * 255;
}

int myyuvb_ctx_create(int device, void* stream, myyuvb_ctx** out) {
  if (!out) return fail(MYYUVB_ERR_ARG, "myyuvb_ctx_create: null output pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(MYYUVB_ERR_CUDA, std::string("CUDA error: no usable CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
  if (device < 0 || device >= count) return fail(MYYUVB_ERR_ARG, "myyuvb_ctx_create: bad device ordinal");
  TraceScope ts("ctx_create (CUDA context, streams, events)");
  CU(cudaSetDevice(device));
  myyuvb_ctx* c = new myyuvb_ctx();
  c->device = device;
  const int rc = [&]() -> int {
    if (stream) {
      c->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
      CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      c->own_stream = true;
    }
    // copy_stream and d2h_stream are created by the first pipelined host-pointer call (ensure_copy_streams): contexts that
    // only serve device-pointer calls stay at one stream each
    for (auto& ev : c->d2h_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c->ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c->ring_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c->kev) CU(cudaEventCreate(&ev));
    return MYYUVB_OK;
  }();
  if (rc) {  // a half-built context is taken apart again (the message of the failed call survives the clean-up)
    const std::string msg = g_err;
    myyuvb_ctx_destroy(c);
    return fail(rc, msg);
  }
  c->grid = codec_grid_size(device, true);
  c->grid_dec = codec_grid_size(device, false);
  *out = c;
  return MYYUVB_OK;
}

void myyuvb_ctx_destroy(myyuvb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
  for (Buffer* b : {&c->d_in, &c->d_out, &c->d_plane_start, &c->d_counters, &c->d_sizes,
                    &c->d_overflow, &c->d_desc, &c->d_offsets, &c->d_scratch, &c->d_tile_pos, &c->d_tile_total, &c->d_tile_prefix, &c->d_heavy_rec, &c->d_heavy_coef, &c->d_heavy_bytes, &c->d_block_slot, &c->d_heavy_list, &c->d_heavy_list2, &c->h_small, &c->h_stage_in, &c->h_stage_out, &c->h_ring})
    b->release();
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ring_ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->kev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->d2h_ev)
    if (ev) cudaEventDestroy(ev);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  delete c;
}

int myyuvb_sync(myyuvb_ctx* c) {
  if (!c) return fail(MYYUVB_ERR_ARG, "null context");
  CU(cudaStreamSynchronize(c->stream));
  return MYYUVB_OK;
}

void* myyuvb_stream(myyuvb_ctx* c) { return c ? (void*)c->stream : nullptr; }

int myyuvb_set_encoder_mode(myyuvb_ctx* c, int mode) {
  if (!c || mode < 0 || mode > 2) return fail(MYYUVB_ERR_ARG, "myyuvb_set_encoder_mode: mode must be 0, 1 or 2");
  c->enc_mode = mode;
  if (mode) c->enc_in_place = mode == 2;
  return MYYUVB_OK;
}

int myyuvb_last_kernel_ms(myyuvb_ctx* c, float* ms) {
  if (!c || !ms) return fail(MYYUVB_ERR_ARG, "null argument");
  CU(cudaEventSynchronize(c->kev[1]));
  CU(cudaEventElapsedTime(ms, c->kev[0], c->kev[1]));
  return MYYUVB_OK;
}

int myyuvb_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(MYYUVB_ERR_ARG, "null output pointer");
  // mapped and portable (what unified addressing implies anyway): the batch_host calls let kernels read and write it
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable));
  return MYYUVB_OK;
}

void myyuvb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------
// device-pointer batch entry points
// ------------------------------------------------------------------------------------------------
namespace {
int convert_dev_impl(myyuvb_ctx* c, const uint8_t* d_px, uint32_t pixel_bytes, uint32_t w, uint32_t h, int bottom_up,
                     uint32_t n_frames, uint8_t* d_iyuv) {
  if (!c || !d_px || !d_iyuv) return fail(MYYUVB_ERR_ARG, "null argument");
  if (w == 0 || h == 0 || (w % 2) || (h % 2)) return fail(MYYUVB_ERR_EVEN, "Error. width and height must be even");
  if ((uint64_t)w * h * 4 > 0xffffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "Error. image does not fit the format's uint32 sizes");
  // Alignment is what the kernel that will run needs: the 8-pixel kernels (width % 8 == 0) use 128-bit (32-bit pixels) or
  // 64-bit (24-bit pixels) loads and 64-bit stores on every frame of the batch; the quad kernel for other widths reads 24-bit
  // pixels bytewise and 32-bit pixels as 64-bit pairs, and stores 16 bits at a time.
  if (w % 8 == 0) {
    if (pixel_bytes == 4 ? ((uintptr_t)d_px & 15) != 0 : (((uintptr_t)d_px & 7) != 0 || (n_frames > 1 && ((uint64_t)w * h * 3) % 8 != 0)))
      return fail(MYYUVB_ERR_ARG, pixel_bytes == 4 ? "device buffers must be 16-byte (input) / 8-byte (output) aligned"
                                                   : "24-bit input: device buffer must be 8-byte aligned and, in a batch, a frame a multiple of 8 bytes");
    if ((uintptr_t)d_iyuv & 7) return fail(MYYUVB_ERR_ARG, "device buffers must be 16-byte (input) / 8-byte (output) aligned");
  } else {
    if (pixel_bytes == 4 && ((uintptr_t)d_px & 7)) return fail(MYYUVB_ERR_ARG, "32-bit input: device buffer must be 8-byte aligned");
    if ((uintptr_t)d_iyuv & 1) return fail(MYYUVB_ERR_ARG, "device output must be 2-byte aligned");
  }
  CU(cudaSetDevice(c->device));
  launch_bgr_to_iyuv(d_px, pixel_bytes, d_iyuv, w, h, bottom_up, n_frames, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}

int convert_host_impl(myyuvb_ctx* c, const uint8_t* px, uint32_t pixel_bytes, uint32_t w, uint32_t h, int bottom_up, uint8_t* iyuv_out) {
  if (!c || !px || !iyuv_out) return fail(MYYUVB_ERR_ARG, "null argument");
  if (w == 0 || h == 0 || (w % 2) || (h % 2)) return fail(MYYUVB_ERR_EVEN, "Error. width and height must be even");
  CU(cudaSetDevice(c->device));
  const size_t in_bytes = (size_t)w * h * pixel_bytes, out_bytes = (size_t)w * h * 3 / 2;
  int rc;
  if ((rc = c->d_in.reserve(in_bytes))) return rc;
  if ((rc = c->d_out.reserve(out_bytes))) return rc;
  if ((rc = copy_async(c->d_in.p, px, in_bytes, cudaMemcpyHostToDevice, c->stream))) return rc;
  if ((rc = convert_dev_impl(c, c->d_in.as<uint8_t>(), pixel_bytes, w, h, bottom_up, 1, c->d_out.as<uint8_t>()))) return rc;
  if ((rc = staged_download(c, iyuv_out, c->d_out.p, out_bytes, c->stream))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return MYYUVB_OK;
}
}  // namespace

int myyuvb_xrgb_to_iyuv_batch_dev(myyuvb_ctx* c, const uint8_t* d_bgrx, uint32_t w, uint32_t h, int bottom_up,
                                  uint32_t n_frames, uint8_t* d_iyuv) {
  return convert_dev_impl(c, d_bgrx, 4, w, h, bottom_up, n_frames, d_iyuv);
}

int myyuvb_bgr24_to_iyuv_batch_dev(myyuvb_ctx* c, const uint8_t* d_bgr, uint32_t w, uint32_t h, int bottom_up,
                                   uint32_t n_frames, uint8_t* d_iyuv) {
  return convert_dev_impl(c, d_bgr, 3, w, h, bottom_up, n_frames, d_iyuv);
}

namespace {
// d_base: nullptr, or a device pointer to the position of the first payload inside d_out (written by an earlier launch)
int compress_dev_impl(myyuvb_ctx* c, const uint8_t* d_iyuv, uint32_t w, uint32_t h, const uint8_t quality[3], uint32_t n_frames,
                      uint8_t* d_out, uint64_t out_capacity, uint64_t* d_offsets, const uint64_t* d_base) {
  int rc;
  if ((uintptr_t)d_iyuv & 7) return fail(MYYUVB_ERR_ARG, "device input must be 8-byte aligned");
  CU(cudaSetDevice(c->device));
  const FrameGeom g = make_geom(w, h, n_frames, kEncTile);
  if ((uint64_t)g.tiles_per_frame * n_frames > 0x7fffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "batch too large");
  Workspace ws{};
  if ((rc = ensure_workspace(c, g, true, &ws, out_capacity))) return rc;
  QTables qt;
  make_qtables(quality, &qt);
  launch_compress(d_iyuv, g, qt, d_out, out_capacity, d_offsets, d_base, ws, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}
}  // namespace

int myyuvb_dct_compress_batch_dev(myyuvb_ctx* c, const uint8_t* d_iyuv, uint32_t w, uint32_t h, const uint8_t quality[3],
                                  uint32_t n_frames, uint8_t* d_out, uint64_t out_capacity, uint64_t* d_offsets) {
  if (!c || !d_iyuv || !quality || !d_out || !d_offsets || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  return compress_dev_impl(c, d_iyuv, w, h, quality, n_frames, d_out, out_capacity, d_offsets, nullptr);
}

// XRGB -> IYUV -> DCT payloads.  Frames go through in chunks: a chunk is converted, then coded while its IYUV bytes are
// still in L2 (126 MB), so the intermediate image of the reference's two-step API (YUV(bmp) then compress) costs no
// HBM read.  Fusing the conversion into the coding kernel itself was rejected: that kernel is instruction-issue bound
//  and every pixel would be converted up to three times (once per plane tile that needs it).
int myyuvb_xrgb_dct_compress_batch_dev(myyuvb_ctx* c, const uint8_t* d_bgrx, uint32_t w, uint32_t h, int bottom_up,
                                       const uint8_t quality[3], uint32_t n_frames, uint32_t chunk_frames, uint8_t* d_iyuv,
                                       uint8_t* d_out, uint64_t out_capacity, uint64_t* d_offsets) {
  if (!c || !d_bgrx || !quality || !d_out || !d_offsets || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  if ((uint64_t)w * h * 4 > 0xffffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "Error. image does not fit the format's uint32 sizes");
  CU(cudaSetDevice(c->device));
  const uint64_t frame_bytes = (uint64_t)w * h * 3 / 2;
  if (chunk_frames == 0) chunk_frames = (uint32_t)std::max<uint64_t>(1, (100ull << 20) / frame_bytes);  // ~100 MB of IYUV per chunk
  chunk_frames = std::min(chunk_frames, n_frames);
  uint8_t* ring = nullptr;
  if (!d_iyuv) {
    if ((rc = c->d_in.reserve((uint64_t)chunk_frames * frame_bytes))) return rc;
    ring = c->d_in.as<uint8_t>();
  }
  for (uint32_t f0 = 0; f0 < n_frames; f0 += chunk_frames) {
    const uint32_t nf = std::min(chunk_frames, n_frames - f0);
    uint8_t* iy = d_iyuv ? d_iyuv + (uint64_t)f0 * frame_bytes : ring;
    if ((rc = myyuvb_xrgb_to_iyuv_batch_dev(c, d_bgrx + (uint64_t)f0 * w * h * 4, w, h, bottom_up, nf, iy))) return rc;
    if ((rc = compress_dev_impl(c, iy, w, h, quality, nf, d_out, out_capacity, d_offsets + f0, f0 ? d_offsets + f0 : nullptr))) return rc;
  }
  return MYYUVB_OK;
}

int myyuvb_dct_decompress_batch_dev(myyuvb_ctx* c, const uint8_t* d_payloads, const uint64_t* d_offsets, uint32_t w,
                                    uint32_t h, const uint8_t quality[3], uint32_t n_frames, uint8_t* d_iyuv) {
  if (!c || !d_payloads || !d_offsets || !quality || !d_iyuv || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  if ((uintptr_t)d_iyuv & 7) return fail(MYYUVB_ERR_ARG, "device output must be 8-byte aligned");
  CU(cudaSetDevice(c->device));
  const FrameGeom g = make_geom(w, h, n_frames, kDecTile);
  if ((uint64_t)g.tiles_per_frame * n_frames > 0x7fffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "batch too large");
  Workspace ws{};
  if ((rc = ensure_workspace(c, g, false, &ws))) return rc;
  QTables qt;
  make_qtables(quality, &qt);
  launch_decompress(d_payloads, d_offsets, g, qt, d_iyuv, ws, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}

int myyuvb_batch_status(myyuvb_ctx* c) {
  if (!c) return fail(MYYUVB_ERR_ARG, "null context");
  CU(cudaSetDevice(c->device));
  if (!c->d_counters.p) {
    CU(cudaStreamSynchronize(c->stream));
    return MYYUVB_OK;
  }
  return read_flags(c);
}

// ------------------------------------------------------------------------------------------------
// host-pointer single image entry points
// ------------------------------------------------------------------------------------------------
int myyuvb_xrgb_to_iyuv(myyuvb_ctx* c, const uint8_t* bgrx, uint32_t w, uint32_t h, int bottom_up, uint8_t* iyuv_out) {
  return convert_host_impl(c, bgrx, 4, w, h, bottom_up, iyuv_out);
}

int myyuvb_bgr24_to_iyuv(myyuvb_ctx* c, const uint8_t* bgr, uint32_t w, uint32_t h, int bottom_up, uint8_t* iyuv_out) {
  return convert_host_impl(c, bgr, 3, w, h, bottom_up, iyuv_out);
}

// One image, two steps: the payload stays in the context's device memory until the caller, who now knows its size, hands
// over a buffer of exactly that size (YUV::data must be new uint8_t[data_size], myyuv_yuv.cpp:243-246).  The device-side
// output buffer starts at half the raw image -- several times what natural content needs at any quality below ~95 -- and
// only a capacity overflow makes the call repeat with the worst-case bound (255 bytes per block), so a first call does
// not allocate 4x the image size three times over (output slot, parking area, host bound buffer).
int myyuvb_dct_compress_begin(myyuvb_ctx* c, const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t quality[3], uint32_t* out_size) {
  if (!c || !iyuv || !quality || !out_size) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  CU(cudaSetDevice(c->device));
  TraceScope ts("dct_compress_begin");
  c->pending_payload = 0;
  const uint64_t frame_bytes = (uint64_t)w * h * 3 / 2;
  const uint64_t bound = myyuvb_compress_bound(w, h);
  if ((rc = c->d_in.reserve(frame_bytes))) return rc;
  if ((rc = c->d_offsets.reserve(16))) return rc;
  if ((rc = c->h_small.reserve(256 + 16))) return rc;
  volatile uint64_t* h_off = reinterpret_cast<uint64_t*>(c->h_small.as<uint8_t>() + 256);
  uint64_t* h_off_dev = nullptr;
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_off_dev), const_cast<uint64_t*>(h_off), 0));
  if ((rc = copy_async(c->d_in.p, iyuv, (size_t)frame_bytes, cudaMemcpyHostToDevice, c->stream))) return rc;
  uint64_t cap = std::min<uint64_t>(bound, std::max<uint64_t>(c->d_out.cap, frame_bytes / 2 + 4096));
  for (;;) {
    if ((rc = c->d_out.reserve(cap))) return rc;
    if ((rc = compress_dev_impl(c, c->d_in.as<uint8_t>(), w, h, quality, 1, c->d_out.as<uint8_t>(), cap, c->d_offsets.as<uint64_t>(), nullptr)))
      return rc;
    launch_publish_words(reinterpret_cast<uint32_t*>(h_off_dev), c->d_offsets.as<uint32_t>(), 4, c->stream);
    rc = read_flags(c);  // synchronises
    if (rc == MYYUVB_ERR_CAPACITY && cap < bound) {
      cap = bound;
      continue;
    }
    if (rc) return rc;
    break;
  }
  const uint64_t size = h_off[1] - h_off[0];
  if (size > 0xffffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "Error. image does not fit the format's uint32 sizes");
  c->pending_payload = size;
  *out_size = (uint32_t)size;
  return MYYUVB_OK;
}

int myyuvb_dct_compress_fetch(myyuvb_ctx* c, uint8_t* out, uint64_t out_capacity) {
  if (!c || !out) return fail(MYYUVB_ERR_ARG, "null argument");
  if (c->pending_payload == 0) return fail(MYYUVB_ERR_ARG, "myyuvb_dct_compress_fetch: no payload pending (call myyuvb_dct_compress_begin first)");
  if (out_capacity < c->pending_payload) return fail(MYYUVB_ERR_CAPACITY, "Error. output buffer is too small for the compressed data");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = staged_download(c, out, c->d_out.p, (size_t)c->pending_payload, c->stream))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->pending_payload = 0;
  return MYYUVB_OK;
}

int myyuvb_dct_compress(myyuvb_ctx* c, const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t quality[3], uint8_t* out,
                        uint64_t out_capacity, uint32_t* out_size) {
  if (!c || !iyuv || !quality || !out || !out_size) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = myyuvb_dct_compress_begin(c, iyuv, w, h, quality, out_size))) return rc;
  return myyuvb_dct_compress_fetch(c, out, out_capacity);
}

int myyuvb_dct_decompress(myyuvb_ctx* c, const uint8_t* payload, uint32_t payload_size, uint32_t w, uint32_t h,
                          const uint8_t quality[3], uint8_t* iyuv_out) {
  if (!c || !payload || !quality || !iyuv_out) return fail(MYYUVB_ERR_ARG, "null argument");
  // DCTYUV::load's first check (DCT.cpp:132-134) needs no device work
  if (payload_size <= 12) {
    int rc;
    if ((rc = check_quality(quality))) return rc;
    return fail(MYYUVB_ERR_DCTYUV_SIZE, "DCTYUV load bad size");
  }
  const uint64_t offsets[2] = {0, payload_size};
  return myyuvb_dct_decompress_batch_host(c, payload, offsets, w, h, quality, 1, iyuv_out);
}

// ------------------------------------------------------------------------------------------------
// host-pointer batch entry points: frames are processed in chunks; the H2D copy of chunk i+1 and the D2H
// copy of chunk i-1 run on the copy stream while chunk i is coded on the compute stream.
// ------------------------------------------------------------------------------------------------
static int compress_batch_host_impl(myyuvb_ctx* c, const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t quality[3],
                                    uint32_t n_frames, uint8_t* out, uint64_t out_capacity, uint64_t* offsets) {
  if (!c || !iyuv || !quality || !out || !offsets || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  CU(cudaSetDevice(c->device));
  if ((rc = ensure_copy_streams(c))) return rc;
  const uint64_t frame_bytes = (uint64_t)w * h * 3 / 2;
  const uint64_t bound = myyuvb_compress_bound(w, h);
  // chunk size: ~32 MB of input per chunk, at least one frame
  const uint32_t per = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_frames, host_chunk_bytes() / frame_bytes));
  if ((rc = c->d_in.reserve(2 * per * frame_bytes))) return rc;
  if ((rc = c->d_out.reserve(2 * per * bound))) return rc;
  if ((rc = c->d_offsets.reserve(2 * (per + 1) * 8))) return rc;
  if ((rc = c->h_small.reserve(256 + 2 * (per + 1) * 8))) return rc;
  uint64_t* h_off = reinterpret_cast<uint64_t*>(c->h_small.as<uint8_t>() + 256);
  uint64_t* h_off_dev = nullptr;  // the same words as the kernels see them (mapped pinned memory)
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_off_dev), h_off, 0));
  uint64_t written = 0;
  offsets[0] = 0;
  const uint32_t n_chunks = (n_frames + per - 1) / per;
  // Software pipeline over chunks: upload(k+1) overlaps code(k).  The host needs chunk k's sizes before it can issue the
  // download; a kernel stores them into mapped host memory, so that wait depends on the kernel stream only, never on a
  // copy engine that may be busy with another context's downloads (a 64 MB download ahead in the queue used to stall
  // this loop for a millisecond per chunk when a decompress call ran next to it).
  auto upload = [&](uint32_t k) -> int {
    const uint32_t f0 = k * per, nf = std::min(per, n_frames - f0), slot = k & 1;
    const uint8_t* src = iyuv + (uint64_t)f0 * frame_bytes;
    uint8_t* dst = c->d_in.as<uint8_t>() + (uint64_t)slot * per * frame_bytes;
    int urc;
    if ((urc = copy_async(dst, src, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, c->copy_stream))) return urc;
    CU(cudaEventRecord(c->ev[slot], c->copy_stream));
    return MYYUVB_OK;
  };
  cudaStream_t dl = own_d2h_stream() ? c->d2h_stream : c->stream;
  if ((rc = upload(0))) return rc;
  for (uint32_t k = 0; k < n_chunks; k++) {
    const uint32_t f0 = k * per, nf = std::min(per, n_frames - f0), slot = k & 1;
    uint8_t* d_src = c->d_in.as<uint8_t>() + (uint64_t)slot * per * frame_bytes;
    uint8_t* d_dst = c->d_out.as<uint8_t>() + (uint64_t)slot * per * bound;
    uint64_t* d_off = c->d_offsets.as<uint64_t>() + (uint64_t)slot * (per + 1);
    CU(cudaStreamWaitEvent(c->stream, c->ev[slot], 0));
    if (k >= 2) CU(cudaStreamWaitEvent(c->stream, c->d2h_ev[slot], 0));  // the payloads of chunk k-2 have left this output slot
    if ((rc = myyuvb_dct_compress_batch_dev(c, d_src, w, h, quality, nf, d_dst, (uint64_t)per * bound, d_off))) return rc;
    launch_publish_words(reinterpret_cast<uint32_t*>(h_off_dev + (uint64_t)slot * (per + 1)), reinterpret_cast<const uint32_t*>(d_off),
                         2 * (nf + 1), c->stream);
    CU(cudaEventRecord(c->ev[2 + slot], c->stream));
    if (k + 1 < n_chunks) {
      // the other input slot was last read by chunk k-1, which has been synchronised below
      if ((rc = upload(k + 1))) return rc;
    }
    CU(cudaEventSynchronize(c->ev[2 + slot]));
    const volatile uint64_t* ho = h_off + (uint64_t)slot * (per + 1);
    const uint64_t first = ho[0], bytes = ho[nf] - first;
    if (written + bytes > out_capacity)
      return fail(MYYUVB_ERR_CAPACITY, "Error. output buffer is too small for the compressed data");
    for (uint32_t i = 0; i <= nf; i++) offsets[f0 + i] = written + (ho[i] - first);
    // the next chunk's kernels use the other output slot
    if ((rc = small_download(c, out + written, d_dst + first, (size_t)bytes, dl))) return rc;
    CU(cudaEventRecord(c->d2h_ev[slot], dl));
    written += bytes;
  }
  if (dl != c->stream) CU(cudaStreamSynchronize(dl));
  return read_flags(c);
}

static int decompress_batch_host_impl(myyuvb_ctx* c, const uint8_t* payloads, const uint64_t* offsets, uint32_t w, uint32_t h,
                                      const uint8_t quality[3], uint32_t n_frames, uint8_t* iyuv_out) {
  if (!c || !payloads || !offsets || !quality || !iyuv_out || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  for (uint32_t f = 0; f < n_frames; f++)
    if (offsets[f + 1] < offsets[f]) return fail(MYYUVB_ERR_ARG, "frame offsets must be non-decreasing");
  CU(cudaSetDevice(c->device));
  TraceScope ts("decompress_batch_host");
  if ((rc = ensure_copy_streams(c))) return rc;
  const uint64_t frame_bytes = (uint64_t)w * h * 3 / 2;
  const uint32_t per = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_frames, host_chunk_bytes() / frame_bytes));
  const uint32_t n_chunks = (n_frames + per - 1) / per;
  uint64_t max_in = 0;
  for (uint32_t k = 0; k < n_chunks; k++) {
    const uint32_t f0 = k * per, nf = std::min(per, n_frames - f0);
    max_in = std::max(max_in, offsets[f0 + nf] - offsets[f0]);
  }
  if ((rc = c->d_in.reserve(2 * (max_in + 16)))) return rc;
  if ((rc = c->d_out.reserve(2 * per * frame_bytes))) return rc;
  if ((rc = c->d_offsets.reserve(2 * (per + 1) * 8))) return rc;
  if ((rc = c->h_small.reserve(256 + 2 * (per + 1) * 8))) return rc;
  uint64_t* h_off = reinterpret_cast<uint64_t*>(c->h_small.as<uint8_t>() + 256);
  const uint64_t in_slot = (max_in + 16) & ~15ull;
  cudaStream_t up = c->copy_stream;
  auto upload = [&](uint32_t k) -> int {
    const uint32_t f0 = k * per, nf = std::min(per, n_frames - f0), slot = k & 1;
    const uint64_t beg = offsets[f0], bytes = offsets[f0 + nf] - beg;
    const uint8_t* src = payloads + beg;
    uint64_t* ho = h_off + (uint64_t)slot * (per + 1);
    for (uint32_t i = 0; i <= nf; i++) ho[i] = offsets[f0 + i] - beg;
    int urc;
    if ((urc = small_upload(c, c->d_in.as<uint8_t>() + (uint64_t)slot * in_slot, src, (size_t)bytes, up))) return urc;
    if ((urc = small_upload(c, c->d_offsets.as<uint64_t>() + (uint64_t)slot * (per + 1), ho, (size_t)(nf + 1) * 8, up))) return urc;
    CU(cudaEventRecord(c->ev[slot], up));
    return MYYUVB_OK;
  };
  // Uploads run on copy_stream, kernels on the context stream, downloads on d2h_stream; events order the reuse of the two
  // input and two output slots.  Measured with a compress call running next to this one on a second context
  // (profiles/e2e_ab.py, 32 4K frames per call, ms per compress+decompress pair): everything on the copy engines 15.9;
  // downloads on their own stream 15.2; payloads and offsets moved by sm_copy_kernel as well 13.1; 32 MB chunks 12.9
  // (the box's raw two-way copy rate allows 11.5).
  cudaStream_t dl = own_d2h_stream() ? c->d2h_stream : c->stream;
  if ((rc = upload(0))) return rc;
  for (uint32_t k = 0; k < n_chunks; k++) {
    const uint32_t f0 = k * per, nf = std::min(per, n_frames - f0), slot = k & 1;
    uint8_t* d_dst = c->d_out.as<uint8_t>() + (uint64_t)slot * per * frame_bytes;
    CU(cudaStreamWaitEvent(c->stream, c->ev[slot], 0));                    // chunk k is on the device
    if (k >= 2) CU(cudaStreamWaitEvent(c->stream, c->d2h_ev[slot], 0));    // chunk k-2 has left this output slot
    if ((rc = myyuvb_dct_decompress_batch_dev(c, c->d_in.as<uint8_t>() + (uint64_t)slot * in_slot,
                                              c->d_offsets.as<uint64_t>() + (uint64_t)slot * (per + 1), w, h, quality, nf, d_dst)))
      return rc;
    CU(cudaEventRecord(c->ev[2 + slot], c->stream));
    CU(cudaStreamWaitEvent(dl, c->ev[2 + slot], 0));
    if ((rc = staged_download(c, iyuv_out + (uint64_t)f0 * frame_bytes, d_dst, (size_t)nf * frame_bytes, dl))) return rc;
    CU(cudaEventRecord(c->d2h_ev[slot], dl));
    if (k + 1 < n_chunks) {
      if (k >= 1) {
        // input slot (k+1)&1 was read by the kernels of chunk k-1; its pinned offsets staging was read by that chunk's upload
        CU(cudaEventSynchronize(c->ev[(k + 1) & 1]));
        CU(cudaStreamWaitEvent(up, c->ev[2 + ((k + 1) & 1)], 0));
      }
      if ((rc = upload(k + 1))) return rc;
    }
  }
  CU(cudaStreamSynchronize(c->d2h_stream));
  return read_flags(c);
}

// Every way out of the pipelined calls that is not success leaves copies in flight on three streams that read or write
// the caller's buffers, and possibly sticky device error flags: wait for the streams and clear the flags, so that the
// caller may free its buffers and the next call on the context starts clean.  The first error's code and text are kept.
static int drain_after_error(myyuvb_ctx* c, int rc) {
  if (!c || !c->stream) return rc;
  const std::string msg = g_err;
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
  if (c->d_counters.p) {
    cudaMemsetAsync(c->d_counters.as<uint32_t>() + 1, 0, 4, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  cudaGetLastError();
  return fail(rc, msg);
}

int myyuvb_dct_compress_batch_host(myyuvb_ctx* c, const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t quality[3],
                                   uint32_t n_frames, uint8_t* out, uint64_t out_capacity, uint64_t* offsets) {
  const int rc = compress_batch_host_impl(c, iyuv, w, h, quality, n_frames, out, out_capacity, offsets);
  return rc ? drain_after_error(c, rc) : rc;
}

int myyuvb_dct_decompress_batch_host(myyuvb_ctx* c, const uint8_t* payloads, const uint64_t* offsets, uint32_t w, uint32_t h,
                                     const uint8_t quality[3], uint32_t n_frames, uint8_t* iyuv_out) {
  const int rc = decompress_batch_host_impl(c, payloads, offsets, w, h, quality, n_frames, iyuv_out);
  return rc ? drain_after_error(c, rc) : rc;
}

// ------------------------------------------------------------------------------------------------
// One image sharded over the GPUs of a box (BASELINE configs[3]): one process (or context) per GPU codes
// a band of macroblock rows and stores it straight into the root's payload buffer over NVLink (kernels.cu, "the exchange
// step").  Buffers that several ranks touch -- every rank's control block, the root's payload / image buffers -- are plain
// device allocations shared through CUDA IPC handles; how the 64-byte handles travel between the processes is the
// caller's business (sharding.py uses torch.distributed's object collectives once, at set-up).
// ------------------------------------------------------------------------------------------------
uint64_t myyuvb_shard_ctrl_bytes(void) { return sizeof(ShardCtrl); }

int myyuvb_shard_rows(uint32_t height, uint32_t world, uint32_t* rows) {
  if (!rows || world == 0 || world > (uint32_t)kShardMaxWorld) return fail(MYYUVB_ERR_ARG, "shard: world must be 1..16");
  if (height % 16) return fail(MYYUVB_ERR_HEIGHT, "Error. height % 8 must be 0");
  const uint32_t mb = height / 16, base = mb / world, extra = mb % world;  // e.g. 270 rows over 8 ranks: 34 x 6 + 33 x 2
  uint32_t r = 0;
  for (uint32_t q = 0; q <= world; q++) {
    rows[q] = r * 16;
    r += base + (q < extra ? 1 : 0);
  }
  return MYYUVB_OK;
}

int myyuvb_ipc_alloc(myyuvb_ctx* c, uint64_t bytes, void** d_ptr, uint8_t handle_out[64]) {
  if (!c || !d_ptr || bytes == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries IPC handles as 64 bytes");
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(d_ptr, bytes));
  CU(cudaMemset(*d_ptr, 0, bytes));
  if (handle_out) {
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, *d_ptr));
    memcpy(handle_out, &h, 64);
  }
  return MYYUVB_OK;
}

int myyuvb_ipc_open(myyuvb_ctx* c, const uint8_t handle[64], void** d_ptr) {
  if (!c || !handle || !d_ptr) return fail(MYYUVB_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MYYUVB_OK;
}

int myyuvb_ipc_close(myyuvb_ctx* c, void* d_ptr) {
  if (!c || !d_ptr) return fail(MYYUVB_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaIpcCloseMemHandle(d_ptr));
  return MYYUVB_OK;
}

int myyuvb_ipc_free(myyuvb_ctx* c, void* d_ptr) {
  if (!c || !d_ptr) return fail(MYYUVB_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaFree(d_ptr));
  return MYYUVB_OK;
}

static int shard_peers(uint32_t height, uint32_t rank, uint32_t world, uint32_t root, const uint32_t* rows, void* const* ctrl, uint32_t epoch,
                       ShardPeers* S) {
  if (!rows || !ctrl || world == 0 || world > (uint32_t)kShardMaxWorld || rank >= world || root >= world)
    return fail(MYYUVB_ERR_ARG, "shard: bad rank / world / root");
  if (rows[0] != 0 || rows[world] != height) return fail(MYYUVB_ERR_ARG, "shard: the bands must cover the image");
  memset(S, 0, sizeof(*S));
  for (uint32_t q = 0; q < world; q++) {
    if (!ctrl[q]) return fail(MYYUVB_ERR_ARG, "shard: null control block");
    if (rows[q + 1] < rows[q] || rows[q] % 16 || rows[q + 1] % 16) return fail(MYYUVB_ERR_ARG, "shard: bands must be whole macroblock rows");
    S->ctrl[q] = static_cast<ShardCtrl*>(ctrl[q]);
  }
  for (uint32_t q = 0; q <= world; q++) S->row[q] = rows[q];
  S->rank = rank; S->world = world; S->root = root; S->epoch = epoch;
  return MYYUVB_OK;
}

// geometry of the band [y0, y1): an image of its own, or rows of a full frame that lies in device memory
static FrameGeom band_geom(uint32_t w, uint32_t h, uint32_t y0, uint32_t y1, bool in_full_frame, uint32_t tile_blocks) {
  FrameGeom g = make_geom(w, y1 - y0, 1, tile_blocks);
  if (in_full_frame) {
    g.plane_off[0] = (uint64_t)y0 * w;
    g.plane_off[1] = (uint64_t)w * h + (uint64_t)(y0 / 2) * (w / 2);
    g.plane_off[2] = (uint64_t)w * h * 5 / 4 + (uint64_t)(y0 / 2) * (w / 2);
  }
  return g;
}

int myyuvb_dct_compress_shard_dev(myyuvb_ctx* c, const uint8_t* d_iyuv, int iyuv_is_full_frame, uint32_t w, uint32_t h,
                                  const uint8_t quality[3], uint32_t rank, uint32_t world, uint32_t root, const uint32_t* rows,
                                  void* const* ctrl, uint8_t* root_out, uint64_t out_capacity, uint32_t epoch) {
  if (!c || !d_iyuv || !quality || !root_out) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  if (h % 16) return fail(MYYUVB_ERR_HEIGHT, "Error. height % 8 must be 0");
  ShardPeers S;
  if ((rc = shard_peers(h, rank, world, root, rows, ctrl, epoch, &S))) return rc;
  if ((uintptr_t)d_iyuv & 7) return fail(MYYUVB_ERR_ARG, "device input must be 8-byte aligned");
  CU(cudaSetDevice(c->device));
  const FrameGeom g = band_geom(w, h, rows[rank], rows[rank + 1], iyuv_is_full_frame != 0, kEncTile);
  const FrameGeom full = make_geom(w, h, 1, kEncTile);
  Workspace ws{};
  if ((rc = ensure_workspace(c, g, true, &ws, (uint64_t)g.nblk_frame * 255))) return rc;  // parking area: the band's worst case
  QTables qt;
  make_qtables(quality, &qt);
  if (rank == root) launch_shard_go(S, c->stream);
  launch_compress_shard(d_iyuv, g, full.nblk, qt, root_out, out_capacity, S, ws, c->stream);
  launch_shard_done(S, ws, 0, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}

int myyuvb_dct_decompress_shard_dev(myyuvb_ctx* c, const uint8_t* root_payload, uint64_t payload_size, uint32_t w, uint32_t h,
                                    const uint8_t quality[3], uint32_t rank, uint32_t world, uint32_t root, const uint32_t* rows,
                                    void* const* ctrl, uint8_t* d_band_out, uint8_t* root_iyuv, uint32_t epoch) {
  if (!c || !root_payload || !quality || !d_band_out) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  if (h % 16) return fail(MYYUVB_ERR_HEIGHT, "Error. height % 8 must be 0");
  ShardPeers S;
  if ((rc = shard_peers(h, rank, world, root, rows, ctrl, epoch, &S))) return rc;
  if ((uintptr_t)d_band_out & 7) return fail(MYYUVB_ERR_ARG, "device output must be 8-byte aligned");
  CU(cudaSetDevice(c->device));
  const uint32_t y0 = rows[rank], y1 = rows[rank + 1];
  const FrameGeom g = band_geom(w, h, y0, y1, false, kDecTile);
  const FrameGeom full = make_geom(w, h, 1, kDecTile);
  Workspace ws{};
  if ((rc = ensure_workspace(c, g, false, &ws))) return rc;
  QTables qt;
  make_qtables(quality, &qt);
  const uint32_t k_lo[3] = {(y0 / 8) * (w / 8), (y0 / 16) * (w / 16), (y0 / 16) * (w / 16)};
  if (rank == root) launch_shard_go(S, c->stream);
  if ((rc = c->d_in.reserve((uint64_t)g.nblk_frame * 256 + 64))) return rc;  // the band's part of the payload, pulled from the root
  launch_decompress_shard(root_payload, payload_size, g, full.nblk, k_lo, qt, d_band_out, c->d_in.as<uint8_t>(), S, ws, c->stream);
  if (root_iyuv && y1 > y0) launch_shard_push(d_band_out, root_iyuv, w, h, y0, y1, c->stream);  // the band into the root's frame
  launch_shard_done(S, ws, (uint64_t)w * h * 3 / 2, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}

int myyuvb_shard_result(myyuvb_ctx* c, const void* ctrl_local, uint64_t* total_size) {
  if (!c || !ctrl_local) return fail(MYYUVB_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  const int rc = myyuvb_batch_status(c);  // synchronises; data-dependent errors of this rank's part
  if (rc) return rc;
  ShardCtrl h;
  CU(cudaMemcpy(&h, ctrl_local, sizeof(h), cudaMemcpyDeviceToHost));
  if (h.status & kFlagShardTimeout) return fail(MYYUVB_ERR_SHARD_TIMEOUT, "shard: a rank of the group did not arrive within 2 s");
  if (total_size) *total_size = h.total;
  return MYYUVB_OK;
}

// ------------------------------------------------------------------------------------------------
// Consumers of decoded frames that stay on the device: plane accessors, getPixel, display RGB
// ------------------------------------------------------------------------------------------------
int myyuvb_iyuv_planes(const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t* planes[3], uint32_t widths[3], uint32_t heights[3]) {
  if (!iyuv || !planes) return fail(MYYUVB_ERR_ARG, "null argument");
  // YUV::getYUVPlanes for IYUV (order Y, U, V; 8 + 2 + 2 bits per pixel, myyuv_yuv.cpp:383-421) and getWidthHeightChannel
  planes[0] = iyuv;
  planes[1] = iyuv + (uint64_t)w * h;
  planes[2] = iyuv + (uint64_t)w * h * 5 / 4;
  if (widths) { widths[0] = w; widths[1] = widths[2] = w / 2; }
  if (heights) { heights[0] = h; heights[1] = heights[2] = h / 2; }
  return MYYUVB_OK;
}

int myyuvb_get_pixels_dev(myyuvb_ctx* c, const uint8_t* d_iyuv, uint32_t w, uint32_t h, uint32_t n, const uint32_t* d_xy, uint8_t* d_yuv_out) {
  if (!c || !d_iyuv || (n && (!d_xy || !d_yuv_out))) return fail(MYYUVB_ERR_ARG, "null argument");
  if (w == 0 || h == 0 || (w % 2) || (h % 2)) return fail(MYYUVB_ERR_EVEN, "Error. width and height must be even");
  if ((uint64_t)w * h * 3 / 2 > 0xffffffffull) return fail(MYYUVB_ERR_TOO_LARGE, "Error. image does not fit the format's uint32 sizes");
  CU(cudaSetDevice(c->device));
  if (!c->d_counters.p) {
    int rc;
    if ((rc = c->d_counters.reserve(64))) return rc;
    CU(cudaMemsetAsync(c->d_counters.p, 0, 64, c->stream));
  }
  launch_get_pixels(d_iyuv, w, h, n, d_xy, d_yuv_out, c->d_counters.as<uint32_t>() + 1, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}

int myyuvb_iyuv_to_rgba_batch_dev(myyuvb_ctx* c, const uint8_t* d_iyuv, uint32_t w, uint32_t h, uint32_t n_frames, int flip_rows, uint8_t* d_rgba) {
  if (!c || !d_iyuv || !d_rgba || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  if (w == 0 || h == 0 || (w % 4) || (h % 2)) return fail(MYYUVB_ERR_WIDTH, "Error. width % 4 and height % 2 must be 0");
  if (((uintptr_t)d_iyuv & 3) || ((uintptr_t)d_rgba & 15)) return fail(MYYUVB_ERR_ARG, "device buffers must be 4-byte (input) / 16-byte (output) aligned");
  CU(cudaSetDevice(c->device));
  launch_iyuv_to_rgba(d_iyuv, d_rgba, w, h, n_frames, flip_rows, c->stream);
  CU(cudaGetLastError());
  return MYYUVB_OK;
}

// decompress + display conversion, chunked like the XRGB -> payload pipeline so that the IYUV frames are converted while
// they are still in L2 (d_iyuv: NULL, or n_frames * w*h*3/2 bytes that receive the decoded frames as well)
int myyuvb_dct_decompress_to_rgba_batch_dev(myyuvb_ctx* c, const uint8_t* d_payloads, const uint64_t* d_offsets, uint32_t w, uint32_t h,
                                            const uint8_t quality[3], uint32_t n_frames, uint32_t chunk_frames, int flip_rows,
                                            uint8_t* d_iyuv, uint8_t* d_rgba) {
  if (!c || !d_payloads || !d_offsets || !quality || !d_rgba || n_frames == 0) return fail(MYYUVB_ERR_ARG, "null argument");
  int rc;
  if ((rc = check_quality(quality))) return rc;
  if ((rc = check_dims(w, h))) return rc;
  CU(cudaSetDevice(c->device));
  const uint64_t frame_bytes = (uint64_t)w * h * 3 / 2;
  if (chunk_frames == 0) chunk_frames = (uint32_t)std::max<uint64_t>(1, (48ull << 20) / frame_bytes);  // IYUV + RGBA of a chunk fit L2
  chunk_frames = std::min(chunk_frames, n_frames);
  uint8_t* ring = nullptr;
  if (!d_iyuv) {
    if ((rc = c->d_in.reserve((uint64_t)chunk_frames * frame_bytes))) return rc;
    ring = c->d_in.as<uint8_t>();
  }
  for (uint32_t f0 = 0; f0 < n_frames; f0 += chunk_frames) {
    const uint32_t nf = std::min(chunk_frames, n_frames - f0);
    uint8_t* iy = d_iyuv ? d_iyuv + (uint64_t)f0 * frame_bytes : ring;
    if ((rc = myyuvb_dct_decompress_batch_dev(c, d_payloads, d_offsets + f0, w, h, quality, nf, iy))) return rc;
    if ((rc = myyuvb_iyuv_to_rgba_batch_dev(c, iy, w, h, nf, flip_rows, d_rgba + (uint64_t)f0 * w * h * 4))) return rc;
  }
  return MYYUVB_OK;
}

}  // extern "C"
