// kernels.h -- internal launch interface between capi.cu (host logic) and kernels.cu (sm_100a kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace myyuvb {

// A tile = the blocks one CTA of a codec kernel takes per ticket, one per thread.  Two shapes per kernel, chosen at build time:
// 128 blocks per 4-warp CTA (blocks sorted across the tile, CTA barriers between the phases) or 32 blocks per 1-warp CTA
// (-DMYB_ENC_WARP_TILES=1 / -DMYB_DEC_WARP_TILES=1: no warp ever waits for another one).
#ifndef MYB_ENC_WARP_TILES
#define MYB_ENC_WARP_TILES 0
#endif
#ifndef MYB_DEC_WARP_TILES
#define MYB_DEC_WARP_TILES 0
#endif
#ifndef MYB_ENC_THREADS
#define MYB_ENC_THREADS (MYB_ENC_WARP_TILES ? 32 : 128)
#endif
#ifndef MYB_DEC_THREADS
#define MYB_DEC_THREADS (MYB_DEC_WARP_TILES ? 32 : 256)  // 256: natural content -5..-7 % against 128 (blocks sorted across a larger tile), headline -0.6 %
#endif
constexpr int kEncThreads = MYB_ENC_THREADS;  // threads per CTA of dct_compress_kernel (32, 64 or 128)
constexpr int kDecThreads = MYB_DEC_THREADS;  // threads per CTA of dct_decompress_kernel
constexpr int kEncPasses = 1;      // compress: passes per tile
constexpr int kEncTile = kEncThreads * kEncPasses;
constexpr int kDecTile = kDecThreads;
constexpr int kCtaThreads = 128;   // threads per CTA of heavy15_kernel

// error bits raised by kernels (OR-ed into Workspace::flags)
enum : uint32_t {
  kFlagCapacity = 1u << 0,
  kFlagDctYuvSize = 1u << 1,
  kFlagPlaneSize = 1u << 2,
  kFlagHuffman = 1u << 3,
};

// Quantisation tables of the three planes, computed on the host with the reference's float expression
// (DCT.cpp:286-290).  q = divisor / dequantisation factor, rq = correctly rounded 1/q.
struct alignas(16) QQuad {
  float rx, ry, nqx, nqy;
};
struct alignas(16) QTables {
  float q[3][64];        // row-major, decoder side (coef * q)
  QQuad rq[3][32];       // [b/2 * 8 + a] = (1/q[a][b], 1/q[a][b+1], -q[a][b], -q[a][b+1]) for even b: the encoder's column-pair
                         // lanes, one 128-bit uniform load per pair of coefficients
};

// Geometry shared by all frames of a batch.
struct FrameGeom {
  uint32_t width, height, n_frames;
  uint32_t pw[3], ph[3];          // plane width / height in pixels
  uint32_t bw[3];                 // plane width in 8x8 blocks
  uint32_t bw_magic[3];           // floor(2^32 / bw): block index -> block row without a division (block_row_col)
  uint32_t nblk[3];               // blocks per plane
  uint32_t tile_blocks;           // blocks per tile (kEncTile or kDecTile)
  uint32_t tiles[3];              // tiles per plane
  uint32_t tiles_per_frame;
  uint32_t nblk_frame;            // blocks per frame
  uint64_t plane_off[3];          // byte offset of each plane inside one IYUV frame
  uint64_t frame_bytes;           // w*h*3/2
};

FrameGeom make_geom(uint32_t width, uint32_t height, uint32_t n_frames, uint32_t tile_blocks);

// Device scratch owned by a context (sized for the current batch by capi.cu).
struct Workspace {
  uint64_t* plane_start;   // [n_frames*3 + 1] code bytes before each plane (compress)
  uint64_t* frame_base;    // [n_frames] code bytes of the batch before each frame (compress)
  uint32_t* counters;      // [0] tile ticket, [1] error flags, [2..3] u64 scratch bump allocator, [4] deferred blocks, [5] of those: more
                           // than 15 symbols, [6] tiles with deferred blocks, [7] blocks with more than 32 symbols, [8] blocks with more than 8 symbols
  uint8_t* chunk_sizes;    // [n_frames * nblk_frame] per-block chunk size, linear block order (compress)
  uint8_t* overflow;       // [grid * kEncTile * 256] staging overflow area (compress; a chunk is at most 255 bytes)
  uint8_t* scratch;        // [scratch_cap] chunk bytes of all tiles in completion order (compress, pass 1)
  uint64_t scratch_cap;
  uint64_t* tile_pos;      // [total tiles] position of each tile's bytes in scratch
  uint32_t* tile_total;    // [total tiles] chunk bytes of each tile
  uint64_t* tile_prefix;   // [total tiles] chunk bytes before each tile: inside its frame (compress,
                           // scan_frame_tiles_kernel) or inside its plane (decompress, dec_scan_planes_kernel)
  // compress: blocks with more symbols than the fast path takes are queued and coded 32 at a time by heavy_blocks_kernel
  uint4* heavy_rec;        // [heavy_cap] {block index in the batch, tile, message length, -}; x = 0xffffffff: slot not used
  uint16_t* heavy_coef;    // [heavy_cap * 64] the block's coefficient words, zigzag order
  uint8_t* heavy_bytes;    // [heavy_cap * 256] its chunk
  uint32_t* block_slot;    // [blocks of the batch] queue slot of a deferred block, 0xffffffff otherwise
  uint32_t* heavy_list;    // [heavy_cap] queue slots of the blocks with more than 15 distinct symbols (heavy15_kernel's overflow); later
                           //             the tiles with queued blocks (place_tiles_kernel)
  uint32_t* heavy_list2;   // [heavy_cap] queue slots of the blocks with more than 32 distinct symbols
  uint32_t heavy_cap;      // 0: nothing is deferred
  void* plane_desc;        // [n_frames*3] PlaneDesc (decompress)
  int code_in_place;       // compress: launch the build of dct_compress_kernel that codes up to 15 symbols in place (see kernels.cu)
  uint32_t* queue_stats;   // compress: device-side address of two mapped host words {blocks queued, blocks} the launch reports, or nullptr
  int grid;                // persistent grid size of the codec kernels
  cudaEvent_t k_begin, k_end;  // recorded around the main codec kernel of each launch (myyuvb_last_kernel_ms)
};

struct PlaneDesc {
  uint64_t sizes_off;     // offset of chunk_size[] in the payload buffer
  uint64_t content_off;   // offset of content[]
  uint32_t content_size;
  uint32_t ok;
};

// ---- one image sharded over the GPUs of a box ----
constexpr int kShardMaxWorld = 16;
constexpr uint32_t kFlagShardTimeout = 1u << 4;
constexpr uint32_t kFlagBounds = 1u << 5;

// Where this rank's band goes in the root's payload buffer; written by shard_exchange_kernel, read by the place / finalize kernels
struct ShardPlace {
  uint64_t content_dst[3];  // byte position of the band's content of plane p
  uint64_t sizes_dst[3];    // byte position of the band's chunk-size segment of plane p
  uint64_t total;           // payload bytes of the whole image
  uint32_t ok, pad;
};

// One control block per rank, in device memory that every rank of the group has mapped (CUDA IPC across processes).
// Peers store into it over NVLink; the owner polls it with volatile loads.  Epochs count calls on the group.
struct ShardCtrl {
  uint32_t sizes[kShardMaxWorld][4];  // [q] = content bytes of band q's Y, U, V planes, then the epoch they belong to (stored last)
  uint32_t done[kShardMaxWorld];      // (root's block) epoch whose band q has arrived in the root's buffer
  uint32_t go;                        // the root is ready to receive this epoch (its buffer is free again)
  uint32_t status;                    // (root's block) error bits of the last assembled image
  uint64_t total;                     // (root's block) payload bytes of the last assembled image
  ShardPlace place;                   // this rank's destinations for the current epoch
};

struct ShardPeers {
  ShardCtrl* ctrl[kShardMaxWorld];    // every rank's control block as mapped on this device; ctrl[rank] is local
  uint32_t row[kShardMaxWorld + 1];   // luma pixel rows [row[q], row[q+1]) = band of rank q (multiples of 16)
  uint32_t rank, world, root, epoch;
};

int codec_grid_size(int device, bool encoder);

// Band of a sharded image (n_frames == 1 in g, plane_off may point into a full frame).  out: the root's payload buffer as mapped
// here.  Coding and the two scans run locally; shard_exchange_kernel trades the three content sizes with every peer over
// NVLink and derives the band's destinations; the place / finalize kernels then store the band straight into the root's
// buffer; shard_done_kernel reports to the root, which waits for all bands.  Everything is stream ordered: no host round trip.
void launch_compress_shard(const uint8_t* d_iyuv, const FrameGeom& g, const uint32_t nblk_full[3], const QTables& qt, uint8_t* out,
                           uint64_t out_cap, const ShardPeers& peers, const Workspace& ws, cudaStream_t s);
// Decoding side: payload = the root's payload buffer as mapped here; the band's part of it is pulled into d_local
// (256 bytes per block of the band) and decoded into d_band (band geometry g).
void launch_decompress_shard(const uint8_t* payload, uint64_t payload_size, const FrameGeom& g, const uint32_t nblk_full[3],
                             const uint32_t k_lo[3], const QTables& qt, uint8_t* d_band, uint8_t* d_local, const ShardPeers& peers,
                             const Workspace& ws, cudaStream_t s);
// the decoded band (an image of its own) into rows [y0, y1) of the root's w x h frame
void launch_shard_push(const uint8_t* d_band, uint8_t* root_frame, uint32_t w, uint32_t h, uint32_t y0, uint32_t y1, cudaStream_t s);
// last kernel of a sharded call: tells the root that this rank's part of `epoch` is in place; on the root, waits for all ranks
void launch_shard_done(const ShardPeers& peers, const Workspace& ws, uint64_t total_if_known, cudaStream_t s);
void launch_shard_go(const ShardPeers& peers, cudaStream_t s);  // root only, first kernel of its call

void launch_xrgb_to_iyuv(const uint8_t* d_bgrx, uint8_t* d_iyuv, uint32_t w, uint32_t h, int bottom_up, uint32_t n_frames,
                         cudaStream_t s);
// pixel_bytes: 4 (B,G,R,X) or 3 (B,G,R)
void launch_bgr_to_iyuv(const uint8_t* d_px, uint32_t pixel_bytes, uint8_t* d_iyuv, uint32_t w, uint32_t h, int bottom_up,
                        uint32_t n_frames, cudaStream_t s);
// d_base: device pointer to the position of the first payload inside d_out (nullptr: 0).  finalize writes
// d_offsets[0 .. n_frames]; a chained launch passes d_base = &offsets[first frame], which the previous chunk wrote.
void launch_compress(const uint8_t* d_iyuv, const FrameGeom& g, const QTables& qt, uint8_t* d_out, uint64_t out_cap,
                     uint64_t* d_offsets, const uint64_t* d_base, const Workspace& ws, cudaStream_t s);
void launch_decompress(const uint8_t* d_payloads, const uint64_t* d_offsets, const FrameGeom& g, const QTables& qt,
                       uint8_t* d_iyuv, const Workspace& ws, cudaStream_t s);

// YUV::getPixel for n coordinate pairs (x, y): out[3 i ..] = {Y, U, V}; out-of-range coordinates raise kFlagBounds in *flags
void launch_get_pixels(const uint8_t* d_iyuv, uint32_t w, uint32_t h, uint32_t n, const uint32_t* d_xy, uint8_t* d_out, uint32_t* flags,
                       cudaStream_t s);
// IYUV -> RGBA8 with the viewer's fragment-shader arithmetic (see kernels.cu); flip: rows bottom-up
void launch_iyuv_to_rgba(const uint8_t* d_iyuv, uint8_t* d_rgba, uint32_t w, uint32_t h, uint32_t n_frames, int flip, cudaStream_t s);

// dst / src: device memory or the device-side address of mapped pinned host memory
void launch_sm_copy(uint8_t* dst, const uint8_t* src, uint64_t bytes, cudaStream_t s);
// h_dst_devptr: device-side address of mapped pinned host memory
void launch_publish_words(uint32_t* h_dst_devptr, const uint32_t* d_src, uint32_t n, cudaStream_t s);  // n 32-bit words

extern thread_local uint64_t g_launches;

// per-phase clock sums of the codec kernels, [2][12] (encoder, decoder); zeros unless built with -DMYYUVB_PHASE_CLOCKS
void read_phase_clocks(unsigned long long* out, int reset);

}  // namespace myyuvb
