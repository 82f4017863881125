/*
 * myyuvb200.h -- C ABI of the B200-native (sm_100a) implementation of myyuv_lib's hot path:
 * BMP XRGB8888 -> IYUV conversion, DCT-q compression, decompression, compressed-payload layout.
 *
 * This is the drop-in boundary.  The reference (mahbhlddnhakkh/yuv-manipulations-2) has no FFI of its
 * own: its plugin boundary is three public static registries of std::function in class myyuv::YUV
 * (myyuv_lib/myyuv_yuv.hpp:106,111,116; entries defined at myyuv_lib/myyuv_yuv.cpp:88,130,146).  Each
 * entry point below is what one of those registry entries binds to; INTEGRATION.md shows the binding.
 * Plain pointers and sizes only; every function returns MYYUVB_OK (0) or an error code, and
 * myyuvb_last_error() returns the message the reference would have thrown for that condition.
 *
 * Payload layout (identical to the reference, DCT.cpp:16-73,112-173), little-endian, packed:
 *   payload := u32 planes_sizes[3]  plane[Y] plane[U] plane[V]
 *   plane   := u32 n_chunks  u32 content_size  u8 chunk_size[n_chunks]  u8 content[content_size]
 *   chunk   := u16 code_bits  u8 table_bytes  group*  code_stream[(code_bits+7)/8]      (Huffman.cpp:279-326)
 * It is what YUV::data holds when header.compression == DCT (header.data_size bytes).
 */
#ifndef MYYUVB200_H
#define MYYUVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MYYUVB_API
#else
#define MYYUVB_API __attribute__((visibility("default")))
#endif

enum {
  MYYUVB_OK = 0,
  MYYUVB_ERR_CUDA = 1,          /* a CUDA runtime call failed (message carries cudaGetErrorString)          */
  MYYUVB_ERR_ARG = 2,           /* null pointer / zero size                                                 */
  MYYUVB_ERR_QUALITY = 3,       /* "Level of quality must be between 1 and 100"       DCT.cpp:378-382,438-442 */
  MYYUVB_ERR_WIDTH = 4,         /* "Error. width % 8 must be 0"                       DCT.cpp:280-282,338-340 */
  MYYUVB_ERR_HEIGHT = 5,        /* "Error. height % 8 must be 0"                      DCT.cpp:283-285,341-343 */
  MYYUVB_ERR_CAPACITY = 6,      /* output buffer smaller than the produced payload                          */
  MYYUVB_ERR_DCTYUV_SIZE = 7,   /* "DCTYUV load bad size"                             DCT.cpp:132-134,143-145 */
  MYYUVB_ERR_PLANE_SIZE = 8,    /* "DCTYUVPlane load bad size" (+ chunks_sizes_size / content_size variants) DCT.cpp:41-55 */
  MYYUVB_ERR_HUFFMAN = 9,       /* "Huffman bad code"                                 Huffman.cpp:121,130,139 */
  MYYUVB_ERR_EVEN = 10,         /* colour conversion needs even width and height      myyuv_yuv.cpp:98       */
  MYYUVB_ERR_TOO_LARGE = 11,    /* sizes do not fit the format's uint32 fields        myyuv_yuv.hpp:20-26    */
  MYYUVB_ERR_SHARD_TIMEOUT = 12,/* sharded image: a rank of the group never arrived                          */
  MYYUVB_ERR_BOUNDS = 13        /* "Image coordinates are out of bounds"              myyuv_yuv.cpp:171-173  */
};

typedef struct myyuvb_ctx myyuvb_ctx; /* one device, one stream, reusable device/pinned scratch; not thread-safe
                                         per context (use one context per thread), contexts are independent. */

/* device: CUDA ordinal.  stream: a cudaStream_t to run on (e.g. the caller's torch stream), or NULL to
 * let the context create its own non-blocking stream. */
MYYUVB_API int myyuvb_ctx_create(int device, void* stream, myyuvb_ctx** out);
MYYUVB_API void myyuvb_ctx_destroy(myyuvb_ctx* ctx);
MYYUVB_API const char* myyuvb_last_error(void); /* thread-local, valid until the next call on this thread */
MYYUVB_API int myyuvb_sync(myyuvb_ctx* ctx);    /* wait for the context's stream */
MYYUVB_API void* myyuvb_stream(myyuvb_ctx* ctx);
/* Device time (CUDA events on the context stream) of the most recent compress launch sequence (code tiles,
 * deferred blocks, two scans, place, headers: 8 kernels) or decompress kernel on this context -- the figure bench.py's roofline uses.
 * Synchronises. */
MYYUVB_API int myyuvb_last_kernel_ms(myyuvb_ctx* ctx, float* ms);

/* Which build of the coding kernel compress launches use on this context: 0 (default) chooses from the share of detailed
 * blocks in the previous launch, 1 always queues blocks with more than 8 distinct symbols for the heavy-block kernels,
 * 2 codes up to 15 symbols in place.  A performance knob only: the bytes produced are identical in every mode. */
MYYUVB_API int myyuvb_set_encoder_mode(myyuvb_ctx* ctx, int mode);

/* Upper bound of a compressed payload for one w x h IYUV frame (a chunk is at most 255 bytes because
 * its size is stored in a uint8, DCT.cpp:19,310). */
MYYUVB_API uint64_t myyuvb_compress_bound(uint32_t width, uint32_t height);

/* ---- host-pointer entry points: what the class API / registries bind (H2D + kernels + D2H inside) ---- */

/* replaces bmp_to_yuv_map[IYUV] (myyuv_yuv.cpp:88-128) with BMP::colorData()'s row flip
 * (myyuv_bmp.cpp:80-103) folded in.  bgrx: width*height*4 bytes, rows as stored in the BMP file;
 * bottom_up != 0 when BMP height > 0.  iyuv_out: width*height*3/2 bytes. */
MYYUVB_API int myyuvb_xrgb_to_iyuv(myyuvb_ctx* ctx, const uint8_t* bgrx, uint32_t width, uint32_t height,
                                   int bottom_up, uint8_t* iyuv_out);

/* The same registry entry for a 24-bit BMP: getYUV444FromRGB2x2 addresses pixel i at byte i * bit_count / 8
 * (myyuv_yuv.cpp:34-41), and the "bit_count == 32" assert at :92 ("TODO: test 24") is compiled out of the Release build
 * the reference's README asks for, so B,G,R triplets are what that build converts.  bgr: width*height*3 bytes, rows as
 * stored in the file (no padding: a valid BMP has width % 4 == 0, myyuv_bmp.cpp:130).  SURVEY 8(f) row 3. */
MYYUVB_API int myyuvb_bgr24_to_iyuv(myyuvb_ctx* ctx, const uint8_t* bgr, uint32_t width, uint32_t height,
                                    int bottom_up, uint8_t* iyuv_out);

/* replaces compress_map[DCT][IYUV] -> myyuvDCT::compress_DCT_planar (myyuv_yuv.cpp:130-143, DCT.cpp:371-430).
 * quality[3]: Y,U,V quality 1..100.  out receives the payload (YUV::data of the compressed image). */
MYYUVB_API int myyuvb_dct_compress(myyuvb_ctx* ctx, const uint8_t* iyuv, uint32_t width, uint32_t height,
                                   const uint8_t quality[3], uint8_t* out, uint64_t out_capacity,
                                   uint32_t* out_size);

/* The same in two steps, for bindings that must return an exact-size allocation (YUV::data is new uint8_t[data_size],
 * released with delete[] in ~YUV, myyuv_yuv.cpp:243-246; alloc site DCT.cpp:162): _begin uploads and codes the image and
 * reports the payload size, the payload stays on the device; _fetch copies it into the caller's buffer. */
MYYUVB_API int myyuvb_dct_compress_begin(myyuvb_ctx* ctx, const uint8_t* iyuv, uint32_t width, uint32_t height,
                                         const uint8_t quality[3], uint32_t* out_size);
MYYUVB_API int myyuvb_dct_compress_fetch(myyuvb_ctx* ctx, uint8_t* out, uint64_t out_capacity);

/* replaces decompress_map[DCT][IYUV] -> myyuvDCT::decompress_DCT_planar (myyuv_yuv.cpp:146-159, DCT.cpp:432-488). */
MYYUVB_API int myyuvb_dct_decompress(myyuvb_ctx* ctx, const uint8_t* payload, uint32_t payload_size, uint32_t width,
                                     uint32_t height, const uint8_t quality[3], uint8_t* iyuv_out);

/* ---- device-pointer batch entry points (asynchronous on the context stream) ----
 * All frames of a batch share width/height/quality.  Frames are independent units. */

/* d_bgrx: n_frames * w*h*4, d_iyuv: n_frames * w*h*3/2. */
MYYUVB_API int myyuvb_xrgb_to_iyuv_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_bgrx, uint32_t width, uint32_t height,
                                             int bottom_up, uint32_t n_frames, uint8_t* d_iyuv);

/* d_bgr: n_frames * w*h*3, d_iyuv: n_frames * w*h*3/2.  Alignment follows the kernel that runs: for width % 8 == 0 (64-bit
 * loads) d_bgr must be 8-byte aligned and, in a batch of more than one frame, w*h*3 a multiple of 8, d_iyuv 8-byte aligned;
 * other widths are read bytewise and need no input alignment (d_iyuv 2-byte aligned). */
MYYUVB_API int myyuvb_bgr24_to_iyuv_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_bgr, uint32_t width, uint32_t height,
                                              int bottom_up, uint32_t n_frames, uint8_t* d_iyuv);

/* Payloads are written back to back into d_out; d_offsets[f] .. d_offsets[f+1] (n_frames+1 entries,
 * device memory) delimit frame f.  Nothing is written past out_capacity: an overflow is reported by
 * myyuvb_batch_status as MYYUVB_ERR_CAPACITY.  n_frames * myyuvb_compress_bound() always suffices. */
MYYUVB_API int myyuvb_dct_compress_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_iyuv, uint32_t width, uint32_t height,
                                             const uint8_t quality[3], uint32_t n_frames, uint8_t* d_out,
                                             uint64_t out_capacity, uint64_t* d_offsets);

/* The whole pipeline of the reference's  YUV(bmp, IYUV).compress(DCT, q)  (myyuv_yuv.cpp:88-128 then DCT.cpp:371-430) for a
 * batch of XRGB8888 frames, device resident: d_bgrx n_frames * w*h*4 (rows as stored in the BMP), payloads and offsets as
 * in myyuvb_dct_compress_batch_dev.  Frames are converted and coded in chunks of chunk_frames (0: about 100 MB of IYUV) so
 * that the intermediate IYUV image is read back from L2, not HBM.  d_iyuv: NULL, or n_frames * w*h*3/2 bytes that receive
 * the IYUV frames as well (the product of the conversion step alone). */
MYYUVB_API int myyuvb_xrgb_dct_compress_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_bgrx, uint32_t width, uint32_t height,
                                                  int bottom_up, const uint8_t quality[3], uint32_t n_frames,
                                                  uint32_t chunk_frames, uint8_t* d_iyuv, uint8_t* d_out,
                                                  uint64_t out_capacity, uint64_t* d_offsets);

/* d_payloads + d_offsets as produced above (any packing is fine as long as frame f occupies
 * [d_offsets[f], d_offsets[f+1])).  d_iyuv: n_frames * w*h*3/2. */
MYYUVB_API int myyuvb_dct_decompress_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_payloads, const uint64_t* d_offsets,
                                               uint32_t width, uint32_t height, const uint8_t quality[3],
                                               uint32_t n_frames, uint8_t* d_iyuv);

/* Synchronises the stream and returns the first data-dependent error raised by the batch calls issued
 * since the previous status call (capacity overflow, malformed payload), MYYUVB_OK otherwise. */
MYYUVB_API int myyuvb_batch_status(myyuvb_ctx* ctx);

/* ---- host-pointer batch entry points: H2D / kernels / D2H pipelined over chunks of frames on three streams.  This is
 * the end-to-end path bench.py times ("e2e").  Any host memory is accepted; with page-locked buffers (myyuvb_host_alloc,
 * cudaHostAlloc, cudaHostRegister) the large transfers run at link speed and the small ones (payloads, offsets) are
 * moved by a kernel, so that two calls running side by side on two contexts do not wait for each other's copies.
 * One context serves one host thread at a time. ---- */
MYYUVB_API int myyuvb_dct_compress_batch_host(myyuvb_ctx* ctx, const uint8_t* iyuv, uint32_t width, uint32_t height,
                                              const uint8_t quality[3], uint32_t n_frames, uint8_t* out,
                                              uint64_t out_capacity, uint64_t* offsets /* n_frames+1 */);
MYYUVB_API int myyuvb_dct_decompress_batch_host(myyuvb_ctx* ctx, const uint8_t* payloads, const uint64_t* offsets,
                                                uint32_t width, uint32_t height, const uint8_t quality[3],
                                                uint32_t n_frames, uint8_t* iyuv_out);

/* pinned host memory helpers (so callers in any language can hand the batch_host calls DMA-able buffers) */
MYYUVB_API int myyuvb_host_alloc(size_t bytes, void** out);
MYYUVB_API void myyuvb_host_free(void* p);

/* number of kernels this library has launched on this thread's contexts since process start (bench.py's gpu_launches) */
MYYUVB_API uint64_t myyuvb_launch_count(void);

/* ---- consumers of decoded frames that stay on the device: what the reference's viewers do with
 * a decoded YUV on the host -- YUV::getYUVPlanes / getWidthHeightChannel (myyuv_yuv.cpp:383-427), YUV::getPixel
 * (yuv_get_pixel_map[IYUV], :162-180) and the fragment shader's YUV -> RGB (myyuv_opengl/viewer/frag_yuv.glsl:18-26) -- on
 * device pointers, so that a viewer can hand GPU-decoded frames to the display without a host round trip (register the
 * plane pointers or the RGBA buffer with cudaGraphicsGLRegisterImage / a pixel-unpack buffer). ---- */
/* plane pointers and sizes of an IYUV frame at `iyuv` (host or device address; pure arithmetic) */
MYYUVB_API int myyuvb_iyuv_planes(const uint8_t* iyuv, uint32_t width, uint32_t height, const uint8_t* planes[3],
                                  uint32_t widths[3], uint32_t heights[3]);
/* getPixel for n coordinate pairs d_xy = {x0, y0, x1, y1, ...}: d_yuv_out[3 i ..] = {Y, U, V}, with the reference's chroma
 * index x / 2 + y * width / 4.  Out-of-range coordinates are reported by myyuvb_batch_status as MYYUVB_ERR_BOUNDS. */
MYYUVB_API int myyuvb_get_pixels_dev(myyuvb_ctx* ctx, const uint8_t* d_iyuv, uint32_t width, uint32_t height, uint32_t n,
                                     const uint32_t* d_xy, uint8_t* d_yuv_out);
/* IYUV -> RGBA8 (R, G, B, 255) with the viewer shader's arithmetic, chroma sampled bilinearly at the luma pixel centres;
 * +-1 LSB of the formula in double precision.  flip_rows != 0: rows bottom-up (GL).  d_rgba: n_frames * w*h*4, 16-byte aligned. */
MYYUVB_API int myyuvb_iyuv_to_rgba_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_iyuv, uint32_t width, uint32_t height,
                                             uint32_t n_frames, int flip_rows, uint8_t* d_rgba);
/* decompress + the conversion above, chunk_frames at a time (0: about 48 MB of IYUV) so that the decoded frames are
 * converted out of L2; d_iyuv: NULL, or n_frames * w*h*3/2 bytes that receive the decoded frames too. */
MYYUVB_API int myyuvb_dct_decompress_to_rgba_batch_dev(myyuvb_ctx* ctx, const uint8_t* d_payloads, const uint64_t* d_offsets,
                                                       uint32_t width, uint32_t height, const uint8_t quality[3],
                                                       uint32_t n_frames, uint32_t chunk_frames, int flip_rows,
                                                       uint8_t* d_iyuv, uint8_t* d_rgba);

/* ---- one very large image sharded over the GPUs of one box (the reference has no counterpart:
 * the property that makes it possible is that every 8x8 block is coded on its own and a plane's content is the blocks'
 * chunks in raster order, DCT.cpp:297-322, layout DCT.cpp:16-33,160-173) ----
 * One context per GPU, each in its own process or thread.  Rank q codes the luma rows [rows[q], rows[q+1]) (whole
 * macroblock rows; myyuvb_shard_rows gives the balanced split, 270 rows of 8K over 8 ranks = 34 x 6 + 33 x 2).
 * ctrl[q] is rank q's control block (myyuvb_shard_ctrl_bytes() of zeroed device memory) as mapped on THIS device;
 * root_out / root_payload / root_iyuv are the root's buffers as mapped on this device.  Across processes the mappings come
 * from the IPC helpers below.  All ranks of a group must make the same call with the same, increasing `epoch` (1, 2, ...):
 * the calls are asynchronous on each context's stream and meet on the device -- content sizes are exchanged by peer
 * stores, every band is stored straight into the root's buffer, the root's stream continues when all bands are in.
 * myyuvb_shard_result (root) synchronises and returns the assembled size.  A rank that never arrives makes the others
 * fail with MYYUVB_ERR_SHARD_TIMEOUT after 2 s instead of hanging. */
MYYUVB_API uint64_t myyuvb_shard_ctrl_bytes(void);
MYYUVB_API int myyuvb_shard_rows(uint32_t height, uint32_t world, uint32_t* rows /* [world + 1] */);
/* d_iyuv: this rank's band as an IYUV image of height rows[rank+1]-rows[rank] (iyuv_is_full_frame == 0), or the whole
 * width x height frame resident on this GPU, of which only the band's rows are read (!= 0). */
MYYUVB_API int myyuvb_dct_compress_shard_dev(myyuvb_ctx* ctx, const uint8_t* d_iyuv, int iyuv_is_full_frame, uint32_t width,
                                             uint32_t height, const uint8_t quality[3], uint32_t rank, uint32_t world,
                                             uint32_t root, const uint32_t* rows, void* const* ctrl, uint8_t* root_out,
                                             uint64_t out_capacity, uint32_t epoch);
/* The band is decoded into d_band_out (an IYUV image of the band's height, local) and, when root_iyuv is not NULL, its
 * three planes are copied into the root's width x height frame. */
MYYUVB_API int myyuvb_dct_decompress_shard_dev(myyuvb_ctx* ctx, const uint8_t* root_payload, uint64_t payload_size,
                                               uint32_t width, uint32_t height, const uint8_t quality[3], uint32_t rank,
                                               uint32_t world, uint32_t root, const uint32_t* rows, void* const* ctrl,
                                               uint8_t* d_band_out, uint8_t* root_iyuv, uint32_t epoch);
MYYUVB_API int myyuvb_shard_result(myyuvb_ctx* ctx, const void* ctrl_local, uint64_t* total_size);
/* device memory that other processes can map: cudaMalloc (zeroed) + cudaIpcGetMemHandle / cudaIpcOpenMemHandle */
MYYUVB_API int myyuvb_ipc_alloc(myyuvb_ctx* ctx, uint64_t bytes, void** d_ptr, uint8_t handle_out[64]);
MYYUVB_API int myyuvb_ipc_open(myyuvb_ctx* ctx, const uint8_t handle[64], void** d_ptr);
MYYUVB_API int myyuvb_ipc_close(myyuvb_ctx* ctx, void* d_ptr);
MYYUVB_API int myyuvb_ipc_free(myyuvb_ctx* ctx, void* d_ptr);

/* Profiling aid: clock sums per phase of the two codec kernels, out24 = [2][12] (compress, decompress).  All zero in the
 * product library; the -DMYYUVB_PHASE_CLOCKS build (lib/libmyyuvb200_clk.so, profiles/phase_clocks.py) fills them. */
MYYUVB_API void myyuvb_phase_clocks(uint64_t* out24, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MYYUVB200_H */
