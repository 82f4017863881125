// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" driver around the UNMODIFIED reference library (built from the
// sources where they lie under /root/reference by oracle/Makefile, output in
// oracle/_ref/).  It lets the Python tests and bench.py's reference arm call the
// reference's own public API -- myyuv::YUV(bmp, IYUV), YUV::compress, YUV::decompress
// (myyuv_lib/myyuv_yuv.hpp:143,313,321) and myyuvDCT::Huffman::fromData/dump/fromDump/
// getData (myyuv_lib/myyuv_DCT/Huffman.hpp:64-86) -- through ctypes.
// Nothing here restates the reference's algorithm; it only marshals buffers.
#include <myyuv.hpp>
#include <myyuv_DCT/Huffman.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
thread_local std::string g_err;
using clk = std::chrono::steady_clock;

int fail(const std::exception& e) { g_err = e.what(); return 1; }

myyuv::YUV make_iyuv(const uint8_t* iyuv, uint32_t w, uint32_t h) {
  myyuv::YUV y;
  y.header.fourcc_format = myyuv::YUV::FourccFormats::IYUV;
  y.header.width = w;
  y.header.height = h;
  y.header.data_size = w * h * 3 / 2;
  y.header.data_pos = sizeof(myyuv::YUVHeader);
  y.data = new uint8_t[y.header.data_size];
  std::memcpy(y.data, iyuv, y.header.data_size);
  return y;
}
}  // namespace

extern "C" {

const char* refshim_last_error() { return g_err.c_str(); }

int refshim_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// px: pixel rows exactly as stored in a BMP file of bit_count bits per pixel; height_signed > 0 = bottom-up file.
int refshim_bmp_to_iyuv(const uint8_t* px, int32_t width, int32_t height_signed, uint32_t bit_count, uint8_t* iyuv_out,
                        double* seconds) {
  try {
    myyuv::BMP bmp;
    bmp.header.width = width;
    bmp.header.height = height_signed;
    bmp.header.bit_count = static_cast<uint16_t>(bit_count);
    bmp.header.planes = 1;
    bmp.header.header_size = bit_count == 32 ? 124 : 40;
    bmp.header.compression = bit_count == 32 ? 3 : 0;
    bmp.header.data_pos = sizeof(myyuv::BMPHeader) + (bit_count == 32 ? sizeof(myyuv::BMPColorHeader) : 0);
    const uint32_t sz = bmp.imageSize();
    bmp.header.file_size = bmp.header.data_pos + sz;
    bmp.data = new uint8_t[sz];
    std::memcpy(bmp.data, px, sz);
    auto t0 = clk::now();
    myyuv::YUV yuv(bmp, myyuv::YUV::FourccFormats::IYUV);
    auto t1 = clk::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    std::memcpy(iyuv_out, yuv.data, yuv.header.data_size);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int refshim_bgrx_to_iyuv(const uint8_t* bgrx, int32_t width, int32_t height_signed, uint8_t* iyuv_out,
                         double* seconds) {
  return refshim_bmp_to_iyuv(bgrx, width, height_signed, 32, iyuv_out, seconds);
}

int refshim_compress(const uint8_t* iyuv, uint32_t w, uint32_t h, const uint8_t* q, uint32_t nq, uint8_t* out,
                     uint32_t out_cap, uint32_t* out_size, double* seconds) {
  try {
    myyuv::YUV src = make_iyuv(iyuv, w, h);
    auto t0 = clk::now();
    myyuv::YUV c = src.compress(myyuv::YUV::Compressions::DCT, q, nq);
    auto t1 = clk::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    *out_size = c.header.data_size;
    if (c.header.data_size > out_cap) { g_err = "refshim: output capacity too small"; return 2; }
    std::memcpy(out, c.data, c.header.data_size);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int refshim_decompress(const uint8_t* payload, uint32_t size, uint32_t w, uint32_t h, const uint8_t* q,
                       uint8_t* iyuv_out, double* seconds) {
  try {
    myyuv::YUV c;
    c.header.fourcc_format = myyuv::YUV::FourccFormats::IYUV;
    c.header.width = w;
    c.header.height = h;
    c.header.compression = myyuv::YUV::Compressions::DCT;
    c.header.compression_params_size = 3;
    c.header.compression_params_pos = sizeof(myyuv::YUVHeader);
    c.header.data_pos = sizeof(myyuv::YUVHeader) + 3;
    c.header.data_size = size;
    c.compression_params = new uint8_t[3]{q[0], q[1], q[2]};
    c.data = new uint8_t[size];
    std::memcpy(c.data, payload, size);
    auto t0 = clk::now();
    myyuv::YUV d = c.decompress();
    auto t1 = clk::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    std::memcpy(iyuv_out, d.data, d.header.data_size);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

// n blocks of 64 int16 (row-major 8x8) -> concatenated chunks, sizes[n].
int refshim_huffman_encode(const int16_t* coef, uint32_t n, uint8_t* out, uint8_t* sizes) {
  try {
    for (uint32_t b = 0; b < n; b++) {
      myyuvDCT::Huffman hf = myyuvDCT::Huffman::fromData(coef + 64 * b);
      uint8_t* p = nullptr;
      uint8_t sz = 0;
      hf.dump(p, sz);
      std::memcpy(out, p, sz);
      delete[] p;
      out += sz;
      sizes[b] = sz;
    }
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int refshim_huffman_decode(const uint8_t* chunks, const uint8_t* sizes, uint32_t n, int16_t* coef_out) {
  try {
    for (uint32_t b = 0; b < n; b++) {
      myyuvDCT::Huffman hf = myyuvDCT::Huffman::fromDump(chunks, sizes[b]);
      hf.getData(coef_out + 64 * b);
      chunks += sizes[b];
    }
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

}  // extern "C"
