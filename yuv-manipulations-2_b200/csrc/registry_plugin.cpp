// registry_plugin.cpp -- INTEGRATION.md form B: the reference's OWN libmyyuv_lib.so stays in place and this plugin
// overwrites the three registry slots of the hot path (myyuv_lib/myyuv_yuv.hpp:106,111,116; default entries at
// myyuv_lib/myyuv_yuv.cpp:88,130,146) in a static initialiser with entries that call the C ABI of libmyyuvb200.so.
// Compiled against the REFERENCE's headers (oracle/Makefile, target `plugin`); loaded with LD_PRELOAD or linked into the
// application.  Everything the reference's classes do around the three slots -- file I/O, validity checks, the CLI -- is
// the reference's own code, which makes this the same-process A/B: MYYUVB_PLUGIN=0 leaves the slots alone.
#include <myyuv.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>

#include "../../include/myyuvb200.h"

namespace {
std::mutex g_mtx;  // a context serves one thread at a time; the reference's functions are stateless

myyuvb_ctx* context() {
  static myyuvb_ctx* c = [] {
    myyuvb_ctx* p = nullptr;
    const char* dev = std::getenv("MYYUVB_DEVICE");
    if (myyuvb_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &p)) throw std::runtime_error(myyuvb_last_error());
    return p;
  }();
  return c;
}
void check(int rc) {
  if (rc) throw std::runtime_error(myyuvb_last_error());  // the text the reference throws for the same condition
}

struct Install {
  Install() {
    using myyuv::BMP;
    using myyuv::YUV;
    const char* sw = std::getenv("MYYUVB_PLUGIN");
    if (sw && sw[0] == '0') return;
    if (std::getenv("MYYUVB_PLUGIN_VERBOSE")) std::fprintf(stderr, "[myyuvb200 plugin] registry slots IYUV / DCT overridden\n");

    // replaces the lambda at myyuv_yuv.cpp:89-127 (+ the row flip of BMP::colorData, myyuv_bmp.cpp:95-98)
    YUV::bmp_to_yuv_map[YUV::FourccFormats::IYUV] = [](const BMP& bmp) {
      if (!bmp.isValid()) throw std::runtime_error("BMP data is invalid");
      const uint32_t w = bmp.trueWidth(), h = bmp.trueHeight();
      YUV out;
      out.header.fourcc_format = YUV::FourccFormats::IYUV;
      out.header.width = w;
      out.header.height = h;
      out.header.data_size = w * h * 3 / 2;
      out.header.data_pos = sizeof(myyuv::YUVHeader);
      out.data = new uint8_t[out.header.data_size];
      const auto convert = bmp.header.bit_count == 24 ? myyuvb_bgr24_to_iyuv : myyuvb_xrgb_to_iyuv;
      std::lock_guard<std::mutex> lock(g_mtx);
      if (bmp.header.width > 0 && bmp.header.height != 0) {
        check(convert(context(), bmp.data, w, h, bmp.header.height > 0 ? 1 : 0, out.data));
      } else {  // reversed pixel order (negative width): the reference's own colorData() reorders, rows are then top-down
        std::unique_ptr<uint8_t[]> px(bmp.colorData());
        check(convert(context(), px.get(), w, h, 0, out.data));
      }
      return out;
    };

    // replaces myyuv_yuv.cpp:132-142 -> myyuvDCT::compress_DCT_planar (DCT.cpp:371-430)
    YUV::compress_map[YUV::Compressions::DCT][YUV::FourccFormats::IYUV] = [](const YUV& src, const void* params, uint32_t n) {
      if (n != 3) throw std::runtime_error("Error compression: incorrect parameters count. 3 parameters required");
      const uint8_t* q = static_cast<const uint8_t*>(params);
      std::lock_guard<std::mutex> lock(g_mtx);
      uint32_t size = 0;
      check(myyuvb_dct_compress_begin(context(), src.data, src.header.width, src.header.height, q, &size));
      YUV out;
      out.header = src.header;  // DCT.cpp:390-396
      out.header.compression = YUV::Compressions::DCT;
      out.header.compression_params_size = 3;
      out.header.compression_params_pos = sizeof(myyuv::YUVHeader);
      out.header.data_pos = sizeof(myyuv::YUVHeader) + 3;
      out.header.data_size = size;
      out.compression_params = new uint8_t[3]{q[0], q[1], q[2]};
      out.data = new uint8_t[size];  // exact size, released by ~YUV with delete[] (myyuv_yuv.cpp:243-246)
      check(myyuvb_dct_compress_fetch(context(), out.data, size));
      return out;
    };

    // replaces myyuv_yuv.cpp:148-158 -> myyuvDCT::decompress_DCT_planar (DCT.cpp:432-488)
    YUV::decompress_map[YUV::Compressions::DCT][YUV::FourccFormats::IYUV] = [](const YUV& src) {
      if (src.header.compression_params_size != 3)
        throw std::runtime_error("Error decompression: incorrect parameters count. 3 parameters required");
      YUV out;
      out.header = src.header;  // DCT.cpp:447-453
      out.header.compression = YUV::Compressions::NONE;
      out.header.compression_params_size = 0;
      out.header.compression_params_pos = 0;
      out.header.data_pos = sizeof(myyuv::YUVHeader);
      out.header.data_size = src.header.width * src.header.height * 3 / 2;
      out.data = new uint8_t[out.header.data_size];
      std::lock_guard<std::mutex> lock(g_mtx);
      check(myyuvb_dct_decompress(context(), src.data, src.header.data_size, src.header.width, src.header.height, src.compression_params,
                                  out.data));
      return out;
    };
  }
};
// the registries are defined in the reference library, which this plugin links against, so they are constructed first
Install g_install;
}  // namespace
