// class_diff.cpp -- TEST INFRASTRUCTURE.  A transcript of the host-side behaviour of the myyuv::BMP / myyuv::YUV class API (everything
// that needs no device: loading, header normalisation, validity rules, orientation handling, accessors, ownership, dump, exception
// texts) for every file of a directory.  The same source is linked once against the unmodified reference library
// (oracle/_ref/serial/libmyyuv_lib.so) and once against the drop-in library (yuv-manipulations-2_b200/lib/libmyyuv_lib.so);
// tests/test_class_diff.py requires the two transcripts to be identical.  Calls that reach the codec (YUV(bmp, fmt), compress /
// decompress of real data) are left to the GPU tests.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include <myyuv.hpp>

using namespace myyuv;

static uint64_t fnv(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}
static uint64_t file_hash(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  std::vector<uint8_t> b((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  return fnv(b.data(), b.size()) ^ (uint64_t)b.size();
}
static void guarded(const char* what, const std::function<void()>& f) {
  try {
    f();
  } catch (const std::exception& e) {
    std::string m = e.what();  // messages that carry a path: keep the text in front of it
    const size_t slash = m.find('/');
    if (slash != std::string::npos) m = m.substr(0, slash);
    printf("  %s -> exception \"%s\"\n", what, m.c_str());
  }
}

static void show(const BMP& b, const char* tag) {
  const BMPHeader& h = b.header;
  printf("  [%s] type %c%c file_size %u data_pos %u header_size %u w %d h %d planes %u bits %u comp %u sizeimg %u ppm %d %d used %u imp %u\n", tag,
         h.type[0], h.type[1], h.file_size, h.data_pos, h.header_size, h.width, h.height, h.planes, h.bit_count, h.compression,
         h.size_image_for_compression, h.x_pixels_per_meter, h.y_pixels_per_meter, h.colors_used, h.colors_important);
  printf("  [%s] masks %08x %08x %08x %08x space %08x valid %d validHeader %d true %u x %u image %u data %016llx\n", tag, b.color_header.red_mask,
         b.color_header.green_mask, b.color_header.blue_mask, b.color_header.alpha_mask, b.color_header.color_space, (int)b.isValid(),
         (int)b.isValidHeader(), b.trueWidth(), b.trueHeight(), b.imageSize(),
         b.data ? (unsigned long long)fnv(b.data, b.imageSize()) : 0ull);
}

static void bmp_case(const std::string& path, const std::string& tmp) {
  printf("BMP %s\n", path.substr(path.rfind('/') + 1).c_str());
  guarded("load", [&] {
    BMP b(path);
    show(b, "loaded");
    guarded("colorData", [&] {
      uint8_t* c = b.colorData();
      printf("  colorData %016llx\n", (unsigned long long)fnv(c, b.imageSize()));
      delete[] c;
    });
    if (!(b.header.width > 0 && b.header.height < 0))  // that branch of the reference loops on an unsigned compare with a negative height
      guarded("colorDataFlipped", [&] {
        uint8_t* c = b.colorDataFlipped();
        printf("  colorDataFlipped %016llx\n", (unsigned long long)fnv(c, b.imageSize()));
        delete[] c;
      });
    BMP copy(b);
    show(copy, "copy");
    BMP small;  // assignment into an empty object, then into one that already owns a larger / smaller buffer
    small = b;
    show(small, "assigned");
    BMP other(b);
    other.header.width = b.header.width / 2 - (b.header.width / 2) % 4;  // pretend it is a smaller image: its buffer is larger than needed
    if (other.header.width != 0) {
      BMP target(b);
      target = other;
      printf("  assign smaller: image %u data %016llx\n", target.imageSize(), (unsigned long long)fnv(target.data, target.imageSize()));
    }
    BMP moved(std::move(copy));
    show(moved, "moved");
    printf("  moved-from: data %s\n", copy.data ? "kept" : "null");
    const std::string out = tmp + "/out.bmp";
    guarded("dump", [&] {
      b.dump(out);
      printf("  dump %016llx\n", (unsigned long long)file_hash(out));
      BMP again(out);
      show(again, "reloaded");
    });
  });
}

template <class A>
static void arr(const char* name, const A& a) {
  printf("  %s", name);
  for (auto v : a) printf(" %llu", (unsigned long long)v);
  printf("\n");
}

static void show(const YUV& y, const char* tag) {
  const YUVHeader& h = y.header;
  printf("  [%s] type %c%c fourcc %08x data_size %u comp %u params %u @%u w %u h %u data_pos %u valid %d validHeader %d compressed %d\n", tag, h.type[0],
         h.type[1], h.fourcc_format, h.data_size, h.compression, h.compression_params_size, h.compression_params_pos, h.width, h.height, h.data_pos,
         (int)y.isValid(), (int)y.isValidHeader(), (int)y.isCompressed());
  printf("  [%s] getters %08x %u %u %u %u group %d/%d data %016llx params %016llx\n", tag, y.getFourccFormat(), y.getCompression(), y.getWidth(),
         y.getHeight(), y.getDataSize(), (int)y.getFormatGroup(), (int)YUV::getFormatGroup(h.fourcc_format),
         y.data ? (unsigned long long)fnv(y.data, h.data_size) : 0ull,
         y.compression_params ? (unsigned long long)fnv(y.compression_params, h.compression_params_size) : 0ull);
}

static void yuv_case(const std::string& path, const std::string& tmp) {
  printf("YUV %s\n", path.substr(path.rfind('/') + 1).c_str());
  guarded("load", [&] {
    YUV y(path);
    show(y, "loaded");
    guarded("getImageSize", [&] { printf("  imageSize %u\n", y.getImageSize()); });
    guarded("fractions", [&] {
      arr("resolutionFraction", y.getResolutionFraction());
      arr("formatSizeBits", y.getFormatSizeBits());
      arr("planesOrder", y.getYUVPlanesOrder());
      for (uint8_t c = 0; c < 3; c++) arr("widthHeightChannel", y.getWidthHeightChannel(c));
    });
    if (!y.isCompressed()) {
      guarded("planes", [&] {
        const auto pl = static_cast<const YUV&>(y).getYUVPlanes();
        printf("  planes");
        for (auto p : pl) printf(" %lld", p ? (long long)(p - y.data) : -1ll);
        printf("\n");
      });
      const uint32_t w = y.getWidth(), h = y.getHeight();
      // (the reference's chroma index runs past its buffer on the last row for x >= w / 2: undefined there, not asked for here)
      const uint32_t xs[] = {0, 1, w / 2, w - 1, w / 2 - 1, w, 5}, ys[] = {0, 1, h / 2, h - 2, h - 1, 3, h};
      for (int i = 0; i < 7; i++)
        guarded("getPixel", [&] { arr("pixel", y.getPixel(xs[i], ys[i])); });
      if (getenv("CLASS_DIFF_EXTRA"))  // the drop-in library alone (sanitizer run): where the reference's index leaves its buffer, the sample is 0
        guarded("getPixel", [&] { arr("pixel (last row, right half)", y.getPixel(w - 1, h - 1)); });
      guarded("decompress (not compressed)", [&] {
        YUV d = y.decompress();
        show(d, "decompress copy");
      });
      guarded("compress (unknown compression)", [&] {
        const uint8_t q[3] = {50, 50, 50};
        YUV c = y.compress(7, q, 3);
        show(c, "?");
      });
      guarded("compress (wrong parameter count)", [&] {
        const uint8_t q[3] = {50, 50, 50};
        YUV c = y.compress(YUV::Compressions::DCT, q, 2);
        show(c, "?");
      });
    } else {
      guarded("compress (already compressed)", [&] {
        const uint8_t q[3] = {50, 50, 50};
        YUV c = y.compress(YUV::Compressions::DCT, q, 3);
        show(c, "?");
      });
    }
    YUV copy(y);
    show(copy, "copy");
    YUV assigned;
    assigned = y;
    show(assigned, "assigned");
    YUV moved(std::move(copy));
    show(moved, "moved");
    printf("  moved-from: data %s params %s\n", copy.data ? "kept" : "null", copy.compression_params ? "kept" : "null");
    const std::string out = tmp + "/out.myyuv";
    guarded("dump", [&] {
      y.dump(out);
      printf("  dump %016llx\n", (unsigned long long)file_hash(out));
      YUV again(out);
      show(again, "reloaded");
    });
  });
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string tmp = argv[1];
  printf("implemented: %d %d %d %d\n", (int)YUV::isImplementedFormat(YUV::FourccFormats::IYUV), (int)YUV::isImplementedFormat(YUV::FourccFormats::IYUV, YUV::Compressions::DCT),
         (int)YUV::isImplementedFormat(0x32595559), (int)YUV::isImplementedFormat(YUV::FourccFormats::IYUV, 9));
  {
    BMP empty;
    show(empty, "default");
    guarded("colorData of an empty BMP", [&] { delete[] empty.colorData(); });
    guarded("colorDataFlipped of an empty BMP", [&] { delete[] empty.colorDataFlipped(); });
    YUV none;
    show(none, "default");
    guarded("getImageSize of an empty YUV", [&] { printf("  %u\n", none.getImageSize()); });
  }
  for (int i = 2; i < argc; i++) {
    const std::string p = argv[i];
    if (p.size() > 4 && p.substr(p.size() - 4) == ".bmp") bmp_case(p, tmp);
    else yuv_case(p, tmp);
  }
  return 0;
}
