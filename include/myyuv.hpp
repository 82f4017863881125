// myyuv.hpp -- drop-in declarations of myyuv::BMP and myyuv::YUV for the B200-native library.
//
// ABI-compatible with the reference's public headers (myyuv_lib/myyuv_bmp.hpp:8-157 and
// myyuv_lib/myyuv_yuv.hpp:11-350): same packed header structs, same data members in the same order, same
// member functions and the same seven public static registries, so code compiled against the reference's
// headers (e.g. the unmodified myyuv_cli/main.cpp) links and runs against lib/libmyyuv_lib.so unchanged.
// The difference is below the registries: the IYUV converter and the DCT compress / decompress entries call
// the CUDA kernels through the C ABI of myyuvb200.h instead of the reference's CPU loops.
#pragma once

#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>

namespace myyuv {

#pragma pack(push, 1)
// BITMAPFILEHEADER + BITMAPINFOHEADER, 54 bytes on disk
struct BMPHeader {
  uint8_t type[2] = { 'B', 'M' };
  uint32_t file_size = 0;
  uint16_t reserved1 = 0;
  uint16_t reserved2 = 0;
  uint32_t data_pos = 0;
  uint32_t header_size = 0;
  int32_t width = 0;    // negative: mirrored columns
  int32_t height = 0;   // positive: rows stored bottom-up
  uint16_t planes = 0;
  uint16_t bit_count = 0;
  uint32_t compression = 0;  // 0 (BI_RGB) or 3 (BI_BITFIELDS)
  uint32_t size_image_for_compression = 0;
  int32_t x_pixels_per_meter = 0;
  int32_t y_pixels_per_meter = 0;
  uint32_t colors_used = 0;
  uint32_t colors_important = 0;
};

// colour masks that follow the info header of 32-bit images, 84 bytes on disk
struct BMPColorHeader {
  uint32_t red_mask = 0x00ff0000;
  uint32_t green_mask = 0x0000ff00;
  uint32_t blue_mask = 0x000000ff;
  uint32_t alpha_mask = 0xff000000;
  uint32_t color_space = 0x73524742;  // "sRGB"
  uint32_t unused[16] = { 0 };
};

// 64-byte header of a .myyuv file
struct YUVHeader {
  uint8_t type[2] = { 'Y', 'U' };
  uint32_t fourcc_format = 0;
  uint32_t data_size = 0;               // bytes of `data` (the payload when compressed)
  uint16_t compression = 0;             // YUV::Compressions
  uint32_t compression_params_size = 0;
  uint32_t compression_params_pos = 0;
  uint32_t width = 0;
  uint32_t height = 0;
  uint32_t data_pos = 0;
  uint8_t unused[32] = { 0 };
};
#pragma pack(pop)

// An uncompressed BMP image held in host memory.  `data` is owned (new[] / delete[]).
class BMP {
public:
  BMPHeader header;
  BMPColorHeader color_header;
  uint8_t* data = nullptr;
public:
  BMP() {}
  explicit BMP(const std::string& path);
  BMP(const BMP& bmp);
  BMP& operator=(const BMP& bmp);
  BMP(BMP&& bmp) noexcept;
  BMP& operator=(BMP&& bmp) noexcept;
  ~BMP();

  uint32_t trueWidth() const noexcept;    // |width|
  uint32_t trueHeight() const noexcept;   // |height|
  uint32_t imageSize() const noexcept;    // bytes of pixel data
  uint8_t* colorData() const;             // new[] copy with the origin at the top-left corner
  uint8_t* colorDataFlipped() const;      // new[] copy with the origin at the bottom-left corner
  bool isValid() const noexcept;
  bool isValidHeader() const noexcept;
  void load(const std::string& path);     // strong exception guarantee
  void dump(const std::string& path) const;
};

// A YUV image, raw or compressed.  `data` and `compression_params` are owned (new[] / delete[]).
class YUV {
public:
  YUVHeader header;
  uint8_t* compression_params = nullptr;
  uint8_t* data = nullptr;
public:
  enum class FormatGroup { UNKNOWN = 0, PACKED, PLANAR, SEMI_PLANAR };

  using FourccFormat = uint32_t;
  struct FourccFormats {
    static constexpr const FourccFormat UNKNOWN = 0;
    static constexpr const FourccFormat IYUV = 0x56555949;
  };

  using Compression = uint16_t;
  struct Compressions {
    static constexpr const Compression NONE = 0;
    static constexpr const Compression DCT = 1;
  };

  static constexpr const uint32_t max_planes = 4;
  static constexpr const uint8_t no_plane = 0xff;

  // ---- registries (public and mutable, exactly as in the reference: this is its plug-in boundary) ----
  static std::unordered_map<FourccFormat, FormatGroup> yuv_format_group_map;
  static std::unordered_map<FourccFormat, std::array<uint8_t, max_planes>> yuv_order_planes_map;
  static std::unordered_map<FourccFormat, std::array<uint32_t, 2>> yuv_resolution_fraction_map;
  // BMP -> YUV converters.  [IYUV] runs xrgb_to_iyuv_kernel (myyuvb_xrgb_to_iyuv).
  static std::unordered_map<FourccFormat, std::function<YUV(const BMP&)>> bmp_to_yuv_map;
  // compressors (params, params_size).  [DCT][IYUV] runs dct_compress_kernel (myyuvb_dct_compress).
  static std::unordered_map<Compression, std::unordered_map<FourccFormat, std::function<YUV(const YUV&, const void*, uint32_t)>>> compress_map;
  // decompressors.  [DCT][IYUV] runs dct_decompress_kernel (myyuvb_dct_decompress).
  static std::unordered_map<Compression, std::unordered_map<FourccFormat, std::function<YUV(const YUV&)>>> decompress_map;
  static std::unordered_map<FourccFormat, std::function<std::array<uint8_t, max_planes>(const YUV&, uint32_t, uint32_t)>> yuv_get_pixel_map;

  YUV() {}
  explicit YUV(const std::string& path);
  explicit YUV(const BMP& bmp, FourccFormat format);
  YUV(const YUV& yuv);
  YUV& operator=(const YUV& yuv);
  YUV(YUV&& yuv) noexcept;
  YUV& operator=(YUV&& yuv) noexcept;
  ~YUV();

  bool isValid() const noexcept;
  bool isValidHeader() const noexcept;
  static bool isImplementedFormat(FourccFormat format, Compression compression = Compressions::NONE) noexcept;
  FourccFormat getFourccFormat() const noexcept;
  Compression getCompression() const noexcept;
  uint32_t getWidth() const noexcept;
  uint32_t getHeight() const noexcept;
  uint32_t getDataSize() const noexcept;
  std::array<uint32_t, 2> getResolutionFraction() const;
  std::array<uint32_t, 2> getWidthHeightChannel(uint8_t channel) const;
  std::array<uint32_t, max_planes> getFormatSizeBits() const;
  std::array<uint8_t, max_planes> getYUVPlanesOrder() const;
  uint32_t getImageSize() const;
  std::array<const uint8_t*, max_planes> getYUVPlanes() const;
  std::array<uint8_t*, max_planes> getYUVPlanes();
  FormatGroup getFormatGroup() const noexcept;
  static FormatGroup getFormatGroup(FourccFormat format) noexcept;
  std::array<uint8_t, max_planes> getPixel(uint32_t x, uint32_t y) const;
  YUV compress(Compression compression, const void* params, uint32_t params_size) const;
  YUV decompress() const;
  bool isCompressed() const noexcept;
  void load(const std::string& path);
  void load(const BMP& bmp, FourccFormat format);
  void dump(const std::string& path) const;
};

}  // namespace myyuv
