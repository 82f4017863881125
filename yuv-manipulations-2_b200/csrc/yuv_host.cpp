// yuv_host.cpp -- host side of myyuv::YUV for the drop-in library: file I/O, header rules, accessors, dispatch
// through the public registries (behaviour of myyuv_lib/myyuv_yuv.cpp:182-536), and the three registry entries
// of the hot path (myyuv_yuv.cpp:88-160), which here call the CUDA library through the C ABI of myyuvb200.h.
// There is no CPU implementation of the codec: if no CUDA device is usable the entries throw.
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <utility>

#include "../../include/myyuv.hpp"
#include "../../include/myyuvb200.h"

namespace myyuv {

namespace {

// One lazily created context for the class API (the reference's free functions are stateless; a context is
// not thread-safe, so calls are serialised).
std::mutex g_ctx_mutex;
myyuvb_ctx* g_ctx = nullptr;

struct CtxLock {
  std::lock_guard<std::mutex> lock;
  CtxLock() : lock(g_ctx_mutex) {
    if (!g_ctx && myyuvb_ctx_create(0, nullptr, &g_ctx) != MYYUVB_OK) throw std::runtime_error(myyuvb_last_error());
  }
  myyuvb_ctx* get() const { return g_ctx; }
};

void check(int rc) {
  if (rc != MYYUVB_OK) throw std::runtime_error(myyuvb_last_error());
}

template <typename Map, typename Key>
bool has(const Map& m, const Key& k) {
  return m.find(k) != m.end();
}

// ---- registry entry: BMP (XRGB8888) -> IYUV, replaces the lambda at myyuv_yuv.cpp:89-127 ----
YUV bmp_to_iyuv(const BMP& bmp) {
  if (!bmp.isValid()) throw std::runtime_error("BMP data is invalid");  // what colorData() would throw (:99)
  // assert(bit_count == 32) at :92 ("TODO: test 24"); the Release build (NDEBUG) goes on and reads 24-bit files as B,G,R
  // triplets through the same pixel_bits / 8 addressing (:34-41), which is what the 24-bit entry point does
  if (bmp.header.bit_count != 32 && bmp.header.bit_count != 24)
    throw std::runtime_error("Error. only 24-bit and 32-bit BMP are supported");
  const auto convert = bmp.header.bit_count == 32 ? myyuvb_xrgb_to_iyuv : myyuvb_bgr24_to_iyuv;
  const uint32_t w = bmp.trueWidth(), h = bmp.trueHeight();
  YUV out;
  out.header.fourcc_format = YUV::FourccFormats::IYUV;
  out.header.width = w;
  out.header.height = h;
  out.header.data_size = w * h * 3 / 2;
  out.header.data_pos = sizeof(YUVHeader);
  out.data = new uint8_t[out.header.data_size];
  CtxLock ctx;
  if (bmp.header.width > 0 && bmp.header.height != 0) {
    // rows stay where they are; the kernel reads them bottom-up when height > 0 (myyuv_bmp.cpp:95-98)
    check(convert(ctx.get(), bmp.data, w, h, bmp.header.height > 0 ? 1 : 0, out.data));
  } else if (bmp.header.width < 0 && bmp.header.height > 0) {
    std::unique_ptr<uint8_t[]> px(bmp.colorData());  // reversed pixel order, rare: reorder on the host
    check(convert(ctx.get(), px.get(), w, h, 0, out.data));
  } else {
    throw std::runtime_error("Unaccounted width and height sign");
  }
  return out;
}

// ---- registry entry: DCT compression of IYUV, replaces myyuv_yuv.cpp:132-142 + DCT.cpp:371-430 ----
YUV compress_dct_iyuv(const YUV& src, const void* params, uint32_t params_size) {
  if (params_size != 3) throw std::runtime_error("Error compression: incorrect parameters count. 3 parameters required");
  const uint8_t* q = static_cast<const uint8_t*>(params);
  // two steps, so that YUV::data is allocated once with its final size (the reference's dump(), DCT.cpp:160-173, allocates
  // totalSize() bytes): the payload waits in device memory while its size comes back
  CtxLock ctx;
  uint32_t size = 0;
  check(myyuvb_dct_compress_begin(ctx.get(), src.data, src.header.width, src.header.height, q, &size));
  YUV out;
  out.header = src.header;  // DCT.cpp:390-394
  out.header.compression = YUV::Compressions::DCT;
  out.header.compression_params_size = 3;
  out.header.compression_params_pos = sizeof(YUVHeader);
  out.header.data_pos = sizeof(YUVHeader) + 3;
  out.header.data_size = size;
  out.compression_params = new uint8_t[3]{q[0], q[1], q[2]};
  out.data = new uint8_t[size];
  check(myyuvb_dct_compress_fetch(ctx.get(), out.data, size));
  return out;
}

// ---- registry entry: DCT decompression, replaces myyuv_yuv.cpp:148-158 + DCT.cpp:432-488 ----
YUV decompress_dct_iyuv(const YUV& src) {
  if (src.header.compression_params_size != 3)
    throw std::runtime_error("Error decompression: incorrect parameters count. 3 parameters required");
  YUV out;
  out.header = src.header;  // DCT.cpp:447-453
  out.header.compression = YUV::Compressions::NONE;
  out.header.compression_params_size = 0;
  out.header.compression_params_pos = 0;
  out.header.data_pos = sizeof(YUVHeader);
  out.header.data_size = src.getImageSize();
  out.data = new uint8_t[out.header.data_size];
  CtxLock ctx;
  check(myyuvb_dct_decompress(ctx.get(), src.data, src.header.data_size, src.header.width, src.header.height, src.compression_params, out.data));
  return out;
}

std::array<uint8_t, YUV::max_planes> iyuv_pixel(const YUV& img, uint32_t x, uint32_t y) {  // myyuv_yuv.cpp:163-179
  const uint32_t w = img.getWidth(), h = img.getHeight();
  if (img.isCompressed()) throw std::runtime_error("Cannot get pixel from compressed image. Decompress first.");
  if (x >= w || y >= h) throw std::runtime_error("Image coordinates are out of bounds");
  // The reference's chroma index (x / 2 + y * width / 4, myyuv_yuv.cpp:175) equals (y / 2) * (width / 2) + x / 2 on even rows only; on
  // odd rows it lies width / 4 further on, and on the LAST row it runs past the V plane for x >= width / 2, where the reference
  // reads whatever follows its buffer.  Here, as in get_pixels_kernel, that sample is 0.
  const uint64_t frame = (uint64_t)w * h * 3 / 2;
  const uint64_t c = x / 2 + (uint64_t)(y * w / 4);
  const uint64_t vi = (uint64_t)w * h * 5 / 4 + c;
  std::array<uint8_t, YUV::max_planes> px{0};
  px[0] = img.data[x + (uint64_t)y * w];
  px[1] = img.data[(uint64_t)w * h + c];
  px[2] = vi < frame ? img.data[vi] : (uint8_t)0;
  return px;
}

}  // namespace

static_assert(sizeof(YUVHeader) == 64 && sizeof(BMPHeader) == 54 && sizeof(BMPColorHeader) == 84, "packed headers");

std::unordered_map<YUV::FourccFormat, YUV::FormatGroup> YUV::yuv_format_group_map = {{FourccFormats::IYUV, FormatGroup::PLANAR}};
std::unordered_map<YUV::FourccFormat, std::array<uint8_t, YUV::max_planes>> YUV::yuv_order_planes_map = {{FourccFormats::IYUV, {0, 1, 2, no_plane}}};
std::unordered_map<YUV::FourccFormat, std::array<uint32_t, 2>> YUV::yuv_resolution_fraction_map = {{FourccFormats::IYUV, {2, 2}}};
std::unordered_map<YUV::FourccFormat, std::function<YUV(const BMP&)>> YUV::bmp_to_yuv_map = {{FourccFormats::IYUV, bmp_to_iyuv}};
std::unordered_map<YUV::Compression, std::unordered_map<YUV::FourccFormat, std::function<YUV(const YUV&, const void*, uint32_t)>>> YUV::compress_map = {
    {Compressions::DCT, {{FourccFormats::IYUV, compress_dct_iyuv}}}};
std::unordered_map<YUV::Compression, std::unordered_map<YUV::FourccFormat, std::function<YUV(const YUV&)>>> YUV::decompress_map = {
    {Compressions::DCT, {{FourccFormats::IYUV, decompress_dct_iyuv}}}};
std::unordered_map<YUV::FourccFormat, std::function<std::array<uint8_t, YUV::max_planes>(const YUV&, uint32_t, uint32_t)>> YUV::yuv_get_pixel_map = {
    {FourccFormats::IYUV, iyuv_pixel}};

YUV::YUV(const std::string& path) : YUV() { load(path); }
YUV::YUV(const BMP& bmp, FourccFormat format) : YUV() { load(bmp, format); }
YUV::YUV(const YUV& other) { *this = other; }

YUV& YUV::operator=(const YUV& other) {
  if (this == &other) return *this;
  // allocate everything first: on bad_alloc *this is unchanged (myyuv_yuv.cpp:194-230)
  std::unique_ptr<uint8_t[]> nd, np;
  if (other.data) {
    nd.reset(new uint8_t[other.header.data_size]);
    std::memcpy(nd.get(), other.data, other.header.data_size);
  }
  if (other.compression_params) {
    np.reset(new uint8_t[other.header.compression_params_size]);
    std::memcpy(np.get(), other.compression_params, other.header.compression_params_size);
  }
  delete[] data;
  delete[] compression_params;
  data = nd.release();
  compression_params = np.release();
  header = other.header;
  return *this;
}

YUV::YUV(YUV&& other) noexcept { *this = std::move(other); }

YUV& YUV::operator=(YUV&& other) noexcept {
  std::swap(header, other.header);
  std::swap(compression_params, other.compression_params);
  std::swap(data, other.data);
  return *this;
}

YUV::~YUV() {
  delete[] data;
  delete[] compression_params;
}

bool YUV::isValid() const noexcept {  // myyuv_yuv.cpp:248-254
  if (data == nullptr) return false;
  const bool params_ok = (header.compression_params_size > 0 && compression_params != nullptr) ||
                         (header.compression == Compressions::NONE && compression_params == nullptr) ||
                         (header.compression_params_size == 0 && compression_params == nullptr);
  return params_ok && isValidHeader();
}

bool YUV::isValidHeader() const noexcept {  // myyuv_yuv.cpp:256-262
  return header.type[0] == 'Y' && header.type[1] == 'U' && isImplementedFormat(getFourccFormat(), getCompression()) && header.width > 0 &&
         header.height > 0 && header.data_pos >= sizeof(YUVHeader) + header.compression_params_size && header.data_size > 0;
}

bool YUV::isImplementedFormat(FourccFormat format, Compression compression) noexcept {  // myyuv_yuv.cpp:264-276
  if (!has(bmp_to_yuv_map, format) || !has(yuv_resolution_fraction_map, format)) return false;
  if (compression == Compressions::NONE) return true;
  if (!has(compress_map, compression) || !has(decompress_map, compression)) return false;
  return has(compress_map.at(compression), format) && has(decompress_map.at(compression), format);
}

bool YUV::isCompressed() const noexcept { return getCompression() != Compressions::NONE; }
YUV::FourccFormat YUV::getFourccFormat() const noexcept { return header.fourcc_format; }
YUV::Compression YUV::getCompression() const noexcept { return header.compression; }
uint32_t YUV::getWidth() const noexcept { return header.width; }
uint32_t YUV::getHeight() const noexcept { return header.height; }
uint32_t YUV::getDataSize() const noexcept { return header.data_size; }

std::array<uint32_t, 2> YUV::getResolutionFraction() const {
  if (!isImplementedFormat(getFourccFormat(), Compressions::NONE)) throw std::runtime_error("Error. Unimplemented format.");
  return yuv_resolution_fraction_map.at(getFourccFormat());
}

std::array<uint32_t, 2> YUV::getWidthHeightChannel(uint8_t channel) const {  // myyuv_yuv.cpp:309-325
  const auto order = getYUVPlanesOrder();
  if (order[channel] == no_plane) return {0, 0};
  if (channel == 1 || channel == 2) {
    const auto frac = getResolutionFraction();
    return {header.width / frac[0], header.height / frac[1]};
  }
  return {header.width, header.height};
}

std::array<uint32_t, YUV::max_planes> YUV::getFormatSizeBits() const {  // myyuv_yuv.cpp:327-343
  if (!isImplementedFormat(getFourccFormat(), Compressions::NONE)) throw std::runtime_error("Error. Unimplemented format.");
  const auto frac = getResolutionFraction();
  const auto order = getYUVPlanesOrder();
  const uint32_t sub = frac[0] * frac[1];
  std::array<uint32_t, max_planes> bits = {8, 8 / sub, 8 / sub, 8};
  for (uint32_t i = 0; i < max_planes; i++)
    if (order[i] == no_plane) bits[i] = 0;
  return bits;
}

std::array<uint8_t, YUV::max_planes> YUV::getYUVPlanesOrder() const {
  if (!isImplementedFormat(getFourccFormat(), Compressions::NONE)) throw std::runtime_error("Error. Unimplemented format.");
  if (!has(yuv_order_planes_map, getFourccFormat())) throw std::runtime_error("Error. Planar type unimplemented (?)");
  return yuv_order_planes_map.at(getFourccFormat());
}

uint32_t YUV::getImageSize() const {  // myyuv_yuv.cpp:374-381 (uint32 arithmetic on purpose)
  const auto bits = getFormatSizeBits();
  uint32_t total = 0;
  for (uint32_t i = 0; i < max_planes; i++) total += header.width * header.height * bits[i] / 8;
  return total;
}

std::array<const uint8_t*, YUV::max_planes> YUV::getYUVPlanes() const {  // myyuv_yuv.cpp:383-423
  if (!has(yuv_order_planes_map, getFourccFormat())) throw std::runtime_error("Error. Planar type unimplemented (?)");
  const auto order = getYUVPlanesOrder();
  const auto bits = getFormatSizeBits();
  const FormatGroup group = getFormatGroup();
  std::array<const uint8_t*, max_planes> planes = {nullptr, nullptr, nullptr, nullptr};
  planes[order[0]] = data;
  uint8_t back = 1;  // distance to the previous plane that exists
  for (uint8_t i = 1; i < max_planes; i++) {
    const uint8_t cur = order[i];
    if (cur == no_plane) {
      back++;
      continue;
    }
    const uint8_t prev = order[i - back];
    planes[cur] = group == FormatGroup::PACKED ? data : planes[prev] + header.width * header.height * bits[prev] / 8;
  }
  for (uint32_t i = 0; i < max_planes; i++) {
    const uint32_t cur = order[i];
    if (cur != no_plane && bits[cur] == 0) planes[cur] = nullptr;
  }
  if (group == FormatGroup::SEMI_PLANAR) {
    if (planes[1] != nullptr) planes[2] = planes[1];
    else if (planes[2] != nullptr) planes[1] = planes[2];
  }
  return planes;
}

std::array<uint8_t*, YUV::max_planes> YUV::getYUVPlanes() {
  const auto c = static_cast<const YUV*>(this)->getYUVPlanes();
  std::array<uint8_t*, max_planes> m{};
  for (uint32_t i = 0; i < max_planes; i++) m[i] = const_cast<uint8_t*>(c[i]);
  return m;
}

YUV::FormatGroup YUV::getFormatGroup() const noexcept { return getFormatGroup(getFourccFormat()); }

YUV::FormatGroup YUV::getFormatGroup(FourccFormat format) noexcept {
  const auto it = yuv_format_group_map.find(format);
  return it == yuv_format_group_map.end() ? FormatGroup::UNKNOWN : it->second;
}

std::array<uint8_t, YUV::max_planes> YUV::getPixel(uint32_t x, uint32_t y) const {  // myyuv_yuv.cpp:441-452
  if (!has(yuv_get_pixel_map, getFourccFormat())) throw std::runtime_error("Unimplemented");
  if (isCompressed()) throw std::runtime_error("Cannot get pixel from compressed image. Decompress first.");
  if (x >= getWidth() || y >= getHeight()) throw std::runtime_error("Image coordinates are out of bounds");
  return yuv_get_pixel_map.at(getFourccFormat())(*this, x, y);
}

YUV YUV::compress(Compression compression, const void* params, uint32_t params_size) const {  // myyuv_yuv.cpp:454-467
  if (getCompression() != Compressions::NONE) throw std::runtime_error("Error already compressed");
  if (!has(compress_map, compression)) throw std::runtime_error("Error this compression is unimplemented");
  const auto& by_format = compress_map.at(compression);
  if (!has(by_format, getFourccFormat())) throw std::runtime_error("Error compression for this format is unimplemented");
  return by_format.at(getFourccFormat())(*this, params, params_size);
}

YUV YUV::decompress() const {  // myyuv_yuv.cpp:469-483
  const Compression compression = getCompression();
  if (compression == Compressions::NONE) return *this;
  if (!has(decompress_map, compression)) throw std::runtime_error("Error this decompression is unimplemented");
  const auto& by_format = decompress_map.at(compression);
  if (!has(by_format, getFourccFormat())) throw std::runtime_error("Error decompression for this format is unimplemented");
  return by_format.at(getFourccFormat())(*this);
}

void YUV::load(const std::string& path) {  // myyuv_yuv.cpp:485-510
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("Error opening file to read " + path);
  YUV tmp;
  in.read(reinterpret_cast<char*>(&tmp.header), sizeof(tmp.header));
  if (!tmp.isValidHeader()) throw std::runtime_error("Error bad header " + path);
  if (tmp.header.compression_params_size > 0) {
    in.seekg(tmp.header.compression_params_pos, in.beg);
    tmp.compression_params = new uint8_t[tmp.header.compression_params_size];
    in.read(reinterpret_cast<char*>(tmp.compression_params), tmp.header.compression_params_size);
  }
  in.seekg(tmp.header.data_pos, in.beg);
  tmp.header.compression_params_pos = sizeof(YUVHeader);
  tmp.header.data_pos = tmp.header.compression_params_pos + tmp.header.compression_params_size;
  if (tmp.getCompression() == Compressions::NONE) tmp.header.data_size = tmp.getImageSize();
  tmp.data = new uint8_t[tmp.header.data_size];
  in.read(reinterpret_cast<char*>(tmp.data), tmp.header.data_size);
  *this = std::move(tmp);
}

void YUV::load(const BMP& bmp, FourccFormat format) {  // myyuv_yuv.cpp:512-523
  if (!bmp.isValid()) throw std::runtime_error("BMP is invalid");
  if (!has(bmp_to_yuv_map, format)) throw std::runtime_error("Incorrect format");
  YUV tmp = bmp_to_yuv_map.at(format)(bmp);
  *this = std::move(tmp);
}

void YUV::dump(const std::string& path) const {  // myyuv_yuv.cpp:525-536
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("Error opening file to write " + path);
  out.write(reinterpret_cast<const char*>(&header), sizeof(header));
  if (compression_params != nullptr) out.write(reinterpret_cast<const char*>(compression_params), header.compression_params_size);
  out.write(reinterpret_cast<const char*>(data), header.data_size);
}

}  // namespace myyuv
