// bmp_host.cpp -- host side of myyuv::BMP for the drop-in library (behaviour of myyuv_lib/myyuv_bmp.cpp,
// written from its observable semantics: file layout, validity rules, orientation handling, exception text).
// Nothing here is hot: the row flip of colorData() is folded into the converter kernel's addressing instead.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <utility>

#include "../../include/myyuv.hpp"

namespace myyuv {

namespace {
// pixel bytes of an image described by a header
uint32_t pixel_bytes(const BMPHeader& h) {
  return static_cast<uint32_t>(std::abs(h.width)) * static_cast<uint32_t>(std::abs(h.height)) * h.bit_count / 8;
}
}  // namespace

BMP::BMP(const std::string& path) : BMP() { load(path); }

BMP::BMP(const BMP& other) : BMP() { *this = other; }

BMP& BMP::operator=(const BMP& other) {
  if (this == &other) return *this;
  const uint32_t need = other.imageSize();
  if (other.data != nullptr) {
    // keep the current buffer when it is large enough (myyuv_bmp.cpp:23-29), otherwise allocate first so that a
    // failed allocation leaves *this untouched
    if (data == nullptr || need > imageSize()) {
      uint8_t* fresh = new uint8_t[need];
      delete[] data;
      data = fresh;
    }
    std::memcpy(data, other.data, need);
  } else {
    delete[] data;
    data = nullptr;
  }
  header = other.header;
  color_header = other.color_header;
  return *this;
}

BMP::BMP(BMP&& other) noexcept : BMP() { *this = std::move(other); }

BMP& BMP::operator=(BMP&& other) noexcept {
  std::swap(header, other.header);
  std::swap(color_header, other.color_header);
  std::swap(data, other.data);
  return *this;
}

BMP::~BMP() { delete[] data; }

uint32_t BMP::trueWidth() const noexcept { return static_cast<uint32_t>(std::abs(header.width)); }
uint32_t BMP::trueHeight() const noexcept { return static_cast<uint32_t>(std::abs(header.height)); }
uint32_t BMP::imageSize() const noexcept { return pixel_bytes(header); }

// Orientation rules of myyuv_bmp.cpp:80-103: (w>0,h<0) rows already top-down; (w>0,h>0) rows bottom-up;
// (w<0,h>0) whole pixel sequence reversed; anything else is rejected.
uint8_t* BMP::colorData() const {
  if (!isValid()) throw std::runtime_error("BMP data is invalid");
  const uint32_t total = imageSize();
  const uint32_t bpp = header.bit_count / 8;
  std::unique_ptr<uint8_t[]> out(new uint8_t[total]);
  if (header.width > 0 && header.height < 0) {
    std::memcpy(out.get(), data, total);
  } else if (header.width > 0 && header.height > 0) {
    const size_t row = static_cast<size_t>(bpp) * header.width;
    for (int32_t r = 0; r < header.height; r++) std::memcpy(out.get() + row * r, data + row * (header.height - 1 - r), row);
  } else if (header.width < 0 && header.height > 0) {
    const uint32_t npx = total / bpp;
    for (uint32_t p = 0; p < npx; p++) std::memcpy(out.get() + static_cast<size_t>(p) * bpp, data + static_cast<size_t>(npx - 1 - p) * bpp, bpp);
  } else {
    throw std::runtime_error("Unaccounted width and height sign");
  }
  return out.release();
}

uint8_t* BMP::colorDataFlipped() const {
  if (!isValid()) throw std::runtime_error("BMP data is invalid");
  const uint32_t total = imageSize();
  const uint32_t bpp = header.bit_count / 8;
  std::unique_ptr<uint8_t[]> out(new uint8_t[total]);
  if (header.width > 0 && header.height > 0) {
    std::memcpy(out.get(), data, total);
  } else if (header.width > 0 && header.height < 0) {
    // the reference iterates `uint32_t i < header.height` with a negative height (myyuv_bmp.cpp:115), i.e. the
    // comparison is done in unsigned arithmetic and the loop copies |2^32 + height| rows: undefined behaviour.
    // Here the rows are simply flipped.
    const int32_t rows = -header.height;
    const size_t row = static_cast<size_t>(bpp) * header.width;
    for (int32_t r = 0; r < rows; r++) std::memcpy(out.get() + row * r, data + row * (rows - 1 - r), row);
  } else {
    throw std::runtime_error("Unaccounted width and height sign");
  }
  return out.release();
}

bool BMP::isValid() const noexcept { return data != nullptr && isValidHeader(); }

// myyuv_bmp.cpp:127-139: no row padding (width % 4), uncompressed, true colour, standard XRGB/ARGB masks, sRGB
bool BMP::isValidHeader() const noexcept {
  if (header.type[0] != 'B' || header.type[1] != 'M') return false;
  if (header.width % 4 != 0 || header.bit_count == 0 || header.header_size == 0) return false;
  if (header.compression != 0 && header.compression != 3) return false;
  if (header.colors_used != 0 || header.colors_important != 0) return false;
  if (color_header.red_mask != 0x00ff0000 || color_header.green_mask != 0x0000ff00 || color_header.blue_mask != 0x000000ff) return false;
  if (color_header.alpha_mask != 0xff000000 && color_header.alpha_mask != 0) return false;
  return color_header.color_space == 0x73524742;
}

void BMP::load(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("Error opening file to read " + path);
  BMP tmp;
  in.read(reinterpret_cast<char*>(&tmp.header), sizeof(tmp.header));
  const bool has_masks = tmp.header.bit_count == 32;
  if (has_masks) in.read(reinterpret_cast<char*>(&tmp.color_header), sizeof(tmp.color_header));
  in.seekg(tmp.header.data_pos, in.beg);
  // positions are normalised to this library's own dump layout (myyuv_bmp.cpp:152-158)
  tmp.header.data_pos = sizeof(BMPHeader) + (has_masks ? sizeof(BMPColorHeader) : 0);
  const uint32_t total = tmp.imageSize();
  tmp.header.file_size = tmp.header.data_pos + total;
  if (!tmp.isValidHeader()) throw std::runtime_error("Error bad header " + path);
  tmp.data = new uint8_t[total];
  in.read(reinterpret_cast<char*>(tmp.data), total);
  *this = std::move(tmp);
}

void BMP::dump(const std::string& path) const {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("Error opening file to write " + path);
  out.write(reinterpret_cast<const char*>(&header), sizeof(header));
  if (header.bit_count == 32) out.write(reinterpret_cast<const char*>(&color_header), sizeof(color_header));
  out.write(reinterpret_cast<const char*>(data), imageSize());
}

}  // namespace myyuv
