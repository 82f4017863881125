// bmp_host.cpp -- host side of myyuv::BMP for the drop-in library.  Written from the observable semantics of the reference's
// class (myyuv_lib/myyuv_bmp.cpp: file layout, validity rules, orientation handling, exception texts) and held to them by
// tests/test_class_diff.py, which runs one program against this library and against the unmodified reference library and
// compares what they do with well-formed and malformed files.  Nothing here is hot: the converter kernels read the rows as the
// file stores them and fold the orientation into their addressing, so colorData() is only for callers that ask for it.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <utility>

#include "../../include/myyuv.hpp"

namespace myyuv {

namespace {

using Bytes = std::unique_ptr<uint8_t[]>;

// How the stored pixel sequence has to be rearranged for a requested origin.
enum class Reorder { kNone, kRows, kPixels, kUndefined };

// The file stores rows bottom-up when height > 0 and top-down when height < 0; a negative width mirrors the columns, which
// together with bottom-up rows is the whole sequence back to front.  Same cases as myyuv_bmp.cpp:87-101 (top-left origin)
// and :110-121 (bottom-left origin); every other sign combination is rejected there and here.
Reorder reorder_for(int32_t width, int32_t height, bool top_left_origin) {
  if (width > 0) {
    const bool stored_top_down = height < 0;
    if (height == 0) return Reorder::kUndefined;
    return stored_top_down == top_left_origin ? Reorder::kNone : Reorder::kRows;
  }
  if (width < 0 && height > 0 && top_left_origin) return Reorder::kPixels;
  return Reorder::kUndefined;
}

// total: bytes of the image (imageSize(): it is what callers index, also for bit counts that are no whole number of bytes,
// where bytes_per_pixel rounds down and the row / pixel loops move fewer bytes than that -- as in the reference)
Bytes rearranged(const uint8_t* src, size_t total, uint32_t columns, uint32_t rows, uint32_t bytes_per_pixel, Reorder how) {
  const size_t row_bytes = static_cast<size_t>(columns) * bytes_per_pixel;
  Bytes out(new uint8_t[total]);
  switch (how) {
    case Reorder::kNone:
      std::memcpy(out.get(), src, total);
      break;
    case Reorder::kRows:
      for (uint32_t r = 0; r < rows; r++) std::memcpy(out.get() + row_bytes * r, src + row_bytes * (rows - 1 - r), row_bytes);
      break;
    case Reorder::kPixels: {
      const size_t pixels = static_cast<size_t>(columns) * rows;
      for (size_t p = 0; p < pixels; p++) std::memcpy(out.get() + p * bytes_per_pixel, src + (pixels - 1 - p) * bytes_per_pixel, bytes_per_pixel);
      break;
    }
    case Reorder::kUndefined:
      throw std::runtime_error("Unaccounted width and height sign");
  }
  return out;
}

template <class T>
void read_pod(std::ifstream& in, T& v) { in.read(reinterpret_cast<char*>(&v), sizeof(T)); }
template <class T>
void write_pod(std::ofstream& out, const T& v) { out.write(reinterpret_cast<const char*>(&v), sizeof(T)); }

}  // namespace

BMP::BMP(const std::string& path) : BMP() { load(path); }

BMP::BMP(const BMP& other) : BMP() { *this = other; }

BMP::BMP(BMP&& other) noexcept : BMP() { *this = std::move(other); }

BMP::~BMP() { delete[] data; }

// Copy: the new pixels are in hand before anything of *this changes (an allocation failure leaves it as it was).
BMP& BMP::operator=(const BMP& other) {
  if (this == &other) return *this;
  Bytes fresh;
  if (other.data != nullptr) {
    const uint32_t bytes = other.imageSize();
    fresh.reset(new uint8_t[bytes]);
    std::memcpy(fresh.get(), other.data, bytes);
  }
  delete[] data;
  data = fresh.release();
  header = other.header;
  color_header = other.color_header;
  return *this;
}

// Move: the two objects trade places; the source releases what it received when it goes out of scope.
BMP& BMP::operator=(BMP&& other) noexcept {
  std::swap(data, other.data);
  std::swap(color_header, other.color_header);
  std::swap(header, other.header);
  return *this;
}

uint32_t BMP::trueWidth() const noexcept { return static_cast<uint32_t>(std::abs(header.width)); }

uint32_t BMP::trueHeight() const noexcept { return static_cast<uint32_t>(std::abs(header.height)); }

uint32_t BMP::imageSize() const noexcept { return trueWidth() * trueHeight() * header.bit_count / 8; }

uint8_t* BMP::colorData() const {
  if (!isValid()) throw std::runtime_error("BMP data is invalid");
  return rearranged(data, imageSize(), trueWidth(), trueHeight(), header.bit_count / 8u, reorder_for(header.width, header.height, true)).release();
}

// The reference's loop for (width > 0, height < 0) compares an unsigned counter with the negative height
// (myyuv_bmp.cpp:115), which is undefined behaviour; here that case simply gets its rows reversed.
uint8_t* BMP::colorDataFlipped() const {
  if (!isValid()) throw std::runtime_error("BMP data is invalid");
  return rearranged(data, imageSize(), trueWidth(), trueHeight(), header.bit_count / 8u, reorder_for(header.width, header.height, false)).release();
}

bool BMP::isValid() const noexcept { return isValidHeader() && data != nullptr; }

// What the reference accepts (myyuv_bmp.cpp:127-139): a "BM" file without row padding, uncompressed true colour, the
// standard XRGB / ARGB channel masks, sRGB.
bool BMP::isValidHeader() const noexcept {
  const BMPHeader& h = header;
  const BMPColorHeader& c = color_header;
  const bool signature = h.type[0] == 'B' && h.type[1] == 'M';
  const bool geometry = h.width % 4 == 0 && h.bit_count != 0 && h.header_size != 0;
  const bool plain = (h.compression == 0 || h.compression == 3) && h.colors_used == 0 && h.colors_important == 0;
  const bool channels = c.red_mask == 0x00ff0000u && c.green_mask == 0x0000ff00u && c.blue_mask == 0x000000ffu &&
                        (c.alpha_mask == 0xff000000u || c.alpha_mask == 0u);
  return signature && geometry && plain && channels && c.color_space == 0x73524742u;
}

// The object is only replaced once the whole file has been taken in.  Positions are rewritten to the layout dump() produces
// (myyuv_bmp.cpp:152-158): headers back to back, pixels right behind them.
void BMP::load(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("Error opening file to read " + path);
  BMP incoming;
  read_pod(in, incoming.header);
  const bool with_masks = incoming.header.bit_count == 32;
  if (with_masks) read_pod(in, incoming.color_header);
  in.seekg(incoming.header.data_pos, std::ios::beg);
  const uint32_t pixels = incoming.imageSize();
  incoming.header.data_pos = static_cast<uint32_t>(sizeof(BMPHeader) + (with_masks ? sizeof(BMPColorHeader) : 0));
  incoming.header.file_size = incoming.header.data_pos + pixels;
  if (!incoming.isValidHeader()) throw std::runtime_error("Error bad header " + path);
  incoming.data = new uint8_t[pixels];
  in.read(reinterpret_cast<char*>(incoming.data), pixels);
  *this = std::move(incoming);
}

void BMP::dump(const std::string& path) const {
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("Error opening file to write " + path);
  write_pod(out, header);
  if (header.bit_count == 32) write_pod(out, color_header);
  out.write(reinterpret_cast<const char*>(data), imageSize());
}

}  // namespace myyuv
