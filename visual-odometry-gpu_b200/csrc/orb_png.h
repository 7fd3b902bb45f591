// orb_png.h -- host-side PNG -> 8-bit gray decoder of the frame ingest.
// Replaces cv::imread(path, cv::IMREAD_GRAYSCALE) of the reference's VO loops (src/feature_matching.cpp:55,59;
// src/feature_tracking.cpp:56,196) for the files the reference reads (KITTI odometry: 8-bit gray, non-interlaced PNG).
// Hand-written inflate (RFC 1951) + zlib framing (RFC 1950) + PNG unfilter; rows land directly in the caller's buffer
// (the pinned staging area of the wave pipeline), so a decoded frame is never copied on the host.
#pragma once
#include <cstddef>
#include <cstdint>

namespace orbpng {

struct Info {
  int width = 0, height = 0;
  int bit_depth = 0;    // 8 or 16
  int color_type = 0;   // 0 gray, 2 RGB, 4 gray+alpha, 6 RGBA
  int channels = 0;
};

// Both return nullptr on success or a static message describing the failure.
const char* read_info(const uint8_t* file, size_t n, Info* info);
// Scratch memory is owned by the caller's Scratch object so that worker threads reuse their allocations.
struct Scratch {
  uint8_t* buf = nullptr; size_t cap = 0;
  ~Scratch();
  uint8_t* need(size_t n);
};
const char* decode_gray8(const uint8_t* file, size_t n, uint8_t* dst, size_t pitch, int expect_w, int expect_h, Scratch* scratch);

// Device-decode path: checks the framing (signature, IHDR, zlib header) and concatenates the IDAT payloads -- the whole
// zlib stream, header included -- into dst (16 zero bytes appended).  *deflate_bytes = length of the raw deflate stream that
// starts at dst + 2 (the Adler-32 trailer behind it is not counted).  With `crcs`, the IDAT chunk checksums are not verified
// here but listed (offset into dst, length, stored CRC) for the device to verify; files with more than crc_cap chunks are
// verified on the host instead (*n_crcs = 0).
struct ChunkCrc { uint32_t offset, len, crc; };
const char* extract_deflate(const uint8_t* file, size_t n, Info* info, uint8_t* dst, size_t dst_cap, size_t* deflate_bytes,
                            ChunkCrc* crcs = nullptr, int crc_cap = 0, int* n_crcs = nullptr);

// raw pieces, exported for the tests
const char* inflate_zlib(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out, size_t* produced, bool verify_adler);
uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0);
uint32_t adler32(const uint8_t* p, size_t n, uint32_t adler = 1);

}  // namespace orbpng
