// decompress_b200/csrc/inflate_stream.h — state of one resumable DEFLATE stream (the streaming half of the C ABI:
// dcmp_inf_* / dcmp_zl_inf_* / dcmp_gz_inf_*, csrc/stream.cu), shared by the kernel (inflate_serial.cu) and the
// host state machine.  It is what De.Inf.decoder (lib/de.ml:651-672) carries between two calls of `decode`:
// the bit cursor, the block being decoded and its code lengths, the 32 KiB window (a device ring) and how much
// of it is filled.
#pragma once
#include <stdint.h>

namespace dcmp {

enum : int32_t { SP_HEADER = 0, SP_CODED = 1, SP_STORED = 2, SP_DONE = 3 };            // InfStreamState.phase
enum : int32_t { SR_AWAIT = 0, SR_FULL = 1, SR_END = 2, SR_ERROR = 3 };                // InfStreamState.result

struct InfStreamState {
  int64_t bitpos;      // next bit of the input buffer to decode (a block header or a token boundary)
  int64_t total_out;   // bytes produced since the stream began (ring write position = total_out & mask)
  int64_t produced;    // out: bytes this launch produced
  int32_t phase;       // SP_*
  int32_t last;        // BFINAL of the block being decoded
  int32_t btype;       // SP_CODED: 1 fixed, 2 dynamic
  int32_t hlit, hdist; // SP_CODED, dynamic: the code lengths kept in lens[]
  int32_t stored_rem;  // SP_STORED: payload bytes still to copy
  int32_t result;      // out: SR_*
  int32_t status;      // out: DCMP_* of SR_ERROR
  uint8_t lens[320];
};

constexpr int64_t STREAM_RING = 1 << 20;            // device ring: the window (32 KiB) + what one launch may produce
constexpr int64_t STREAM_MAX_OUT = STREAM_RING - 32768 - 512;

} // namespace dcmp
