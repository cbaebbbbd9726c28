// decompress_b200/csrc/status_strings.h — the reference's error vocabulary (lib/de.ml:702-730, 1557-1567;
// lib/zl.ml:177-183; lib/gz.ml:285-296; lib/lzo.ml:4-12) for the DCMP_* status codes of include/dcmp.h.
#pragma once
#include "../../include/dcmp.h"

namespace dcmp {
inline const char *status_string(int s) {
  switch (s) {
  case DCMP_OK: return "Ok";
  case DCMP_UNEXPECTED_END_OF_INPUT: return "Unexpected end of input";
  case DCMP_UNEXPECTED_END_OF_OUTPUT: return "Unexpected end of output";
  case DCMP_INVALID_KIND_OF_BLOCK: return "Invalid kind of block";
  case DCMP_INVALID_DICTIONARY: return "Invalid dictionary";
  case DCMP_INVALID_COMPLEMENT_OF_LENGTH: return "Invalid complement of length";
  case DCMP_INVALID_DISTANCE: return "Invalid distance";
  case DCMP_INVALID_DISTANCE_CODE: return "Invalid distance code";
  case DCMP_REF_EXCEPTION: return "Reference raises an exception (index out of bounds)";
  case DCMP_INVALID_HEADER: return "Invalid header";
  case DCMP_INVALID_CHECKSUM: return "Invalid checksum";
  case DCMP_INVALID_GZIP_HEADER: return "Invalid GZip header";
  case DCMP_INVALID_GZIP_HEADER_CHECKSUM: return "Invalid GZip header checksum";
  case DCMP_INVALID_INPUT_SIZE: return "Invalid input size";
  case DCMP_LZO_MALFORMED: return "Malformed";
  case DCMP_LZO_INVALID_ARGUMENT: return "Input is malformed or output is not large enough";
  case DCMP_LZO_INVALID_DICTIONARY: return "Invalid dictionary";
  case DCMP_INVALID_COMPRESSION_LEVEL: return "Invalid compression level";
  default: return "?";
  }
}
} // namespace dcmp
