// Internal header of the library's DEFLATE decoder (csrc/rsb_inflate.cpp); see there.
#ifndef RSB_INFLATE_H
#define RSB_INFLATE_H

#include <stddef.h>
#include <stdint.h>

namespace rsb {

// The decoder reads and writes a little past the ends it is given instead of checking every access:
static constexpr size_t kInflateInPad = 64;   // readable bytes required after in[in_len)
static constexpr size_t kInflateOutPad = 64;  // writable bytes required after out[out_len)

// Raw DEFLATE stream -> exactly out_len bytes. 0 on success (*consumed = input bytes used); negative: -1 unsupported host,
// -2 input or output exhausted, -3 invalid stream, -4 stream ended before out_len bytes.
int rsb_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, size_t* consumed);
// zlib-wrapped stream (RFC 1950) with the same padding requirements; -5: Adler-32 mismatch.
int rsb_inflate_zlib_padded(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
uint32_t rsb_adler32(const uint8_t* p, size_t n);

}  // namespace rsb

#endif
