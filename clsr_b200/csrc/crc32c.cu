// CRC-32C (Castagnoli), slice-by-8, host only.  TensorFlow's tensor-bundle reader verifies the masked
// crc32c of every tensor it restores (BundleEntryProto.crc32c), so the checkpoint writer
// (clsr_b200/tf_bundle.py; reference: tf.train.Saver at base_model.py:58) must store it; a pure-Python
// byte loop over a multi-GB table is too slow, hence this entry point.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/clsr_b200.h"

namespace {

struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};

}  // namespace

extern "C" uint32_t clsr_crc32c(const void* data, uint64_t bytes, uint32_t crc) {
  static const Crc32cTables T;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  crc ^= 0xFFFFFFFFu;
  while (bytes && (reinterpret_cast<uintptr_t>(p) & 7)) {
    crc = T.t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    --bytes;
  }
  while (bytes >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= crc;
    crc = T.t[7][w & 0xFF] ^ T.t[6][(w >> 8) & 0xFF] ^ T.t[5][(w >> 16) & 0xFF] ^ T.t[4][(w >> 24) & 0xFF] ^
          T.t[3][(w >> 32) & 0xFF] ^ T.t[2][(w >> 40) & 0xFF] ^ T.t[1][(w >> 48) & 0xFF] ^ T.t[0][(w >> 56) & 0xFF];
    p += 8;
    bytes -= 8;
  }
  while (bytes--) crc = T.t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return crc ^ 0xFFFFFFFFu;
}
