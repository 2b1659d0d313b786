// 2-bit packing of query bases on the host (see host_pack.cpp).
#pragma once
#include <cstdint>

namespace cls {

// Packs `len` ASCII bases into ceil(len / 16) words at dst; false if any byte is not A/C/G/T/a/c/g/t.
bool pack_read(const uint8_t *s, uint32_t len, uint32_t *dst);           // run-time dispatched (AVX2+BMI2 or SWAR)
bool pack_read_portable(const uint8_t *s, uint32_t len, uint32_t *dst);  // SWAR only (tests compare the two)

}  // namespace cls
