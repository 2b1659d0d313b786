// 2-bit packing of query bases on the host (see host_pack.cpp).
#pragma once
#include <cstdint>

namespace cls {

// Packs `len` ASCII bases into ceil(len / 16) words at dst; false if any byte is not A/C/G/T/a/c/g/t.
bool pack_read(const uint8_t *s, uint32_t len, uint32_t *dst);           // run-time dispatched (AVX-512, AVX2+BMI2 or SWAR)
bool pack_read_portable(const uint8_t *s, uint32_t len, uint32_t *dst);  // SWAR only

// One named body (tests hold them all equal): 1 / 0 as pack_read, -1 if this CPU cannot run the variant.
enum { kPackAuto = 0, kPackPortable = 1, kPackAvx2 = 2, kPackAvx512 = 3 };
int pack_read_variant(int variant, const uint8_t *s, uint32_t len, uint32_t *dst);

}  // namespace cls
