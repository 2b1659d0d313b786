// TEST INFRASTRUCTURE.  Every body of the host 2-bit packer (classeq2_b200/csrc/host_pack.cpp) on exactly-sized heap
// buffers under AddressSanitizer: no byte is read past the read, no word is written past ceil(len / 16), and all bodies
// agree.  Built and run by tests/test_text_fuzz.py.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../classeq2_b200/csrc/host_pack.hpp"

int main() {
    uint64_t x = 88172645463325252ull;
    long bad = 0, ran = 0;
    for (uint32_t L = 0; L <= 700; ++L) {
        uint8_t *in = new uint8_t[L ? L : 1];
        for (uint32_t i = 0; i < L; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; in[i] = "ACGTacgt"[x & 7]; }
        const uint32_t nw = (L + 15) / 16;
        std::vector<uint32_t> ref;
        for (int v = 1; v <= 3; ++v) {
            uint32_t *out = new uint32_t[nw ? nw : 1];
            memset(out, 0xEE, (nw ? nw : 1) * 4);
            const int rc = cls::pack_read_variant(v, in, L, out);
            if (rc >= 0) {
                ++ran;
                if (rc != 1) ++bad;
                if (v == 1) ref.assign(out, out + nw);
                else if (nw && memcmp(ref.data(), out, nw * 4) != 0) ++bad;
                if (L) {                                   // an invalid byte anywhere is reported
                    const uint32_t p = (uint32_t)(x % L);
                    const uint8_t keep = in[p];
                    in[p] = 'N';
                    if (cls::pack_read_variant(v, in, L, out) != 0) ++bad;
                    in[p] = keep;
                }
            }
            delete[] out;
        }
        delete[] in;
    }
    printf("bad=%ld ran=%ld\n", bad, ran);
    return bad ? 1 : 0;
}
