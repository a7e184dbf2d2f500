// CPU check of the staging copy of the host-buffer pipeline (myldpccppapi_b200/csrc/ldpc_tables.cpp: stage_copy_nt):
// every source / destination offset modulo a cache line and sizes around the 64- and 128-byte steps give memcpy's bytes,
// and nothing outside [dst, dst + n) changes.  Built and run by tests/test_host.py (no GPU needed).
#include <cstdio>
#include <cstring>
#include <vector>

#include "ldpc_tables.h"

int main() {
    std::vector<char> a(1 << 18), b(1 << 18);
    for (size_t i = 0; i < a.size(); ++i) a[i] = (char)(i * 131 + 7);
    int bad = 0, cases = 0;
    const size_t sizes[] = {0, 1, 15, 63, 64, 65, 127, 128, 129, 255, 256, 4095, 4096, 4097, 65536 + 17};
    for (size_t so = 0; so < 67; ++so)
        for (size_t d0 = 0; d0 < 67; d0 += 3)
            for (size_t n : sizes) {
                std::memset(b.data(), 0x5a, 512 + d0 + n + 512);
                ldpc_b200::stage_copy_nt(b.data() + 512 + d0, a.data() + so, n);
                ++cases;
                if (std::memcmp(b.data() + 512 + d0, a.data() + so, n)) ++bad;
                for (size_t k = 0; k < 512 + d0; ++k) if (b[k] != 0x5a) { ++bad; break; }
                for (size_t k = 512 + d0 + n; k < 512 + d0 + n + 512; ++k) if (b[k] != 0x5a) { ++bad; break; }
            }
    std::printf("cases=%d bad=%d\n", cases, bad);
    return bad != 0;
}
