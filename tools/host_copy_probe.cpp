// Host staging probe: how fast do T threads copy a pageable buffer in 4 MB chunks into a small ring of buffers --
// glibc memcpy against a loop of non-temporal (streaming) stores.  g++ -O2 -pthread tools/host_copy_probe.cpp -o build/host_copy_probe
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

__attribute__((target("avx2"))) static void copy_nt(void* dst, const void* src, size_t n) {
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
    }
    if (i < n) std::memcpy(d + i, s + i, n - i);
    _mm_sfence();
}

int main() {
    const size_t total = (size_t)65536 * 576 * 4, chunk = (size_t)4 << 20;
    char* src = static_cast<char*>(std::malloc(total));
    std::memset(src, 1, total);
    const int S = 8;
    char* ring[S];
    for (int i = 0; i < S; ++i) { ring[i] = static_cast<char*>(std::aligned_alloc(4096, chunk)); std::memset(ring[i], 0, chunk); }
    std::printf("avx2 %d, hardware threads %u\n", __builtin_cpu_supports("avx2"), std::thread::hardware_concurrency());
    for (int nt = 0; nt < 2; ++nt)
        for (int T : {1, 2, 4, 8, 16}) {
            double best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                std::atomic<size_t> next{0};
                const size_t nch = (total + chunk - 1) / chunk;
                const auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t)
                    th.emplace_back([&, t]() {
                        for (;;) {
                            const size_t j = next.fetch_add(1);
                            if (j >= nch) break;
                            const size_t off = j * chunk, n = std::min(chunk, total - off);
                            if (nt) copy_nt(ring[j % S], src + off, n); else std::memcpy(ring[j % S], src + off, n);
                        }
                    });
                for (auto& x : th) x.join();
                best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
            std::printf("%s  threads %2d  %.2f ms  %.1f GB/s\n", nt ? "streaming stores" : "memcpy          ", T, best * 1e3, total / best / 1e9);
        }
    return 0;
}
