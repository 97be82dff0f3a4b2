// fuzz_deflate.cpp -- the built-in DEFLATE encoder / decoder of the BAM readers and writers against zlib:
//   * fastinflate on streams zlib produced (levels 0-9; default / filtered / Huffman-only / RLE / fixed strategies; several
//     blocks per stream through Z_FULL_FLUSH / Z_SYNC_FLUSH) and on fastdeflate's own output: identical bytes, nothing
//     written outside the output, a wrong expected size refused;
//   * corrupted streams: no crash, no out-of-bounds access (run under ASan / UBSan by tests/test_native_fuzz.py);
//   * with a file argument: throughput of both decoders on its 0xff00-byte blocks.
#include "../../fade_b200/csrc/host/fastinflate.hpp"
#include "../../fade_b200/csrc/host/fastdeflate.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <zlib.h>
#include <chrono>
#include <random>
static std::vector<uint8_t> zdeflate(const std::vector<uint8_t>& in, int level, int strategy, bool split) {
    std::vector<uint8_t> out(in.size() * 2 + 1024);
    z_stream zs; memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    if (split && in.size() > 10) {
        size_t a = in.size() / 3, b = in.size() * 2 / 3;
        zs.next_in = const_cast<uint8_t*>(in.data()); zs.avail_in = (uInt)a; deflate(&zs, Z_FULL_FLUSH);
        zs.avail_in = (uInt)(b - a); deflate(&zs, Z_SYNC_FLUSH);
        zs.avail_in = (uInt)(in.size() - b); deflate(&zs, Z_FINISH);
    } else { zs.next_in = const_cast<uint8_t*>(in.data()); zs.avail_in = (uInt)in.size(); deflate(&zs, Z_FINISH); }
    out.resize(zs.total_out); deflateEnd(&zs);
    return out;
}
static bool zlib_reads_back(const std::vector<uint8_t>& in) {
    std::vector<uint8_t> out(in.size() + 64), back(in.size() + 1);
    const size_t outsz = fastdeflate::compress(in.data(), in.size(), out.data(), out.size());
    if (!outsz) return false;
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = out.data(); zs.avail_in = (uInt)outsz; zs.next_out = back.data(); zs.avail_out = (uInt)back.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = zs.total_out; const bool consumed = zs.avail_in == 0;
    inflateEnd(&zs);
    return rc == Z_STREAM_END && got == in.size() && consumed && (in.empty() || memcmp(back.data(), in.data(), in.size()) == 0);
}
int main(int argc, char** argv) {
    std::mt19937_64 rng(777);
    long cases = 0, fails = 0, corrupt_ok = 0, corrupt_rej = 0;
    const int iters = argc > 2 ? atoi(argv[2]) : 1500;
    for (int it = 0; it < iters && fails < 5; ++it) {
        size_t n = (it % 40 == 0) ? 65536 - (rng() % 200) : (it % 9 == 0 ? rng() % 50 : rng() % 65537);
        std::vector<uint8_t> v(n);
        int kind = it % 8;
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
            case 0: v[i] = (uint8_t)rng(); break;
            case 1: v[i] = 0; break;
            case 2: v[i] = (uint8_t)("ACGT"[rng() & 3]); break;
            case 3: v[i] = (uint8_t)(33 + rng() % 40); break;
            case 4: v[i] = (uint8_t)(i % 251); break;
            case 5: v[i] = (i > 300 && (rng() % 100) < 95) ? v[i - 300] : (uint8_t)rng(); break;
            case 6: v[i] = (uint8_t)(rng() % 3); break;
            default: v[i] = (i > 40000 && (rng() % 100) < 90) ? v[i - 32768] : (uint8_t)(rng() % 17); break;
            }
        }
        static const int strategies[5] = { Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED };
        for (int variant = 0; variant < 4; ++variant) {
            std::vector<uint8_t> c;
            if (variant == 3) {
                if (n > 65535) continue;
                if (!zlib_reads_back(v)) { ++fails; printf("FAIL encoder it=%d kind=%d n=%zu\n", it, kind, n); }
                c.resize(n + 64); c.resize(fastdeflate::compress(v.data(), n, c.data(), c.size()));
            }
            else c = zdeflate(v, (int)(rng() % 10), strategies[rng() % 5], variant == 2);
            const size_t clen = c.size();
            c.resize(clen + 16);                                 // the BGZF trailer and the read slack
            for (int k = 0; k < 16; ++k) c[clen + k] = (uint8_t)rng();
            std::vector<uint8_t> back(n + 1, 0xAA);
            ++cases;
            const bool ok = fastinflate::inflate(c.data(), clen, back.data(), n);
            if (!ok || (n && memcmp(back.data(), v.data(), n) != 0) || back[n] != 0xAA) { ++fails; printf("FAIL it=%d kind=%d n=%zu variant=%d ok=%d\n", it, kind, n, variant, (int)ok); }
            // wrong expected size must be refused
            if (n > 0) { std::vector<uint8_t> b2(n + 8); if (fastinflate::inflate(c.data(), clen, b2.data(), n - 1)) { ++fails; printf("FAIL short accepted\n"); } }
            // corrupted stream: no crash, no write outside (ASAN / canary)
            if (clen > 4 && it % 3 == 0) {
                std::vector<uint8_t> c2 = c;
                for (int f = 0; f < 1 + (int)(rng() % 3); ++f) c2[rng() % clen] ^= (uint8_t)(1u << (rng() % 8));
                std::vector<uint8_t> b3(n + 1, 0x55);
                const bool ok3 = fastinflate::inflate(c2.data(), clen, b3.data(), n);
                if (b3[n] != 0x55) { ++fails; printf("FAIL canary\n"); }
                ok3 ? ++corrupt_ok : ++corrupt_rej;
            }
        }
    }
    printf("cases %ld fails %ld (corrupted streams: %ld decoded to something, %ld refused)\n", cases, fails, corrupt_ok, corrupt_rej);
    if (argc > 1 && argv[1][0] != '-') {
        FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> d(64u << 20); d.resize(fread(d.data(), 1, d.size(), f)); fclose(f);
        for (int enc = 0; enc < 2; ++enc) {
            std::vector<std::vector<uint8_t>> blocks; std::vector<size_t> sizes;
            for (size_t a = 0; a < d.size(); a += 0xff00) {
                std::vector<uint8_t> v(d.begin() + a, d.begin() + std::min(d.size(), a + 0xff00));
                std::vector<uint8_t> c;
                if (enc) { c.resize(v.size() + 64); c.resize(fastdeflate::compress(v.data(), v.size(), c.data(), c.size())); } else c = zdeflate(v, 6, Z_DEFAULT_STRATEGY, false);
                c.resize(c.size() + 16); blocks.push_back(c); sizes.push_back(v.size());
            }
            std::vector<uint8_t> out(0x10000);
            auto t0 = std::chrono::steady_clock::now(); long bad = 0;
            for (size_t k = 0; k < blocks.size(); ++k) bad += !fastinflate::inflate(blocks[k].data(), blocks[k].size() - 16, out.data(), sizes[k]);
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            auto t1 = std::chrono::steady_clock::now();
            for (size_t k = 0; k < blocks.size(); ++k) { z_stream zs; memset(&zs, 0, sizeof(zs)); inflateInit2(&zs, -15); zs.next_in = blocks[k].data(); zs.avail_in = (uInt)(blocks[k].size() - 16); zs.next_out = out.data(); zs.avail_out = (uInt)sizes[k]; inflate(&zs, Z_FINISH); inflateEnd(&zs); }
            double dz = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
            printf("%s blocks: fastinflate %.0f MB/s (bad %ld), zlib %.0f MB/s\n", enc ? "fastdeflate" : "zlib-6", d.size() / 1e6 / dt, bad, d.size() / 1e6 / dz);
        }
    }
    return fails != 0;
}
