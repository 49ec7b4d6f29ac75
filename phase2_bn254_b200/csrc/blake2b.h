// blake2b.h -- unkeyed BLAKE2b-512 (RFC 7693) for the host side of the C ABI.
// The reference hashes with blake2 0.8.1 `Blake2b::default()` (powersoftau/src/utils.rs:20-27) and blake2-rfc
// `Blake2b::new(64)` (phase2/src/hash_writer.rs:24-36): both are plain unkeyed BLAKE2b with a 64-byte digest.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace p2b {

class Blake2b {
public:
    Blake2b() {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        for (int i = 0; i < 8; i++) h_[i] = iv[i];
        h_[0] ^= 0x01010000ULL ^ 64ULL;   // digest length 64, no key, fanout = depth = 1
        t_[0] = t_[1] = 0;
        fill_ = 0;
    }
    void update(const void *data, size_t len) {
        const uint8_t *p = (const uint8_t *)data;
        while (len > 0) {
            if (fill_ == 128) {           // buffer full and more input follows: compress a non-final block
                add_counter(128);
                compress(false);
                fill_ = 0;
            }
            size_t take = 128 - fill_;
            if (take > len) take = len;
            memcpy(buf_ + fill_, p, take);
            fill_ += take;
            p += take;
            len -= take;
        }
    }
    void finish(uint8_t out[64]) {
        add_counter(fill_);
        memset(buf_ + fill_, 0, 128 - fill_);
        compress(true);
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h_[i] >> (8 * j));
    }
    static void hash(const void *data, size_t len, uint8_t out[64]) {
        Blake2b b;
        b.update(data, len);
        b.finish(out);
    }

private:
    uint64_t h_[8], t_[2];
    uint8_t buf_[128];
    size_t fill_;
    void add_counter(uint64_t n) {
        t_[0] += n;
        if (t_[0] < n) t_[1]++;
    }
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(bool last) {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; i++) {
            uint64_t w = 0;
            for (int j = 7; j >= 0; j--) w = (w << 8) | buf_[8 * i + j];
            m[i] = w;
        }
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[i + 8] = iv[i]; }
        v[12] ^= t_[0];
        v[13] ^= t_[1];
        if (last) v[14] = ~v[14];
#define P2B_B2G(a, b, c, d, x, y)                   \
    v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); \
    v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; r++) {
            const uint8_t *s = sigma[r];
            P2B_B2G(0, 4, 8, 12, m[s[0]], m[s[1]])
            P2B_B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
            P2B_B2G(2, 6, 10, 14, m[s[4]], m[s[5]])
            P2B_B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
            P2B_B2G(0, 5, 10, 15, m[s[8]], m[s[9]])
            P2B_B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
            P2B_B2G(2, 7, 8, 13, m[s[12]], m[s[13]])
            P2B_B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef P2B_B2G
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
    }
};

}  // namespace p2b
