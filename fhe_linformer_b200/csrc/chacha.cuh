// chacha.cuh -- ChaCha20 block function (RFC 8439 core, 64-bit block counter + 64-bit nonce as in the original cipher), host and
// device.  The production key / error / encryption randomness of the scheme layer is ChaCha20 in counter mode keyed from the
// operating system (getrandom), with independent keys per purpose; the SplitMix64 streams stay behind the seeded, test-only entry
// points (fl_keygen_seeded, fl_encrypt_seeded) that the bit-exact parity tests need.  Stands in for the CSPRNG OpenFHE draws from
// inside KeyGen / Encrypt (/root/reference/src/FHEController.cpp:47,49,248,378).
#pragma once
#include <cstdint>

namespace flk {

struct ChaChaKey {
    uint32_t k[8];
};

#if defined(__CUDACC__)
#define FLK_HD __host__ __device__ __forceinline__
#else
#define FLK_HD inline
#endif

FLK_HD uint32_t chacha_rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }

#define FLK_CHACHA_QR(a, b, c, d)                                                                        \
    a += b; d ^= a; d = chacha_rotl(d, 16); c += d; b ^= c; b = chacha_rotl(b, 12);                       \
    a += b; d ^= a; d = chacha_rotl(d, 8);  c += d; b ^= c; b = chacha_rotl(b, 7);

// 64-byte key-stream block number `counter` of stream `nonce`
FLK_HD void chacha20_block(const ChaChaKey& key, uint64_t counter, uint64_t nonce, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                      key.k[4],    key.k[5],    key.k[6],    key.k[7],    (uint32_t)counter, (uint32_t)(counter >> 32),
                      (uint32_t)nonce, (uint32_t)(nonce >> 32)};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = s[i];
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        FLK_CHACHA_QR(x[0], x[4], x[8], x[12]) FLK_CHACHA_QR(x[1], x[5], x[9], x[13]) FLK_CHACHA_QR(x[2], x[6], x[10], x[14])
        FLK_CHACHA_QR(x[3], x[7], x[11], x[15]) FLK_CHACHA_QR(x[0], x[5], x[10], x[15]) FLK_CHACHA_QR(x[1], x[6], x[11], x[12])
        FLK_CHACHA_QR(x[2], x[7], x[8], x[13]) FLK_CHACHA_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

// words 2k, 2k+1 of a block as one 64-bit value
FLK_HD uint64_t chacha_u64(const uint32_t blk[16], int k) { return (uint64_t)blk[2 * k] | ((uint64_t)blk[2 * k + 1] << 32); }

// |X| cumulative distribution of the discrete Gaussian, sigma = 3.19, scaled to 2^64 (DESIGN.md "Randomness"); one copy for the
// host and one initialised in the device image (no upload, hence no ordering hazard with the engine stream)
#define FLK_GAUSS_CDT_INIT                                                                                                       \
    {0x2003F343659528D0ull, 0x5CF9E7DE0F06D96Bull, 0x9194FD0BB0694AF3ull, 0xBABAA2EF1EEC1101ull, 0xD7E6AB30AA084360ull,          \
     0xEAA5B92100F77DABull, 0xF591040AC34992E9ull, 0xFB54CB2CA496FFFAull, 0xFE1702297749D972ull, 0xFF4953F8BD4AE9D1ull,          \
     0xFFC1C20EF5DE7233ull, 0xFFECAC7F021E2BA0ull, 0xFFFA892133378B29ull, 0xFFFE9810099A70DBull, 0xFFFFABC31CFFB430ull,          \
     0xFFFFEE138CF385DBull, 0xFFFFFC88B8F21ED7ull, 0xFFFFFF641EF54A94ull, 0xFFFFFFE7207A46BAull, 0xFFFFFFFC6560DA3Aull,          \
     0xFFFFFFFF86A24B98ull, 0xFFFFFFFFF1822EC9ull, 0xFFFFFFFFFE6DFC66ull, 0xFFFFFFFFFFD877EFull, 0xFFFFFFFFFFFC7916ull,          \
     0xFFFFFFFFFFFFB6EAull, 0xFFFFFFFFFFFFFAA2ull, 0xFFFFFFFFFFFFFFA4ull, 0xFFFFFFFFFFFFFFFAull, 0xFFFFFFFFFFFFFFFFull}

}  // namespace flk
