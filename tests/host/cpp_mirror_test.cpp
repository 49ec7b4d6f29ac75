// tests/host/cpp_mirror_test.cpp -- drives the C++ host mirror (include/p2b.hpp) the way the reference's binaries drive the
// Rust items it mirrors: new_constrained -> compute_constrained (transform) -> decompress, plus dense_multiexp and an
// EvaluationDomain round trip.  Inputs and outputs are raw files so that tests/test_gpu_cpp_mirror.py can compare every
// byte with the oracle.  Exit code 3 = no CUDA device (the mirror has no CPU fallback).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include "../../include/p2b.hpp"

using namespace p2b;
using namespace p2b::powersoftau;

static std::vector<uint8_t> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void spit(const std::string &path, const uint8_t *p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char *)p, (std::streamsize)n);
}
static Scalar scalar_from(const std::vector<uint8_t> &v, size_t off) {
    Scalar s;
    memcpy(s.data(), v.data() + off, 32);
    return s;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::cerr << "usage: cpp_mirror_test <dir> <size>\n"; return 2; }
    const std::string dir = argv[1];
    const size_t size = (size_t)atoi(argv[2]);
    try {
        Context ctx(0);
        CeremonyParams params(size, 16);
        // new_constrained: blank hash prefix is the caller's; the accumulator is all generators
        std::vector<uint8_t> challenge(params.accumulator_size, 0);
        BatchedAccumulator::generate_initial(challenge.data(), challenge.size(), params);
        spit(dir + "/challenge", challenge.data(), challenge.size());
        // compute_constrained: keys from a file (tau | alpha | beta, 32-byte BE each)
        std::vector<uint8_t> keys = slurp(dir + "/keys");
        PrivateKey key{scalar_from(keys, 0), scalar_from(keys, 32), scalar_from(keys, 64)};
        std::vector<uint8_t> response(params.contribution_size, 0);
        BatchedAccumulator::transform(ctx, challenge.data(), challenge.size(), response.data(), response.size(), UseCompression::No,
                                      UseCompression::Yes, CheckForCorrectness::No, key, params);
        spit(dir + "/response", response.data(), response.size());
        // verify_transform_constrained's last step: decompress the response into the next challenge
        std::vector<uint8_t> next(params.accumulator_size, 0);
        BatchedAccumulator::decompress(ctx, response.data(), response.size() - params.public_key_size, next.data(), next.size(),
                                       CheckForCorrectness::Yes, params);
        spit(dir + "/new_challenge", next.data(), next.size());
        // prepare_phase2 for m = size
        std::vector<uint8_t> radix = prepare_phase2(ctx, response.data(), response.size() - params.public_key_size, params, (uint32_t)size);
        spit(dir + "/phase1radix2m", radix.data(), radix.size());
        // bellman: dense_multiexp over tau_g1 of the new challenge with scalars from a file, EvaluationDomain round trip
        std::vector<uint8_t> sc = slurp(dir + "/scalars");
        const size_t n = sc.size() / 32;
        auto sum = bellman::dense_multiexp_g1(ctx, next.data() + 64, sc.data(), n);
        spit(dir + "/msm", sum.data(), sum.size());
        // the verifier's power_pairs over the COMPRESSED tau_g1 section of the response, coefficients from a fixed seed
        {
            uint8_t seed[32];
            for (int i = 0; i < 32; i++) seed[i] = (uint8_t)(7 * i + 3);
            auto pp = bellman::power_pairs_g1(ctx, response.data() + 64, n, seed, 253, P2B_ENC_COMPRESSED, P2B_REJECT_INFINITY);
            std::vector<uint8_t> both(pp.s.begin(), pp.s.end());
            both.insert(both.end(), pp.sx.begin(), pp.sx.end());
            spit(dir + "/power_pairs", both.data(), both.size());
        }
        std::vector<Scalar> coeffs(n);
        for (size_t i = 0; i < n; i++) memcpy(coeffs[i].data(), &sc[32 * i], 32);
        auto dom = bellman::EvaluationDomain::from_coeffs(coeffs);
        dom.fft(ctx);
        spit(dir + "/fft", dom.coeffs.data(), dom.coeffs.size());
        dom.ifft(ctx);
        spit(dir + "/fft_roundtrip", dom.coeffs.data(), dom.coeffs.size());
        // verifier host side: tau_g1[0..2] of the new challenge is (G, tau G) and tau_g2[0..2] is (H, tau H): same ratio, and not
        // the other way round (test_same_ratio_bn256, powersoftau/src/utils.rs:76-88)
        {
            const uint8_t *tg1 = next.data() + 64, *tg2 = next.data() + 64 + params.powers_g1_length * 64;
            if (!p2b::same_ratio(tg1, tg1 + 64, tg2, tg2 + 128) || p2b::same_ratio(tg1 + 64, tg1, tg2, tg2 + 128)) {
                std::cerr << "same_ratio mismatch\n";
                return 1;
            }
            // QAP-style sparse evaluation over tau_g1: row 0 = 1*P0 + 1*P1, row 1 empty, row 2 = s0*P2
            std::vector<uint64_t> ro = {0, 2, 2, 3};
            std::vector<uint32_t> cols = {0, 1, 2};
            Scalar one{};
            one[31] = 1;
            std::vector<Scalar> cf = {one, one, coeffs[0]};
            auto ev = phase2::sparse_eval_g1(ctx, tg1, params.powers_g1_length, ro, cols, cf);
            spit(dir + "/sparse", ev.data(), ev.size());
            // the key-generation RNG is deterministic in its seed
            const uint32_t seed[8] = {1, 2, 3, 4, 5, 6, 7, 8};
            p2b::ChaChaRng r1(seed), r2(seed);
            if (r1.gen_fr() != r2.gen_fr() || r1.gen_g1() != r2.gen_g1() || r1.gen_g2() != r2.gen_g2()) return 1;
            auto hg = p2b::hash_to_g2(response.data());
            spit(dir + "/hash_to_g2", hg.data(), hg.size());
        }
        // error behaviour: a point at infinity in the input is DeserializationError::PointAtInfinity
        challenge[64] = 0x40;
        memset(challenge.data() + 65, 0, 63);
        try {
            BatchedAccumulator::transform(ctx, challenge.data(), challenge.size(), response.data(), response.size(), UseCompression::No,
                                          UseCompression::Yes, CheckForCorrectness::No, key, params);
            std::cerr << "expected PointAtInfinity\n";
            return 1;
        } catch (const DeserializationError &e) {
            if (e.kind != DeserializationError::PointAtInfinity || e.index != 0) return 1;
        }
        std::cout << "cpp mirror ok\n";
        return 0;
    } catch (const Error &e) {
        std::cerr << "p2b error " << e.code << ": " << e.what() << "\n";
        return e.code == P2B_ECUDA ? 3 : 1;
    }
}
