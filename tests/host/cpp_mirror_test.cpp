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
        std::vector<Scalar> coeffs(n);
        for (size_t i = 0; i < n; i++) memcpy(coeffs[i].data(), &sc[32 * i], 32);
        auto dom = bellman::EvaluationDomain::from_coeffs(coeffs);
        dom.fft(ctx);
        spit(dir + "/fft", dom.coeffs.data(), dom.coeffs.size());
        dom.ifft(ctx);
        spit(dir + "/fft_roundtrip", dom.coeffs.data(), dom.coeffs.size());
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
