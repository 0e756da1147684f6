// Host-only component timing of worker_verify (decompression, scalar multiplication, Miller loops, final exponentiation).
// g++ -O3 -std=c++17 -o build/verify_bench tools/verify_bench.cpp ; runs without a GPU.
#include "../zkp_subnet_b200/csrc/host/pairing.hpp"
#include <chrono>
#include <cstdio>
using namespace zkp::host;
template <class F> double us(F f, int reps) {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) f();
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}
int main() {
    G1J g = g1_generator();
    uint64_t k[4] = {0x123456789abcdefull, 0xfedcba987654321ull, 0x1111111122222222ull, 0x0123456701234567ull};
    G1J p = g.mul(k, 4);
    uint8_t c48[48];
    g1_compress(c48, p);
    G1J q;
    volatile bool sink = false;
    printf("g1_decompress (with subgroup check): %.1f us\n", us([&] { sink = g1_decompress(q, c48); }, 200));
    printf("g1_decompress (no subgroup check):   %.1f us\n", us([&] { sink = g1_decompress(q, c48, false); }, 200));
    printf("g1 scalar mul 255 bit:               %.1f us\n", us([&] { q = p.mul(k, 4); }, 100));
    printf("g1 scalar mul 255 bit, 4-bit windows: %.1f us\n", us([&] { q = p.mul_w4(k, 4); }, 100));
    G2Lines l1 = g2_precompute(g2_generator()), l2 = g2_precompute(g2_generator());
    std::vector<G1AffineHost> ps = {g1_affine_host(p), g1_affine_host(p.neg())};
    std::vector<const G2Lines*> qs = {&l1, &l2};
    printf("pairing_product_is_one (2 pairs):    %.1f us (%d)\n", us([&] { sink = pairing_product_is_one(ps, qs); }, 50), (int)sink);
    { Fq12 f = Fq12::one(); f.c0.c0.c0 = p.x; f.c1.c1.c1 = p.y; f.c0.c2.c0 = p.y; printf("final_exponentiation:                %.1f us\n", us([&] { f = final_exponentiation(f); }, 50)); }
    printf("g1_affine_host:                      %.1f us\n", us([&] { ps[0] = g1_affine_host(q); }, 200));
}
