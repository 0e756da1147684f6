// Host-side (emulated carry flag) exerciser for ff.cuh: prints operands and results as hex so a
// Python big-int checker (tests/test_host_limbs.py) can verify the limb algorithms without a GPU.
#include <cstdio>
#include <cstdlib>
#include "../../zkp_subnet_b200/csrc/ff.cuh"
using namespace zkp;
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
template <class F> static void pr(const char* tag, const F& a) {
    printf("%s ", tag);
    for (int i = F::N - 1; i >= 0; i--) printf("%08x", a.v[i]);
    printf("\n");
}
template <class F, class P> static F rand_fe(int mode) {
    F a;
    for (int i = 0; i < F::N; i++) a.v[i] = (uint32_t)rnd();
    if (mode == 2) for (int i = 0; i < F::N; i++) a.v[i] = 0;
    if (mode == 2) a.v[0] = (uint32_t)(rnd() % 3);
    if (mode == 3) for (int i = 0; i < F::N; i++) a.v[i] = (rnd() & 1) ? 0xffffffffu : 0u;
    // reduce into [0,p): clear top bits then conditional subtract
    { uint32_t m = P::MOD[F::N - 1] >> 1; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; a.v[F::N - 1] &= m >> 1; }
    if (mode == 1) {  // p - k, k in 1..3, with borrow
        uint32_t k = 1 + (uint32_t)(rnd() % 3);
        for (int i = 0; i < F::N; i++) { uint32_t m = P::MOD[i]; a.v[i] = m - k; k = m < k ? 1 : 0; }
    }
    return a;
}
template <class F, class P> static void run(const char* name, int iters) {
    for (int it = 0; it < iters; it++) {
        F a = rand_fe<F, P>(it % 7 == 3 ? 1 : it % 11 == 5 ? 2 : it % 13 == 6 ? 3 : 0);
        F b = rand_fe<F, P>(it % 5 == 4 ? 1 : it % 17 == 7 ? 2 : 0);
        printf("field %s\n", name);
        pr("a", a); pr("b", b);
        pr("mul", a * b); pr("add", a + b); pr("sub", a - b); pr("neg", a.neg());
        pr("sqr", a.sqr()); pr("mul2", F::mul2(a, b, b, a.sqr())); pr("tom", a.to_mont()); pr("fromm", a.from_mont());
        if (it < 8) pr("inv", a.inverse());
    }
}
int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 200;
    run<Fq, FqParams>("fq", iters);
    run<Fr, FrParams>("fr", iters);
    return 0;
}
