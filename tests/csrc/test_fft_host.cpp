// Host harness for lpc_fft.cuh: runs the SAME in-register DFTs, Stockham pass index math,
// padding and twiddle tables on the CPU (threads emulated by loops, the CTA barrier by the
// loop boundary between the load+compute phase and the store phase) and compares against a
// naive O(N^2) DFT in long double.  Built and run by tests/test_fft_host.py (no GPU needed).
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "lpc_fft_plan.h"

using namespace lpc;

template <typename T, int DIR> static void naive(const std::vector<std::complex<long double>>& x, std::vector<std::complex<long double>>& X) {
    const int N = (int)x.size();
    X.assign(N, 0);
    for (int k = 0; k < N; ++k) {
        std::complex<long double> acc = 0;
        for (int n = 0; n < N; ++n) {
            long double ang = DIR * 2.0L * M_PIl * (long double)(((long long)n * k) % N) / N;
            acc += x[n] * std::complex<long double>(cosl(ang), sinl(ang));
        }
        X[k] = acc;
    }
}

template <typename T, int R, int DIR> static double test_dft() {
    using C = typename CxT<T>::type;
    C v[R];
    std::vector<std::complex<long double>> x(R), X;
    for (int i = 0; i < R; ++i) {
        v[i].x = (T)(rand() / (double)RAND_MAX - 0.5);
        v[i].y = (T)(rand() / (double)RAND_MAX - 0.5);
        x[i] = {v[i].x, v[i].y};
    }
    naive<T, DIR>(x, X);
    Dft<T, R, DIR>::run(v);
    double err = 0;
    for (int i = 0; i < R; ++i) err = fmax(err, (double)std::abs(std::complex<long double>(v[i].x, v[i].y) - X[i]));
    return err;
}

template <typename T, int R, int DIR>
static void emu_pass(std::vector<typename CxT<T>::type>& sm, const typename CxT<T>::type* tw, const FftPlan& plan, int Ns, int nbatch, int bs, int es) {
    using C = typename CxT<T>::type;
    const int nb = plan.N / R, total = nb * nbatch;
    std::vector<C> regs((size_t)total * R);
    for (int g = 0; g < total; ++g) {  // "threads": load + compute phase
        int b, j;
        if (es > 1) { b = g % nbatch; j = g / nbatch; } else { b = g / nb; j = g - b * nb; }
        fft_bfly_load<T, R>(sm.data(), b * bs, es, plan.padsh, plan.N, j, &regs[(size_t)g * R]);
        fft_bfly_compute<T, R, DIR>(&regs[(size_t)g * R], tw, Ns, j);
    }
    for (int g = 0; g < total; ++g) {  // barrier, then store phase
        int b, j;
        if (es > 1) { b = g % nbatch; j = g / nbatch; } else { b = g / nb; j = g - b * nb; }
        fft_bfly_store<T, R>(sm.data(), b * bs, es, plan.padsh, Ns, j, &regs[(size_t)g * R]);
    }
}

template <typename T, int DIR>
static void emu_fft(std::vector<typename CxT<T>::type>& sm, const FftPlan& plan, const std::vector<typename CxT<T>::type>& tw, int nbatch, int bs, int es) {
    for (int p = 0; p < plan.npass; ++p) {
        const auto* t = tw.data() + plan.tw_off[p];
        switch (plan.radix[p]) {
#define CASE(R) case R: emu_pass<T, R, DIR>(sm, t, plan, plan.Ns[p], nbatch, bs, es); break;
            CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(8) CASE(9) CASE(10) CASE(12) CASE(15) CASE(16)
#undef CASE
            default: printf("bad radix %d\n", plan.radix[p]); exit(2);
        }
    }
}

template <typename T, int DIR> static double test_fft(int N, int nbatch, bool interleaved) {
    using C = typename CxT<T>::type;
    FftPlan plan;
    if (!make_fft_plan(N, &plan)) { printf("no plan for %d\n", N); exit(2); }
    std::vector<C> tw;
    fill_fft_twiddles(plan, &tw);
    const int es = interleaved ? nbatch : 1;
    const int bs = interleaved ? 1 : N;  // batch stride in LOGICAL elements (padding applies to the sum)
    std::vector<C> sm(fft_padded_len(N * nbatch, plan.padsh) + 8);
    std::vector<std::vector<std::complex<long double>>> x(nbatch, std::vector<std::complex<long double>>(N));
    for (int b = 0; b < nbatch; ++b)
        for (int i = 0; i < N; ++i) {
            C c; c.x = (T)(rand() / (double)RAND_MAX - 0.5); c.y = (T)(rand() / (double)RAND_MAX - 0.5);
            sm[fft_pad(b * bs + i * es, plan.padsh)] = c;
            x[b][i] = {c.x, c.y};
        }
    emu_fft<T, DIR>(sm, plan, tw, nbatch, bs, es);
    double err = 0, mx = 0;
    for (int b = 0; b < nbatch; ++b) {
        std::vector<std::complex<long double>> X;
        naive<T, DIR>(x[b], X);
        for (int i = 0; i < N; ++i) {
            C c = sm[fft_pad(b * bs + i * es, plan.padsh)];
            err = fmax(err, (double)std::abs(std::complex<long double>(c.x, c.y) - X[i]));
            mx = fmax(mx, (double)std::abs(X[i]));
        }
    }
    printf("N=%5d batch=%d %s dir=%+d radices:", N, nbatch, interleaved ? "interleaved" : "planar     ", DIR);
    for (int p = 0; p < plan.npass; ++p) printf(" %d", plan.radix[p]);
    printf("  padsh=%d tw=%d  rel err %.2e\n", plan.padsh, plan.tw_len, err / mx);
    return err / mx;
}


// ---- in-place decimation passes (fft_ip_bfly): forward DIF -> permuted spectrum -> inverse DIT
template <typename T, int R, int DIR, bool DIF>
static void emu_ip_pass(std::vector<typename CxT<T>::type>& sm, const typename CxT<T>::type* tw, int N, int M, int nbatch) {
    const int total = (N / R) * nbatch;
    for (int g = 0; g < total; ++g) {  // butterflies touch disjoint addresses: any order is valid
        const int b = g % nbatch, j = g / nbatch;
        fft_ip_bfly<T, R, DIR, DIF>(sm.data() + b, nbatch, M, j, tw);
    }
}
template <typename T, int DIR, bool DIF>
static void emu_ip_dispatch(int R, std::vector<typename CxT<T>::type>& sm, const typename CxT<T>::type* tw, int N, int M, int nbatch) {
    switch (R) {
#define CASE(RR) case RR: emu_ip_pass<T, RR, DIR, DIF>(sm, tw, N, M, nbatch); break;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(8) CASE(9) CASE(10) CASE(12) CASE(15) CASE(16)
#undef CASE
        default: printf("bad radix %d\n", R); exit(2);
    }
}
template <typename T> static double test_ip(int N, int nbatch) {
    using C = typename CxT<T>::type;
    FftPlan plan;
    if (!make_fft_plan(N, &plan)) { printf("no plan for %d\n", N); exit(2); }
    std::vector<C> tw;
    if (!fill_fft_ip_twiddles(N, plan.radix, plan.npass, &tw)) { printf("no ip twiddles for %d\n", N); exit(2); }
    std::vector<C> sm((size_t)N * nbatch);
    std::vector<std::vector<std::complex<long double>>> x(nbatch, std::vector<std::complex<long double>>(N));
    for (int b = 0; b < nbatch; ++b)
        for (int i = 0; i < N; ++i) {
            C c; c.x = (T)(rand() / (double)RAND_MAX - 0.5); c.y = (T)(rand() / (double)RAND_MAX - 0.5);
            sm[(size_t)i * nbatch + b] = c;
            x[b][i] = {c.x, c.y};
        }
    std::vector<int> off(plan.npass), Ms(plan.npass);
    int M = N, o = 0;
    for (int i = 0; i < plan.npass; ++i) { Ms[i] = M; off[i] = o; o += (plan.radix[i] - 1) * (M / plan.radix[i]); M /= plan.radix[i]; }
    for (int i = 0; i < plan.npass; ++i) emu_ip_dispatch<T, -1, true>(plan.radix[i], sm, tw.data() + off[i], N, Ms[i], nbatch);
    double err = 0, mx = 0;
    for (int b = 0; b < nbatch; ++b) {
        std::vector<std::complex<long double>> X;
        naive<T, -1>(x[b], X);
        for (int p = 0; p < N; ++p) {
            const int k = fft_ip_freq_of_pos(N, plan.radix, plan.npass, p);
            C c = sm[(size_t)p * nbatch + b];
            err = fmax(err, (double)std::abs(std::complex<long double>(c.x, c.y) - X[k]));
            mx = fmax(mx, (double)std::abs(X[k]));
        }
    }
    // mirror image back: innermost pass first
    for (int i = plan.npass - 1; i >= 0; --i) emu_ip_dispatch<T, +1, false>(plan.radix[i], sm, tw.data() + off[i], N, Ms[i], nbatch);
    double err2 = 0;
    for (int b = 0; b < nbatch; ++b)
        for (int i = 0; i < N; ++i) {
            C c = sm[(size_t)i * nbatch + b];
            err2 = fmax(err2, (double)std::abs(std::complex<long double>(c.x / N, c.y / N) - x[b][i]));
        }
    printf("N=%5d batch=%d in-place DIF/DIT radices:", N, nbatch);
    for (int p = 0; p < plan.npass; ++p) printf(" %d", plan.radix[p]);
    printf("  fwd rel err %.2e  round-trip err %.2e\n", err / mx, err2);
    return fmax(err / mx, err2);
}

// ---- the warp-local pair of outermost passes (fft_ip_outer2_phase): classes walked ONE AFTER THE OTHER, each through its
//      two phases -- correct only if a class really depends on nothing outside itself.  WIN: zero padding outside [lo, hi):
//      the forward side must not read outside the window (garbage is planted there), the inverse side only has to
//      deliver the window.
template <typename T, int N, int R1, int R2, int R3, bool WIN> static double test_ip_outer2() {
    using C = typename CxT<T>::type;
    constexpr int NB = 4, M1 = N / R1, S2 = M1 / R2, LPB = 8;
    static_assert(R1 * R2 * R3 == N, "three-pass sequence");
    const int radix[3] = {R1, R2, R3};
    std::vector<C> tw;
    if (!fill_fft_ip_twiddles(N, radix, 3, &tw)) { printf("no ip twiddles for %d\n", N); exit(2); }
    const int lo = WIN ? N / 4 + 1 : 0, hi = WIN ? N / 4 + N / 2 : N;
    std::vector<C> sm((size_t)N * NB);
    std::vector<std::vector<std::complex<long double>>> x(NB, std::vector<std::complex<long double>>(N));
    for (int b = 0; b < NB; ++b)
        for (int i = 0; i < N; ++i) {
            C c; c.x = (T)(rand() / (double)RAND_MAX - 0.5); c.y = (T)(rand() / (double)RAND_MAX - 0.5);
            sm[(size_t)i * NB + b] = c;
            x[b][i] = (i >= lo && i < hi) ? std::complex<long double>(c.x, c.y) : std::complex<long double>(0, 0);
        }
    const C* tw3 = tw.data() + (R1 - 1) * (N / R1) + (R2 - 1) * S2;
    auto outer = [&](auto dif) {
        constexpr bool DIF = decltype(dif)::value;
        for (int c = 0; c < S2; ++c)
            for (int phase = 0; phase < 2; ++phase)
                for (int b = 0; b < NB; ++b)
                    for (int li = 0; li < LPB; ++li)
                        fft_ip_outer2_phase<T, DIF ? -1 : +1, DIF, NB, N, R1, R2, WIN>(sm.data(), tw.data(), phase, c, b, li, LPB, lo, hi);
    };
    outer(std::true_type{});
    emu_ip_pass<T, R3, -1, true>(sm, tw3, N, R3, NB);
    double err = 0, mx = 0;
    for (int b = 0; b < NB; ++b) {
        std::vector<std::complex<long double>> X;
        naive<T, -1>(x[b], X);
        for (int p = 0; p < N; ++p) {
            const int k = fft_ip_freq_of_pos(N, radix, 3, p);
            C c = sm[(size_t)p * NB + b];
            err = fmax(err, (double)std::abs(std::complex<long double>(c.x, c.y) - X[k]));
            mx = fmax(mx, (double)std::abs(X[k]));
        }
    }
    emu_ip_pass<T, R3, +1, false>(sm, tw3, N, R3, NB);
    outer(std::false_type{});
    double err2 = 0;
    for (int b = 0; b < NB; ++b)
        for (int i = lo; i < hi; ++i) {
            C c = sm[(size_t)i * NB + b];
            err2 = fmax(err2, (double)std::abs(std::complex<long double>(c.x / N, c.y / N) - x[b][i]));
        }
    printf("N=%5d warp-local outer passes (%d,%d | %d) %s  fwd rel err %.2e  round-trip err %.2e\n", N, R1, R2, R3,
           WIN ? "windowed" : "full    ", err / mx, err2);
    return fmax(err / mx, err2);
}

int main() {
    int bad = 0;
    double e;
#define TD(R) \
    e = test_dft<double, R, -1>(); if (e > 1e-14) { printf("dft<double,%d,fwd> err %.2e\n", R, e); bad++; } \
    e = test_dft<double, R, +1>(); if (e > 1e-14) { printf("dft<double,%d,inv> err %.2e\n", R, e); bad++; } \
    e = test_dft<float, R, -1>(); if (e > 5e-6) { printf("dft<float,%d,fwd> err %.2e\n", R, e); bad++; }
    TD(2) TD(3) TD(4) TD(5) TD(6) TD(8) TD(9) TD(10) TD(12) TD(15) TD(16)
    printf("in-register DFTs: %s\n", bad ? "FAILED" : "ok");
    const int sizes[] = {2, 3, 4, 5, 6, 8, 9, 16, 27, 40, 45, 48, 54, 60, 64, 75, 96, 125, 270, 405, 480, 768, 1024, 2160, 3840};
    for (int N : sizes) {
        if (test_fft<double, -1>(N, 1, false) > 1e-13) bad++;
        if (test_fft<double, +1>(N, 2, false) > 1e-13) bad++;
        if (test_fft<float, -1>(N, 4, true) > 2e-6) bad++;
        if (test_fft<float, +1>(N, 3, true) > 2e-6) bad++;
    }
    for (int N : sizes) {
        if (test_ip<double>(N, 1) > 1e-13) bad++;
        if (test_ip<float>(N, 4) > 2e-6) bad++;
    }
#define TW(N, A, B, C3) \
    if (test_ip_outer2<double, N, A, B, C3, false>() > 1e-13) bad++; \
    if (test_ip_outer2<float, N, A, B, C3, false>() > 2e-6) bad++;   \
    if (test_ip_outer2<double, N, A, B, C3, true>() > 1e-13) bad++;  \
    if (test_ip_outer2<float, N, A, B, C3, true>() > 2e-6) bad++;
    TW(2160, 16, 15, 9) TW(1080, 8, 15, 9) TW(405, 15, 9, 3) TW(400, 16, 5, 5) TW(2160, 15, 16, 9)
#undef TW
    printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
    return bad ? 1 : 0;
}
