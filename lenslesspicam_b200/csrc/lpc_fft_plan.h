// lpc_fft_plan.h -- host-side planner for lpc_fft.cuh: radix sequence and twiddle tables.
#pragma once
#include <cmath>
#include <vector>

#include "lpc_fft.cuh"

namespace lpc {

// Factorise N (5-smooth) into radices from {16,15,12,10,9,8,6,5,4,3,2} with the fewest passes
// whose radix does not exceed max_radix; an odd radix goes first when there is one (see
// fft_pad), the rest in decreasing order.  Returns false if N is not 5-smooth.
inline bool make_fft_plan(int N, FftPlan* plan, int max_radix = 16) {
    static const int cand[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};
    if (N < 1) return false;
    // dynamic programme over divisors: best[n] = fewest radices whose product is n
    std::vector<int> best(N + 1, 1 << 20), choice(N + 1, 0);
    best[1] = 0;
    for (int n = 2; n <= N; ++n) {
        if (N % n) continue;
        for (int r : cand) {
            if (r > max_radix || n % r) continue;
            if (best[n / r] + 1 < best[n]) { best[n] = best[n / r] + 1; choice[n] = r; }
        }
    }
    if (N > 1 && best[N] >= (1 << 20)) return false;
    std::vector<int> rad;
    for (int n = N; n > 1; n /= choice[n]) rad.push_back(choice[n]);
    if ((int)rad.size() > LPC_FFT_MAX_PASSES) return false;
    // order: decreasing, then move the largest odd radix to the front
    for (size_t i = 0; i < rad.size(); ++i)
        for (size_t j = i + 1; j < rad.size(); ++j)
            if (rad[j] > rad[i]) { int t = rad[i]; rad[i] = rad[j]; rad[j] = t; }
    int odd = -1;
    for (size_t i = 0; i < rad.size(); ++i)
        if (rad[i] & 1) { odd = (int)i; break; }
    if (odd > 0) {
        int r = rad[odd];
        rad.erase(rad.begin() + odd);
        rad.insert(rad.begin(), r);
    }
    plan->N = N;
    plan->npass = (int)rad.size();
    plan->padsh = (odd >= 0 || rad.empty()) ? 30 : 4;
    int Ns = 1, off = 0;
    for (int p = 0; p < plan->npass; ++p) {
        plan->radix[p] = rad[p];
        plan->Ns[p] = Ns;
        plan->tw_off[p] = off;
        off += (p == 0) ? 0 : Ns * (rad[p] - 1);
        Ns *= rad[p];
    }
    plan->tw_len = off > 0 ? off : 1;
    return true;
}

// plan from an explicit radix sequence (the compile-time FftSeq menus of lpc_fused.cuh)
inline bool make_fft_plan_from(int N, const int* radices, int npass, FftPlan* plan) {
    if (npass < 1 || npass > LPC_FFT_MAX_PASSES) return false;
    long long prod = 1;
    for (int p = 0; p < npass; ++p) prod *= radices[p];
    if (prod != N) return false;
    plan->N = N;
    plan->npass = npass;
    plan->padsh = (radices[0] & 1) ? 30 : 4;
    int Ns = 1, off = 0;
    for (int p = 0; p < npass; ++p) {
        plan->radix[p] = radices[p];
        plan->Ns[p] = Ns;
        plan->tw_off[p] = off;
        off += (Ns == 1) ? 0 : Ns * (radices[p] - 1);
        Ns *= radices[p];
    }
    plan->tw_len = off > 0 ? off : 1;
    return true;
}

// forward-sign tables: pass p, entry [(m-1)*Ns + k] = exp(-2*pi*i*m*k/(Ns*R)); computed in double
template <typename C>
inline void fill_fft_twiddles(const FftPlan& plan, std::vector<C>* out) {
    out->assign(plan.tw_len, C{});
    for (int p = 0; p < plan.npass; ++p) {
        const int R = plan.radix[p], Ns = plan.Ns[p];
        if (Ns == 1) continue;
        for (int m = 1; m < R; ++m)
            for (int k = 0; k < Ns; ++k) {
                const long long num = ((long long)m * k) % ((long long)Ns * R);
                const double ang = -2.0 * M_PI * (double)num / ((double)Ns * R);
                C w;
                w.x = (decltype(w.x))std::cos(ang);
                w.y = (decltype(w.y))std::sin(ang);
                (*out)[plan.tw_off[p] + (m - 1) * Ns + k] = w;
            }
    }
}

// twiddle tables of the in-place decimation passes (fft_ip_bfly): pass i over blocks of size
// M_i = N / (r[0]*...*r[i-1]) with stride S_i = M_i / r[i] owns (r[i]-1)*S_i entries at offset
// sum_{i'<i} (r[i']-1)*S_i', entry [(k1-1)*S + n2] = exp(-2*pi*i*n2*k1/M); computed in double
template <typename C>
inline bool fill_fft_ip_twiddles(int N, const int* r, int npass, std::vector<C>* out) {
    long long prod = 1;
    for (int i = 0; i < npass; ++i) prod *= r[i];
    if (prod != N) return false;
    out->clear();
    int M = N;
    for (int i = 0; i < npass; ++i) {
        const int R = r[i], S = M / R;
        for (int k1 = 1; k1 < R; ++k1)
            for (int n2 = 0; n2 < S; ++n2) {
                const long long num = ((long long)n2 * k1) % M;
                const double ang = -2.0 * M_PI * (double)num / (double)M;
                C w;
                w.x = (decltype(w.x))std::cos(ang);
                w.y = (decltype(w.y))std::sin(ang);
                out->push_back(w);
            }
        M = S;
    }
    if (out->empty()) out->push_back(C{});
    return true;
}

}  // namespace lpc
